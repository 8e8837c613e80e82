#!/bin/bash
# ncu_pick.sh <rep> : the handful of raw metrics the roofline discussion needs, from an .ncu-rep
ncu -i "$1" --page raw --csv 2>/dev/null | python3 -c '
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; units=rows[1]
want=["gpu__time_duration.sum","sm__cycles_elapsed.max","dram__bytes_read.sum","dram__bytes_write.sum","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active","lts__t_sector_hit_rate.pct","lts__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__m_xbar2l1tex_read_bytes.sum","lts__t_bytes.sum","lts__t_sectors_srcunit_tex_op_read.sum","sm__throughput.avg.pct_of_peak_sustained_elapsed","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","launch__registers_per_thread","launch__grid_size","launch__block_size","launch__shared_mem_per_block_dynamic","smsp__inst_executed.sum","sm__inst_executed_pipe_uniform.sum","smsp__cycles_active.avg","launch__cluster_size"]
for r in rows[2:]:
    name=r[hdr.index("Kernel Name")][:60]
    print("==",name)
    for w in want:
        if w in hdr:
            i=hdr.index(w); print(f"{w:75s} {r[i]:>18s} {units[i]}")
'
