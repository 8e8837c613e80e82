"""Per-launch summary of an `ncu -i X.ncu-rep --page raw --csv` export (the .ncu-rep itself is too large to bring back from
the GPU box): kernel, grid/block, duration, DRAM bytes read/written, achieved DRAM GB/s, L2 and SM throughput, occupancy.
Usage: python tools/ncu_raw_summary.py gpurun_out/r2_ncu_layout_kernels.csv > profiles/r2_ncu_layout_kernels.txt"""
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "us", 1e3 if False else None),
]


def f(row, idx, name, default=0.0):
    i = idx.get(name)
    if i is None or row[i] == "":
        return default
    return float(row[i].replace(",", ""))


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    unit = {h: u for h, u in zip(hdr, units)}

    def in_bytes(r, name):   # ncu prints adaptive units per column
        v = f(r, idx, name)
        u = unit.get(name, "byte").lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

    def in_us(r, name):
        v = f(r, idx, name)
        u = unit.get(name, "ns").lower()
        return v * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "nsecond": 1e-3}.get(u, 1)

    print(f"# {path}: {len(data)} launches (ncu --set full, --clock-control none; cold-cache, serialised launches)")
    print(f"{'kernel':44s} {'grid':>7s} {'blk':>5s} {'us':>8s} {'rdMB':>8s} {'wrMB':>8s} {'GB/s':>8s} {'dram%':>6s} {'l2%':>6s} {'sm%':>6s} {'occ%':>6s} {'regs':>5s}")
    agg = {}
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0][:44]
        us = in_us(r, "gpu__time_duration.sum")
        rd, wr = in_bytes(r, "dram__bytes_read.sum"), in_bytes(r, "dram__bytes_write.sum")
        gbs = (rd + wr) / (us * 1e-6) / 1e9 if us else 0
        dram = f(r, idx, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
        l2 = f(r, idx, "lts__throughput.avg.pct_of_peak_sustained_elapsed")
        sm = f(r, idx, "sm__throughput.avg.pct_of_peak_sustained_elapsed")
        occ = f(r, idx, "sm__warps_active.avg.pct_of_peak_sustained_active")
        regs = f(r, idx, "launch__registers_per_thread")
        print(f"{name:44s} {r[idx['Grid Size']]:>7s} {r[idx['Block Size']]:>5s} {us:8.2f} {rd / 1e6:8.3f} {wr / 1e6:8.3f} {gbs:8.1f} {dram:6.1f} {l2:6.1f} {sm:6.1f} {occ:6.1f} {regs:5.0f}")
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += us; a[2] += rd; a[3] += wr
    print("\n# per kernel: launches, total us, total DRAM MB (read + write), aggregate GB/s")
    for name, (n, us, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:44s} {n:4d} {us:10.2f} {(rd + wr) / 1e6:10.3f} {(rd + wr) / (us * 1e-6) / 1e9 if us else 0:8.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
