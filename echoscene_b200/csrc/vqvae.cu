// VQ-VAE decode: VQVAE.decode_no_quant (vqvae_networks/network.py:95-103) -- the step right after the shape chain
// (EchoToShape.rel2shape, echo2shape.py:522).  latents (n, 3, 16, 16, 16) -> SDF (n, 1, 64, 64, 64):
//   VectorQuantizer (quantizer.py:68-99, nearest of 8192 codes per voxel, straight-through value z + (e - z))
//   -> post_quant_conv 1x1x1 -> Decoder3D (vqvae_modules.py:377-409): conv_in, ResnetBlock, AttnBlock (one head over all
//   4096 voxels), ResnetBlock, then per level ResnetBlock(s) + nearest x2 upsample + conv, GroupNorm, GELU, conv_out.
// Same conventions as shape.cu: channels-last activations, one workspace sized by a dry run, every kernel on the
// caller's stream.  Contractions reuse the implicit-GEMM kernels (tcgen05 in ECHO_PREC_BF16 where the shape allows, fp32
// FMA otherwise and always in ECHO_PREC_FP32, the parity mode); attention over 4096 tokens x 256 channels uses the fp32
// kernels in both modes (it is 2 % of the decoder's FLOPs).
#include "unet.cuh"

#include <math.h>
#include <stdlib.h>

using namespace echo;

namespace {

// one thread per voxel: squared distance to every code as the reference forms it, d = |z|^2 + |e|^2 - 2 z.e
// (quantizer.py:79-82), first minimum wins (torch.argmin), value z + (e - z) (:96), then post_quant_conv (network.py:101).
// Codebook and |e|^2 live in shared memory.  in: NCDHW (n, 3, vox); out: channels-last (n, vox, 3).
__global__ void __launch_bounds__(256) vq_quantize_kernel(const float* __restrict__ z, long long vox, long long total,
                                                          const float* __restrict__ codebook, int n_embed, const float* __restrict__ pw,
                                                          const float* __restrict__ pb, float* __restrict__ out, int* __restrict__ indices) {
  extern __shared__ float sm[];   // [n_embed][3] codes, [n_embed] squared norms
  float* ee = sm + 3 * n_embed;
  for (int i = threadIdx.x; i < n_embed; i += blockDim.x) {
    const float a = codebook[3 * i], b = codebook[3 * i + 1], c = codebook[3 * i + 2];
    sm[3 * i] = a; sm[3 * i + 1] = b; sm[3 * i + 2] = c;
    ee[i] = a * a + b * b + c * c;
  }
  __syncthreads();
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (long long)gridDim.x * blockDim.x) {
    const long long obj = v / vox, r = v - obj * vox;
    const float* zp = z + obj * 3 * vox + r;
    const float z0 = zp[0], z1 = zp[vox], z2 = zp[2 * vox];
    const float zz = z0 * z0 + z1 * z1 + z2 * z2;
    float best = INFINITY;
    int bi = 0;
    for (int i = 0; i < n_embed; ++i) {
      const float dot = fmaf(z2, sm[3 * i + 2], fmaf(z1, sm[3 * i + 1], z0 * sm[3 * i]));
      const float d = (zz + ee[i]) - 2.f * dot;
      if (d < best) { best = d; bi = i; }
    }
    const float q0 = z0 + (sm[3 * bi] - z0), q1 = z1 + (sm[3 * bi + 1] - z1), q2 = z2 + (sm[3 * bi + 2] - z2);
    float* o = out + v * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = fmaf(pw[3 * c + 2], q2, fmaf(pw[3 * c + 1], q1, pw[3 * c] * q0)) + pb[c];
    if (indices) indices[v] = bi;
  }
}

// nearest x2 in d, h, w of a channels-last tensor (F.interpolate(scale_factor=2, mode="nearest"), vqvae_modules.py:36)
template <class T>
__global__ void upsample_dhw2_kernel(const T* __restrict__ x, int n, int d, int h, int w, int cv, long long nvec, T* __restrict__ out) {
  constexpr int VE = 16 / (int)sizeof(T);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv); long long r = i / cv;
    const int ow = (int)(r % (2 * w)); r /= 2 * w;
    const int oh = (int)(r % (2 * h)); r /= 2 * h;
    const int od = (int)(r % (2 * d)); const long long obj = r / (2 * d);
    const long long src = (((obj * d + (od >> 1)) * h + (oh >> 1)) * w + (ow >> 1)) * (long long)cv + c;
    reinterpret_cast<uint4*>(out)[i] = __ldg(reinterpret_cast<const uint4*>(x) + src);
    (void)VE;
  }
}

template <class T>
__global__ void gelu_kernel(T* __restrict__ x, long long count) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const float v = (float)x[i];
    x[i] = (T)(0.5f * v * (1.f + erff(v * 0.70710678118654752440f)));   // nn.GELU(), exact erf form
  }
}

struct ResnetW {
  int cin = 0, cout = 0;
  NormW n1, n2;
  ConvW c1, c2, nin;
  bool has_nin = false;
};

// state_dict entries -> prepared device copies in the handle's pool (conv kernels repacked tap-major, optional bf16 copy)
struct VqLoader {
  const WeightMap& wm;
  DevPool& pool;
  bool bf;
  cudaStream_t s;
  const float* vec(const std::string& name, int n) const {
    const WView& v = wm.get(name, {n});
    float* o = pool.alloc_n<float>(n);
    ECHO_CUDA(cudaMemcpyAsync(o, v.p, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
    return o;
  }
  const __nv_bfloat16* to_bf16(const float* w, size_t n) const {
    __nv_bfloat16* o = pool.alloc_n<__nv_bfloat16>(n);
    convert(w, F32, o, BF16, (int64_t)n, s);
    return o;
  }
  ConvW conv(const std::string& p, int cin, int cout, int k) const {
    ConvW c;
    c.cin = cin; c.cout = cout; c.taps = k * k * k;
    const WView& v = wm.get(p + ".weight", {cout, cin, k, k, k});
    const size_t n_el = (size_t)cout * cin * c.taps;
    float* o = pool.alloc_n<float>(n_el);
    if (c.taps == 1) ECHO_CUDA(cudaMemcpyAsync(o, v.p, sizeof(float) * n_el, cudaMemcpyDeviceToDevice, s));
    else repack_conv_weight(v.p, cout, cin, c.taps, o, s);
    c.w = o;
    if (bf) c.wb = to_bf16(o, n_el);
    c.b = vec(p + ".bias", cout);
    return c;
  }
  NormW norm(const std::string& p, int c) const {
    NormW n;
    n.c = c;
    n.g = vec(p + ".weight", c);
    n.b = vec(p + ".bias", c);
    return n;
  }
  ResnetW resnet(const std::string& p, int cin, int cout) const {
    ResnetW r;
    r.cin = cin; r.cout = cout;
    ECHO_CHECK(cin % 32 == 0 && cout % 32 == 0, "vqvae: Normalize() with 32 groups needs channels %% 32 == 0 (%d, %d)", cin, cout);
    r.n1 = norm(p + ".norm1", cin);
    r.c1 = conv(p + ".conv1", cin, cout, 3);
    r.n2 = norm(p + ".norm2", cout);
    r.c2 = conv(p + ".conv2", cout, cout, 3);
    r.has_nin = cin != cout;
    if (r.has_nin) r.nin = conv(p + ".nin_shortcut", cin, cout, 1);
    return r;
  }
  // q, k, v 1x1 convs of an AttnBlock stacked [3C, C] (+ biases) for one projection launch (fp32)
  ConvW qkv_stack(const std::string& p, int C) const {
    float* w = pool.alloc_n<float>((size_t)3 * C * C);
    float* b = pool.alloc_n<float>((size_t)3 * C);
    const char* names[3] = {"q", "k", "v"};
    for (int i = 0; i < 3; ++i) {
      const WView& wv = wm.get(p + names[i] + ".weight", {C, C, 1, 1, 1});
      const WView& bv = wm.get(p + names[i] + ".bias", {C});
      ECHO_CUDA(cudaMemcpyAsync(w + (size_t)i * C * C, wv.p, sizeof(float) * C * C, cudaMemcpyDeviceToDevice, s));
      ECHO_CUDA(cudaMemcpyAsync(b + (size_t)i * C, bv.p, sizeof(float) * C, cudaMemcpyDeviceToDevice, s));
    }
    ConvW q;
    q.cin = C; q.cout = 3 * C; q.taps = 1; q.w = w; q.b = b;
    return q;
  }
};

}  // namespace

struct echo_vqvae {
  echo_vqvae_desc_t d;
  DevPool pool;
  Arena arena;
  bool dry = false;
  int prec = ECHO_PREC_FP32;
  DT adt = F32;
  const float *codebook = nullptr, *pq_w = nullptr, *pq_b = nullptr;
  ConvW conv_in, conv_out, qkv, attn_out;
  ConvW conv_out_pad;          // bf16 mode: conv_out zero-padded to 32 output channels so it runs on the tcgen05 kernel
  ConvW q_s, k_w, v_w;         // bf16 mode: q (pre-scaled by C^-0.5), k, v projections as separate dense matrices
  NormW attn_norm, norm_out;
  ResnetW mid1, mid2;
  std::vector<std::vector<ResnetW>> up_blocks;   // [level][block]
  std::vector<ConvW> up_conv;                    // [level] (level 0: unused)
  std::vector<ConvW> up_fold;                    // bf16 mode: the same convs folded per output phase of the x2 upsample
  // encoder handles (vqvae_encoder_create): Encoder3D + quant_conv; conv_in / mid1 / attn_* / mid2 / norm_out / conv_out above
  // hold the ENCODER's tensors then
  bool is_encoder = false;
  int in_ch = 1;
  std::vector<std::vector<ResnetW>> down_blocks;   // [level][block]
  std::vector<ConvW> down_conv;                    // [level] (last level: unused)
  ConvW quant_conv;

  Act new_act(int n, int dd, int h, int w, int c, DT dt) {
    Act a;
    a.n = n; a.d = dd; a.h = h; a.w = w; a.c = c; a.dt = dt;
    a.p = arena.alloc(a.bytes());
    return a;
  }
  // stride / pad: Conv3d(k, stride, padding = pad) with pad < 0 meaning k / 2; taps that fall outside the input read zeros, so
  // Downsample's explicit (0, 1) zero pad + Conv3d(k3, stride 2, padding 0) (vqvae_modules.py:42-58) is stride 2, pad 0 with
  // an output grid of half the input
  void contract(const Act& x, const ConvW& w, int k, const Act* res, const Act& out, cudaStream_t s, int stride = 1, int pad = -1) {
    if (dry) return;
    GemmArgs g;
    g.A = x.p; g.a_dt = x.dt; g.n = x.n; g.d = x.d; g.h = x.h; g.w = x.w; g.cin = x.c; g.lda = x.c;
    g.od = out.d; g.oh = out.h; g.ow = out.w;
    g.kd = g.kh = g.kw = k; g.pd = g.ph = g.pw = pad < 0 ? k / 2 : pad;
    g.sd = g.sh = g.sw = stride;
    g.W = w.w; g.w_dt = F32; g.w_stride_n = (int64_t)w.taps * w.cin; g.cout = w.cout; g.bias = w.b;
    if (res) { g.res = res->p; g.res_dt = res->dt; g.ld_res = res->c; }
    g.out = out.p; g.out_dt = out.dt; g.ldo = out.c;
    ECHO_CHECK(w.cin == x.c && w.cout == out.c && w.taps == k * k * k, "vqvae: weight/activation mismatch (cin %d vs %d, cout %d vs %d)", w.cin,
               x.c, w.cout, out.c);
    if (prec == ECHO_PREC_BF16 && w.wb && x.dt == BF16 && stride == 1 && tc_available()) {
      GemmArgs t = g;
      t.W = w.wb; t.w_dt = BF16;
      if (gemm_tc_supported(t)) { gemm_tc(t, s); return; }
    }
    gemm_simt(g, s);
  }
  Act gn(const Act& x, const NormW& nw, bool swish, cudaStream_t s) {
    float* stats = arena.alloc_n<float>((size_t)x.n * 32 * 2);
    float* partial = arena.alloc_n<float>(gn_partial_floats(x, 32));
    Act o = new_act(x.n, x.d, x.h, x.w, x.c, x.dt);
    if (!dry) {
      gn_stats(x, 32, 1e-6f, stats, partial, s);            // Normalize(): GroupNorm(32, eps 1e-6), vqvae_modules.py:13-22
      gn_apply(x, stats, nw.g, nw.b, 32, swish, o, s);
    }
    return o;
  }
  // ResnetBlock.forward, temb = None (vqvae_modules.py:107-127)
  Act resnet(const Act& x, const ResnetW& r, cudaStream_t s) {
    Act out = new_act(x.n, x.d, x.h, x.w, r.cout, x.dt);
    const size_t m = arena.mark();
    Act a1 = gn(x, r.n1, true, s);
    Act h1 = new_act(x.n, x.d, x.h, x.w, r.cout, x.dt);
    contract(a1, r.c1, 3, nullptr, h1, s);
    Act a2 = gn(h1, r.n2, true, s);
    if (r.has_nin) {
      Act sk = new_act(x.n, x.d, x.h, x.w, r.cout, x.dt);
      contract(x, r.nin, 1, nullptr, sk, s);
      contract(a2, r.c2, 3, &sk, out, s);
    } else {
      contract(a2, r.c2, 3, &x, out, s);
    }
    arena.release(m);
    return out;
  }
  // dense [rows, cin] x [cout, cin]^T on the tcgen05 kernel (rows presented as a 1 x 1 x rows "voxel" line)
  void tc_matmul(const void* A, int64_t rows, int cin, const void* W, const float* bias, int cout, void* out, DT out_dt, int64_t ldo,
                 int out_t, cudaStream_t s) {
    GemmArgs g;
    g.A = A; g.a_dt = BF16; g.n = 1; g.d = 1; g.h = 1; g.w = (int)rows; g.cin = cin; g.lda = cin;
    g.od = 1; g.oh = 1; g.ow = (int)rows;
    g.W = W; g.w_dt = BF16; g.w_stride_n = cin; g.cout = cout; g.bias = bias;
    g.out = out; g.out_dt = out_dt; g.ldo = ldo; g.out_t = out_t;
    ECHO_CHECK(gemm_tc_supported(g), "vqvae: %lld x %d x %d matmul not supported by the tcgen05 kernel", (long long)rows, cin, cout);
    gemm_tc(g, s);
  }
  // AttnBlock on the tensor cores (bf16 mode): per object S = (Q C^-0.5) K^T (fp32, materialised: 64 MB), row softmax in
  // fp32, P rounded to bf16, O = P V with V projected straight into V^T.  One head of 256 channels over 4096 tokens is
  // GEMM-shaped work (8.6 GFLOP per product and object), not flash-attention-shaped: the tcgen05 GEMM takes it as is.
  Act attn_tc(const Act& x, cudaStream_t s) {
    Act out = new_act(x.n, x.d, x.h, x.w, x.c, x.dt);
    const size_t m = arena.mark();
    const int C = x.c, tokens = (int)x.voxels();
    const int64_t rows = x.rows();
    Act xn = gn(x, attn_norm, false, s);
    __nv_bfloat16* q = arena.alloc_n<__nv_bfloat16>((size_t)rows * C);
    __nv_bfloat16* k = arena.alloc_n<__nv_bfloat16>((size_t)rows * C);
    __nv_bfloat16* vt = arena.alloc_n<__nv_bfloat16>((size_t)rows * C);          // [(obj, c), tokens]
    float* sc = arena.alloc_n<float>((size_t)tokens * tokens);                   // one object's scores at a time
    __nv_bfloat16* pb = arena.alloc_n<__nv_bfloat16>((size_t)tokens * tokens);
    Act o = new_act(x.n, x.d, x.h, x.w, C, BF16);
    if (!dry) {
      GemmArgs g;   // projections over all objects at once (per-object transposed store for V)
      g.A = xn.p; g.a_dt = BF16; g.n = x.n; g.d = x.d; g.h = x.h; g.w = x.w; g.cin = C; g.lda = C;
      g.od = x.d; g.oh = x.h; g.ow = x.w; g.w_dt = BF16; g.w_stride_n = C; g.cout = C; g.out_dt = BF16; g.ldo = C;
      GemmArgs gq = g; gq.W = q_s.wb; gq.bias = q_s.b; gq.out = q;
      GemmArgs gk = g; gk.W = k_w.wb; gk.bias = k_w.b; gk.out = k;
      GemmArgs gv = g; gv.W = v_w.wb; gv.bias = v_w.b; gv.out = vt; gv.out_t = 1;
      ECHO_CHECK(gemm_tc_supported(gq) && gemm_tc_supported(gv), "vqvae: attention projections not supported by the tcgen05 kernel");
      gemm_tc(gq, s);
      gemm_tc(gk, s);
      gemm_tc(gv, s);
      for (int ob = 0; ob < x.n; ++ob) {
        const __nv_bfloat16* qo = q + (size_t)ob * tokens * C;
        const __nv_bfloat16* ko = k + (size_t)ob * tokens * C;
        tc_matmul(qo, tokens, C, ko, nullptr, tokens, sc, F32, tokens, 0, s);                       // S = Q K^T (scale folded into Q)
        softmax_rows(sc, tokens, tokens, s);                                                        // vqvae_modules.py:178
        convert(sc, F32, pb, BF16, (int64_t)tokens * tokens, s);
        tc_matmul(pb, tokens, tokens, vt + (size_t)ob * C * tokens, nullptr, C, (__nv_bfloat16*)o.p + (size_t)ob * tokens * C, BF16, C, 0, s);
      }
    }
    contract(o, attn_out, 1, &x, out, s);
    arena.release(m);
    return out;
  }

  // AttnBlock.forward (vqvae_modules.py:158-195): one head over all voxels, scale C^-0.5
  Act attn(const Act& x, cudaStream_t s) {
    static const bool no_tc_attn = getenv("ECHO_VQ_NO_TC_ATTN") != nullptr;
    if (prec == ECHO_PREC_BF16 && x.dt == BF16 && q_s.wb && !no_tc_attn && x.voxels() % 128 == 0) return attn_tc(x, s);
    Act out = new_act(x.n, x.d, x.h, x.w, x.c, x.dt);
    const size_t m = arena.mark();
    const int C = x.c, tokens = (int)x.voxels();
    Act xn = gn(x, attn_norm, false, s);
    Act qkvt = new_act(x.n, x.d, x.h, x.w, 3 * C, F32);
    contract(xn, qkv, 1, nullptr, qkvt, s);
    float* ws = arena.alloc_n<float>(attention_f32_ws_floats(x.n, tokens, 1));
    Act o32 = new_act(x.n, x.d, x.h, x.w, C, F32);
    if (!dry) attention_f32((const float*)qkvt.p, x.n, tokens, 1, C, ws, (float*)o32.p, s);
    Act o = o32;
    if (x.dt != F32) {
      o = new_act(x.n, x.d, x.h, x.w, C, x.dt);
      if (!dry) convert(o32.p, F32, o.p, o.dt, o.rows() * C, s);
    }
    contract(o, attn_out, 1, &x, out, s);
    arena.release(m);
    return out;
  }
  Act upsample(const Act& x, cudaStream_t s) {
    Act o = new_act(x.n, 2 * x.d, 2 * x.h, 2 * x.w, x.c, x.dt);
    if (!dry) {
      const int cv = (int)(x.c * dt_size(x.dt) / 16);
      ECHO_CHECK(x.c * dt_size(x.dt) % 16 == 0, "vqvae: upsample needs 16-byte channel vectors");
      const long long nvec = (long long)o.rows() * cv;
      long long blocks = (nvec + 255) / 256;
      if (blocks > 148 * 32) blocks = 148 * 32;
      if (x.dt == F32) upsample_dhw2_kernel<float><<<(int)blocks, 256, 0, s>>>((const float*)x.p, x.n, x.d, x.h, x.w, cv, nvec, (float*)o.p);
      else upsample_dhw2_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, s>>>((const __nv_bfloat16*)x.p, x.n, x.d, x.h, x.w, cv, nvec, (__nv_bfloat16*)o.p);
      ECHO_LAUNCH_CHECK();
    }
    return o;
  }

  void gelu_inplace(const Act& x, cudaStream_t s) {
    if (dry) return;
    const long long cnt = (long long)x.rows() * x.c;
    long long blocks = (cnt + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (x.dt == F32) gelu_kernel<float><<<(int)blocks, 256, 0, s>>>((float*)x.p, cnt);
    else gelu_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, s>>>((__nv_bfloat16*)x.p, cnt);
    ECHO_LAUNCH_CHECK();
  }

  // VQVAE.encode_no_quant (network.py:84-88): Encoder3D.forward (vqvae_modules.py:256-289) -> quant_conv.
  // sdf (n, in_ch, R, R, R) NCDHW with in_ch == 1 (so it is channels-last as it stands) -> latents (n, embed_dim, L, L, L) NCDHW
  void run_encode(const float* sdf, int n, float* latents_out, cudaStream_t s) {
    ECHO_CHECK(is_encoder, "vqvae_encode: this handle was created by echo_vqvae_create (decoder); use echo_vqvae_encoder_create");
    ECHO_CHECK(n >= 0 && n <= d.max_objects, "vqvae_encode: %d objects exceed the handle's capacity %d", n, d.max_objects);
    if (n == 0) return;
    arena.release(0);
    const int R = d.latent_size << (d.num_levels - 1);
    Act x;
    x.n = n; x.d = x.h = x.w = R; x.c = in_ch; x.dt = F32; x.p = const_cast<float*>(sdf);
    Act h = new_act(n, R, R, R, conv_in.cout, adt);
    contract(x, conv_in, 3, nullptr, h, s);
    for (int lvl = 0; lvl < d.num_levels; ++lvl) {
      for (auto& r : down_blocks[lvl]) h = resnet(h, r, s);
      if (lvl != d.num_levels - 1) {
        Act o = new_act(n, h.d / 2, h.h / 2, h.w / 2, down_conv[lvl].cout, adt);
        contract(h, down_conv[lvl], 3, nullptr, o, s, 2, 0);
        h = o;
      }
    }
    h = resnet(h, mid1, s);
    h = attn(h, s);
    h = resnet(h, mid2, s);
    Act hn = gn(h, norm_out, false, s);
    gelu_inplace(hn, s);                                            // activ = 'gelu', vqvae_modules.py:199-201, 288
    Act z = new_act(n, hn.d, hn.h, hn.w, conv_out.cout, F32);
    contract(hn, conv_out, 3, nullptr, z, s);
    Act q = new_act(n, hn.d, hn.h, hn.w, quant_conv.cout, F32);
    contract(z, quant_conv, 1, nullptr, q, s);
    if (!dry) cl_to_ncdhw(q.p, F32, n, q.c, q.voxels(), q.c, latents_out, s);
  }

  void run(const float* latents, int n, float* sdf_out, int* indices_out, cudaStream_t s) {
    ECHO_CHECK(!is_encoder, "vqvae_decode: this handle was created by echo_vqvae_encoder_create (encoder)");
    ECHO_CHECK(n >= 0 && n <= d.max_objects, "vqvae_decode: %d objects exceed the handle's capacity %d", n, d.max_objects);
    if (n == 0) return;
    arena.release(0);
    const int L = d.latent_size;
    const long long vox = (long long)L * L * L;
    Act zq = new_act(n, L, L, L, d.z_channels, F32);
    if (!dry) {
      const size_t smem = (size_t)d.n_embed * 4 * sizeof(float);
      vq_quantize_kernel<<<148 * 2, 256, smem, s>>>(latents, vox, (long long)n * vox, codebook, d.n_embed, pq_w, pq_b, (float*)zq.p, indices_out);
      ECHO_LAUNCH_CHECK();
    }
    Act h = new_act(n, L, L, L, conv_in.cout, adt);
    contract(zq, conv_in, 3, nullptr, h, s);                 // 3 input channels: fp32 FMA kernel in both modes
    h = resnet(h, mid1, s);
    h = attn(h, s);
    h = resnet(h, mid2, s);
    for (int lvl = d.num_levels - 1; lvl >= 0; --lvl) {
      for (auto& r : up_blocks[lvl]) h = resnet(h, r, s);
      if (lvl != 0) {
        static const bool no_fold = getenv("ECHO_VQ_NO_UPFOLD") != nullptr;   // A/B: explicit upsample + 27-tap conv
        if (prec == ECHO_PREC_BF16 && h.dt == BF16 && up_fold[lvl].wb && !no_fold) {
          // nearest x2 folded into the conv: eight output phases, 2x2x2 taps each, on the low-resolution tensor
          Act o = new_act(h.n, 2 * h.d, 2 * h.h, 2 * h.w, up_conv[lvl].cout, adt);
          if (!dry) {
            GemmArgs g;
            g.A = h.p; g.a_dt = BF16; g.n = h.n; g.d = h.d; g.h = h.h; g.w = h.w; g.cin = h.c; g.lda = h.c;
            g.od = o.d; g.oh = o.h; g.ow = o.w; g.up2 = 2;
            g.kd = g.kh = g.kw = 3; g.pd = g.ph = g.pw = 1;
            g.W = up_fold[lvl].wb; g.w_dt = BF16; g.w_stride_n = (int64_t)64 * h.c; g.cout = o.c; g.bias = up_fold[lvl].b;
            g.out = o.p; g.out_dt = BF16; g.ldo = o.c;
            ECHO_CHECK(gemm_tc_supported(g), "vqvae: folded upsample conv not supported by the tcgen05 kernel");
            gemm_tc(g, s);
          }
          h = o;
        } else {
          Act u = upsample(h, s);
          Act o = new_act(u.n, u.d, u.h, u.w, up_conv[lvl].cout, adt);
          contract(u, up_conv[lvl], 3, nullptr, o, s);
          h = o;
        }
      }
    }
    Act hn = gn(h, norm_out, false, s);
    if (!dry) {
      const long long cnt = (long long)hn.rows() * hn.c;
      long long blocks = (cnt + 255) / 256;
      if (blocks > 148 * 32) blocks = 148 * 32;
      if (hn.dt == F32) gelu_kernel<float><<<(int)blocks, 256, 0, s>>>((float*)hn.p, cnt);
      else gelu_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, s>>>((__nv_bfloat16*)hn.p, cnt);
      ECHO_LAUNCH_CHECK();
    }
    // conv_out: out_ch = 1, so channels-last == the reference's NCDHW
    static const bool no_pad = getenv("ECHO_VQ_NO_OUTPAD") != nullptr;
    if (prec == ECHO_PREC_BF16 && hn.dt == BF16 && conv_out_pad.wb && !no_pad) {
      // a 1-column GEMM wastes the tensor core far less than a scalar reduction wastes the SM: 32 padded output channels
      Act e32 = new_act(n, hn.d, hn.h, hn.w, 32, F32);
      contract(hn, conv_out_pad, 3, nullptr, e32, s);
      if (!dry) copy_cols((const float*)e32.p, 32, e32.rows(), 1, sdf_out, 1, s);
    } else {
      Act e;
      e.n = n; e.d = hn.d; e.h = hn.h; e.w = hn.w; e.c = d.out_ch; e.dt = F32; e.p = sdf_out;
      contract(hn, conv_out, 3, nullptr, e, s);
    }
  }
};

namespace echo {

echo_vqvae* vqvae_create(const echo_vqvae_desc_t* desc, const echo_weight_t* weights, int n_weights) {
  ECHO_CHECK(desc, "vqvae: null desc");
  echo_vqvae* h = new echo_vqvae();
  try {
    h->d = *desc;
    const echo_vqvae_desc_t& d = h->d;
    ECHO_CHECK(d.embed_dim == 3 && d.z_channels == 3, "vqvae: embed_dim / z_channels must be 3 (config/vqvae_snet.yaml)");
    ECHO_CHECK(d.out_ch == 1, "vqvae: out_ch must be 1 (the output is written as NCDHW == channels-last)");
    ECHO_CHECK(d.num_levels >= 1 && d.num_levels <= 8 && d.max_objects > 0 && d.n_embed > 0 && d.latent_size > 0, "vqvae: bad config");
    ECHO_CHECK((size_t)d.n_embed * 16 <= 200 * 1024, "vqvae: codebook does not fit shared memory");
    h->prec = d.precision;
    ECHO_CHECK(h->prec == ECHO_PREC_FP32 || h->prec == ECHO_PREC_BF16, "vqvae: unknown precision %d", h->prec);
    if (h->prec == ECHO_PREC_BF16 && !tc_available())
      fail(ECHO_ERR_UNSUPPORTED, "vqvae: ECHO_PREC_BF16 needs the sm_100a tcgen05 kernels on a B200-class device");
    h->adt = h->prec == ECHO_PREC_BF16 ? BF16 : F32;
    WeightMap wm;
    wm.load(weights, n_weights);
    cudaStream_t s = 0;
    DevPool& pool = h->pool;
    const bool bf = h->prec == ECHO_PREC_BF16;
    const VqLoader ld{wm, pool, bf, s};
    auto vec = [&](const std::string& name, int n) { return ld.vec(name, n); };
    auto to_bf16 = [&](const float* w, size_t n) { return ld.to_bf16(w, n); };
    auto conv = [&](const std::string& p, int cin, int cout, int k) { return ld.conv(p, cin, cout, k); };
    auto norm = [&](const std::string& p, int c) { return ld.norm(p, c); };
    auto resnet = [&](const std::string& p, int cin, int cout) { return ld.resnet(p, cin, cout); };
    {
      const WView& cb = wm.get("quantize.embedding.weight", {d.n_embed, d.embed_dim});
      float* o = pool.alloc_n<float>(cb.numel());
      ECHO_CUDA(cudaMemcpyAsync(o, cb.p, sizeof(float) * cb.numel(), cudaMemcpyDeviceToDevice, s));
      h->codebook = o;
      const WView& pw = wm.get("post_quant_conv.weight", {d.z_channels, d.embed_dim, 1, 1, 1});
      float* w = pool.alloc_n<float>(9);
      ECHO_CUDA(cudaMemcpyAsync(w, pw.p, sizeof(float) * 9, cudaMemcpyDeviceToDevice, s));
      h->pq_w = w;
      h->pq_b = vec("post_quant_conv.bias", d.z_channels);
    }
    int block_in = d.ch * d.ch_mult[d.num_levels - 1];
    h->conv_in = conv("decoder.conv_in", d.z_channels, block_in, 3);
    h->conv_in.wb = nullptr;                                   // 3 input channels: fp32 kernel
    h->mid1 = resnet("decoder.mid.block_1", block_in, block_in);
    h->attn_norm = norm("decoder.mid.attn_1.norm", block_in);
    {   // q, k, v 1x1 convs stacked [3C, C] (+ biases) for one projection launch
      const int C = block_in;
      float* w = pool.alloc_n<float>((size_t)3 * C * C);
      float* b = pool.alloc_n<float>((size_t)3 * C);
      const char* names[3] = {"q", "k", "v"};
      for (int i = 0; i < 3; ++i) {
        const WView& wv = wm.get(std::string("decoder.mid.attn_1.") + names[i] + ".weight", {C, C, 1, 1, 1});
        const WView& bv = wm.get(std::string("decoder.mid.attn_1.") + names[i] + ".bias", {C});
        ECHO_CUDA(cudaMemcpyAsync(w + (size_t)i * C * C, wv.p, sizeof(float) * C * C, cudaMemcpyDeviceToDevice, s));
        ECHO_CUDA(cudaMemcpyAsync(b + (size_t)i * C, bv.p, sizeof(float) * C, cudaMemcpyDeviceToDevice, s));
      }
      h->qkv.cin = C; h->qkv.cout = 3 * C; h->qkv.taps = 1; h->qkv.w = w; h->qkv.b = b;   // fp32 output for the fp32 attention
      if (bf) {   // tensor-core attention: separate dense projections, the softmax scale C^-0.5 folded into q (weights and bias)
        h->q_s = conv("decoder.mid.attn_1.q", C, C, 1);
        h->k_w = conv("decoder.mid.attn_1.k", C, C, 1);
        h->v_w = conv("decoder.mid.attn_1.v", C, C, 1);
        const float sc = 1.0f / sqrtf((float)C);
        std::vector<float> hw((size_t)C * C), hb(C);
        ECHO_CUDA(cudaStreamSynchronize(s));
        ECHO_CUDA(cudaMemcpy(hw.data(), h->q_s.w, hw.size() * sizeof(float), cudaMemcpyDeviceToHost));
        ECHO_CUDA(cudaMemcpy(hb.data(), h->q_s.b, hb.size() * sizeof(float), cudaMemcpyDeviceToHost));
        for (auto& v : hw) v *= sc;
        for (auto& v : hb) v *= sc;
        float* ws = pool.alloc_n<float>(hw.size());
        float* bs = pool.alloc_n<float>(hb.size());
        ECHO_CUDA(cudaMemcpy(ws, hw.data(), hw.size() * sizeof(float), cudaMemcpyHostToDevice));
        ECHO_CUDA(cudaMemcpy(bs, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice));
        h->q_s.w = ws; h->q_s.b = bs; h->q_s.wb = to_bf16(ws, hw.size());
      }
    }
    h->attn_out = conv("decoder.mid.attn_1.proj_out", block_in, block_in, 1);
    h->mid2 = resnet("decoder.mid.block_2", block_in, block_in);
    h->up_blocks.resize(d.num_levels);
    h->up_conv.resize(d.num_levels);
    h->up_fold.resize(d.num_levels);
    for (int lvl = d.num_levels - 1; lvl >= 0; --lvl) {
      const int block_out = d.ch * d.ch_mult[lvl];
      for (int i = 0; i < d.num_res_blocks; ++i) {
        h->up_blocks[lvl].push_back(resnet("decoder.up." + std::to_string(lvl) + ".block." + std::to_string(i), block_in, block_out));
        block_in = block_out;
      }
      if (lvl != 0) {
        h->up_conv[lvl] = conv("decoder.up." + std::to_string(lvl) + ".upsample.conv", block_in, block_in, 3);
        if (bf) {
          const size_t n_src = (size_t)block_in * 27 * block_in, n_dst = (size_t)block_in * 64 * block_in;
          std::vector<float> hsrc(n_src), hdst(n_dst);
          ECHO_CUDA(cudaStreamSynchronize(s));
          ECHO_CUDA(cudaMemcpy(hsrc.data(), h->up_conv[lvl].w, n_src * sizeof(float), cudaMemcpyDeviceToHost));
          fold_upsample_weight(hsrc.data(), block_in, block_in, hdst.data(), true);
          float* o = pool.alloc_n<float>(n_dst);
          ECHO_CUDA(cudaMemcpy(o, hdst.data(), n_dst * sizeof(float), cudaMemcpyHostToDevice));
          h->up_fold[lvl] = h->up_conv[lvl];
          h->up_fold[lvl].taps = 64;
          h->up_fold[lvl].w = o;
          h->up_fold[lvl].wb = to_bf16(o, n_dst);
        }
      }
    }
    h->norm_out = norm("decoder.norm_out", block_in);
    h->conv_out = conv("decoder.conv_out", block_in, d.out_ch, 3);
    h->conv_out.wb = nullptr;                                  // one output channel: fp32 kernel
    if (bf) {   // ... or 32 zero-padded output channels on the tensor cores
      const size_t row = (size_t)27 * block_in;
      float* o = pool.alloc_n<float>(32 * row);
      float* b = pool.alloc_n<float>(32);
      ECHO_CUDA(cudaMemsetAsync(o, 0, 32 * row * sizeof(float), s));
      ECHO_CUDA(cudaMemsetAsync(b, 0, 32 * sizeof(float), s));
      ECHO_CUDA(cudaMemcpyAsync(o, h->conv_out.w, d.out_ch * row * sizeof(float), cudaMemcpyDeviceToDevice, s));
      ECHO_CUDA(cudaMemcpyAsync(b, h->conv_out.b, d.out_ch * sizeof(float), cudaMemcpyDeviceToDevice, s));
      h->conv_out_pad = h->conv_out;
      h->conv_out_pad.cout = 32;
      h->conv_out_pad.w = o;
      h->conv_out_pad.b = b;
      h->conv_out_pad.wb = to_bf16(o, 32 * row);
    }
    ECHO_CUDA(cudaFuncSetAttribute(vq_quantize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)d.n_embed * 16)));
    ECHO_CUDA(cudaStreamSynchronize(s));
    // size the workspace with a dry run at full capacity
    h->dry = true;
    h->arena.base = nullptr;
    h->arena.cap = ~size_t(0) >> 1;
    h->arena.off = h->arena.high = 0;
    h->run(nullptr, d.max_objects, nullptr, nullptr, s);
    h->dry = false;
    h->arena.init(h->arena.high + (size_t(1) << 20));
    return h;
  } catch (...) {
    h->arena.destroy();
    h->pool.destroy();
    delete h;
    throw;
  }
}

// Encoder3D + quant_conv.  fp32 only: the strided convs and the 1-channel stem have no tensor-core route yet,
// and the training path that calls encode_no_quant (echo2shape.py:334-364) wants the reference's fp32 latents.
echo_vqvae* vqvae_encoder_create(const echo_vqvae_desc_t* desc, const echo_weight_t* weights, int n_weights) {
  ECHO_CHECK(desc, "vqvae_encoder: null desc");
  echo_vqvae* h = new echo_vqvae();
  try {
    h->d = *desc;
    h->is_encoder = true;
    const echo_vqvae_desc_t& d = h->d;
    ECHO_CHECK(d.num_levels >= 1 && d.num_levels <= 8 && d.max_objects > 0 && d.latent_size > 0 && d.ch > 0 && d.num_res_blocks >= 0 &&
                   d.z_channels > 0 && d.embed_dim > 0,
               "vqvae_encoder: bad config");
    ECHO_CHECK(d.out_ch == 1, "vqvae_encoder: in_channels (== out_ch) must be 1 (the SDF volume is read as NCDHW == channels-last)");
    if (d.precision != ECHO_PREC_FP32) fail(ECHO_ERR_UNSUPPORTED, "vqvae_encoder: only ECHO_PREC_FP32 is implemented for encode_no_quant");
    h->prec = ECHO_PREC_FP32;
    h->adt = F32;
    h->in_ch = d.out_ch;
    WeightMap wm;
    wm.load(weights, n_weights);
    cudaStream_t s = 0;
    const VqLoader ld{wm, h->pool, false, s};
    h->conv_in = ld.conv("encoder.conv_in", h->in_ch, d.ch, 3);
    h->down_blocks.resize(d.num_levels);
    h->down_conv.resize(d.num_levels);
    int block_in = d.ch;
    for (int lvl = 0; lvl < d.num_levels; ++lvl) {
      block_in = d.ch * (lvl == 0 ? 1 : d.ch_mult[lvl - 1]);          // in_ch_mult = (1,) + ch_mult, vqvae_modules.py:213-218
      const int block_out = d.ch * d.ch_mult[lvl];
      for (int i = 0; i < d.num_res_blocks; ++i) {
        h->down_blocks[lvl].push_back(ld.resnet("encoder.down." + std::to_string(lvl) + ".block." + std::to_string(i), block_in, block_out));
        block_in = block_out;
      }
      if (lvl != d.num_levels - 1) h->down_conv[lvl] = ld.conv("encoder.down." + std::to_string(lvl) + ".downsample.conv", block_in, block_in, 3);
    }
    h->mid1 = ld.resnet("encoder.mid.block_1", block_in, block_in);
    h->attn_norm = ld.norm("encoder.mid.attn_1.norm", block_in);
    h->qkv = ld.qkv_stack("encoder.mid.attn_1.", block_in);
    h->attn_out = ld.conv("encoder.mid.attn_1.proj_out", block_in, block_in, 1);
    h->mid2 = ld.resnet("encoder.mid.block_2", block_in, block_in);
    h->norm_out = ld.norm("encoder.norm_out", block_in);
    h->conv_out = ld.conv("encoder.conv_out", block_in, d.z_channels, 3);   // double_z: False (config/vqvae_snet.yaml)
    h->quant_conv = ld.conv("quant_conv", d.z_channels, d.embed_dim, 1);
    ECHO_CUDA(cudaStreamSynchronize(s));
    h->dry = true;
    h->arena.base = nullptr;
    h->arena.cap = ~size_t(0) >> 1;
    h->arena.off = h->arena.high = 0;
    h->run_encode(nullptr, d.max_objects, nullptr, s);
    h->dry = false;
    h->arena.init(h->arena.high + (size_t(1) << 20));
    return h;
  } catch (...) {
    h->arena.destroy();
    h->pool.destroy();
    delete h;
    throw;
  }
}

void vqvae_encode(echo_vqvae* h, const float* sdf, int n, float* latents_out, cudaStream_t s) { h->run_encode(sdf, n, latents_out, s); }

void vqvae_destroy(echo_vqvae* h) {
  if (!h) return;
  h->arena.destroy();
  h->pool.destroy();
  delete h;
}

void vqvae_decode(echo_vqvae* h, const float* latents, int n, float* sdf_out, int* indices_out, cudaStream_t s) {
  h->run(latents, n, sdf_out, indices_out, s);
}

}  // namespace echo
