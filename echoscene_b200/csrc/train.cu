// Training-side pieces of SURVEY 8f-3 that do not need a backward pass of the denoisers:
//
//   * q_sample           x_t = sqrt(ac[t_i]) x0 + sqrt(1 - ac[t_i]) noise, one timestep per row
//                        (GaussianDiffusion.q_sample diffusion_ddpm.py:191-201, EchoToShape.q_sample echo2shape.py:254-258)
//   * the diffusion losses given the denoiser output: per-row mean squared error and its column-range parts
//                        (diffusion_loss diffusion_ddpm.py:451-477; get_loss(.., mean=False).mean([1,2,3,4]) echo2shape.py:297-318)
//   * the optimizer step of scripts/train_3dfront.py:247-259 as ONE pass over the parameters: clip_grad_norm_ of the shape
//     denoiser's gradients (:251), the per-parameter "isnan(grad).any() -> grad[isnan] = 0" loop (:252-256, ~700 host
//     synchronisations in the reference) and optimizerFULL.step() (AdamW, :258), multi-tensor, no host synchronisation.
//
// All of it is HBM-bound streaming: the step reads p, g, m, v and writes p, m, v (28 bytes per parameter) once.
#include "model.cuh"

#include <math.h>

#include <vector>

namespace echo {
struct OptTensor {
  float* p;
  float* g;
  float* m;
  float* v;
  int64_t n;
  int clip;   // member of the clip group
  int pad;
};
struct OptChunk {
  int tensor;
  int pad;
  int64_t begin;
};

namespace {

inline int grid_for(int64_t work_items, int threads) {   // enough blocks for the work, at most a few waves of the chip
  int64_t b = (work_items + threads - 1) / threads;
  const int64_t cap = 148 * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

__global__ void q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise, const int64_t* __restrict__ t,
                                const float* __restrict__ sqrt_ac, const float* __restrict__ sqrt_1mac, int64_t rows, int64_t row_len,
                                float* __restrict__ out) {
  griddep_launch();
  griddep_wait();
  const int64_t total = rows * row_len;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / row_len;
    const int64_t ti = t[r];
    // a * x0 + b * noise, two roundings of the products and one of the sum, as torch evaluates it
    out[i] = __fadd_rn(__fmul_rn(__ldg(sqrt_ac + ti), x0[i]), __fmul_rn(__ldg(sqrt_1mac + ti), noise[i]));
  }
}

// out[r][k] = mean over columns [c0_k, c1_k) of (target - pred)^2, k < n_ranges; one warp per row, fixed summation order
__global__ void mse_rows_kernel(const float* __restrict__ pred, const float* __restrict__ target, int64_t rows, int64_t row_len,
                                const int* __restrict__ ranges, int n_ranges, float* __restrict__ out) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* p = pred + r * row_len;
  const float* q = target + r * row_len;
  for (int k = 0; k < n_ranges; ++k) {
    const int c0 = ranges[2 * k], c1 = ranges[2 * k + 1];
    float s = 0.f;
    for (int c = c0 + lane; c < c1; c += 32) {
      const float d = q[c] - p[c];
      s = fmaf(d, d, s);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[r * n_ranges + k] = s / (float)(c1 - c0);
  }
}

// ---- the fused optimizer step -------------------------------------------------------------------------------------------
// One chunk = up to OPT_CHUNK consecutive elements of one tensor.  Pass 1 (clip group only): per-chunk sum of squares of the
// gradient (NaNs propagate, as in torch.nn.utils.clip_grad_norm_).  Pass 2: every chunk -- scale by the clip coefficient when
// the tensor belongs to the clip group, NaN -> 0, AdamW.
constexpr int OPT_CHUNK = 16384;

__global__ void opt_sumsq_kernel(const OptTensor* __restrict__ tensors, const OptChunk* __restrict__ chunks, int n_chunks,
                                 double* __restrict__ partial) {
  const int c = blockIdx.x;
  if (c >= n_chunks) return;
  const OptChunk ch = chunks[c];
  const OptTensor t = tensors[ch.tensor];
  double s = 0.0;
  if (t.clip) {
    const int64_t end = min(ch.begin + (int64_t)OPT_CHUNK, t.n);
    int64_t i0 = ch.begin;
    if (((uintptr_t)t.g & 15) == 0) {   // 16-byte vectors over the aligned body (chunk starts are multiples of OPT_CHUNK)
      const int64_t nv = (end - ch.begin) >> 2;
      const float4* g4 = reinterpret_cast<const float4*>(t.g + ch.begin);
      for (int64_t q = threadIdx.x; q < nv; q += blockDim.x) {
        const float4 g = g4[q];
        s += (double)g.x * g.x + (double)g.y * g.y + (double)g.z * g.z + (double)g.w * g.w;
      }
      i0 = ch.begin + (nv << 2);
    }
    for (int64_t i = i0 + threadIdx.x; i < end; i += blockDim.x) {
      const double g = (double)t.g[i];
      s += g * g;
    }
  }
  __shared__ double red[8];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += red[w];
    partial[c] = a;
  }
}

// total norm (fixed order) -> the clip coefficient of clip_grad_norm_: min(1, max_norm / (norm + 1e-6)); NaN when the norm is NaN
__global__ void opt_clip_coef_kernel(const double* __restrict__ partial, int n_chunks, float max_norm, float* __restrict__ out2) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n_chunks; i += blockDim.x) s += partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float norm = (float)sqrt(red[0]);
    const float coef = max_norm / (norm + 1e-6f);
    out2[0] = norm;
    out2[1] = coef < 1.f ? coef : (coef >= 1.f ? 1.f : coef);   // clamp(max = 1); NaN stays NaN
  }
}

struct AdamW {   // scalars evaluated on the host in double, as torch evaluates its python floats, then rounded once
  float decay;        // 1 - lr * weight_decay
  float w1, w2;       // 1 - beta1, 1 - beta2
  float beta2, eps;
  float step_size;    // lr / (1 - beta1^t)
  float bias2_sqrt;   // sqrt(1 - beta2^t)
};

__global__ void __launch_bounds__(256) opt_step_kernel(const OptTensor* __restrict__ tensors, const OptChunk* __restrict__ chunks, int n_chunks,
                                                       const float* __restrict__ clip2, AdamW h, unsigned long long* __restrict__ nan_count) {
  const int c = blockIdx.x;
  if (c >= n_chunks) return;
  const OptChunk ch = chunks[c];
  const OptTensor t = tensors[ch.tensor];
  const float coef = t.clip && clip2 ? clip2[1] : 1.f;
  const int64_t end = min(ch.begin + (int64_t)OPT_CHUNK, t.n);
  unsigned nans = 0;
  const bool clipped = t.clip && clip2;
  // one element: clip, scrub, AdamW.  Returns the updated (p, m, v); g is updated in place (the reference leaves the clipped,
  // scrubbed gradient behind)
  auto one = [&](float& g, float& p, float& m, float& v) {
    if (clipped) g = __fmul_rn(g, coef);                                           // clip_grad_norm_: grads.mul_(clip_coef_clamped)
    if (g != g) { g = 0.f; ++nans; }                                               // p.grad[torch.isnan(p.grad)] = 0
    p = __fmul_rn(p, h.decay);                                                     // param.mul_(1 - lr * weight_decay)
    m = fmaf(h.w1, __fsub_rn(g, m), m);                                            // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(__fmul_rn(h.w2, g), g, __fmul_rn(v, h.beta2));                        // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), h.bias2_sqrt), h.eps);  // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    p = fmaf(-h.step_size, __fdiv_rn(m, denom), p);                                // param.addcdiv_(exp_avg, denom, value = -step_size)
  };
  int64_t i0 = ch.begin;
  if ((((uintptr_t)t.p | (uintptr_t)t.g | (uintptr_t)t.m | (uintptr_t)t.v) & 15) == 0) {   // 16-byte vectors over the aligned body
    const int64_t nv = (end - ch.begin) >> 2;
    float4* p4 = reinterpret_cast<float4*>(t.p + ch.begin);
    float4* g4 = reinterpret_cast<float4*>(t.g + ch.begin);
    float4* m4 = reinterpret_cast<float4*>(t.m + ch.begin);
    float4* v4 = reinterpret_cast<float4*>(t.v + ch.begin);
    for (int64_t q = threadIdx.x; q < nv; q += blockDim.x) {
      float4 g = g4[q], p = p4[q], m = m4[q], v = v4[q];
      one(g.x, p.x, m.x, v.x); one(g.y, p.y, m.y, v.y); one(g.z, p.z, m.z, v.z); one(g.w, p.w, m.w, v.w);
      g4[q] = g; p4[q] = p; m4[q] = m; v4[q] = v;
    }
    i0 = ch.begin + (nv << 2);
  }
  for (int64_t i = i0 + threadIdx.x; i < end; i += blockDim.x) {
    float g = t.g[i], p = t.p[i], m = t.m[i], v = t.v[i];
    one(g, p, m, v);
    t.g[i] = g; t.p[i] = p; t.m[i] = m; t.v[i] = v;
  }
  if (nan_count && nans) atomicAdd(nan_count, (unsigned long long)nans);
}

}  // namespace

void q_sample(const float* x0, const float* noise, const int64_t* t, const float* sqrt_ac, const float* sqrt_1mac, int64_t rows, int64_t row_len,
              float* out, cudaStream_t s) {
  if (rows * row_len == 0) return;
  launch_pdl(q_sample_kernel, dim3(grid_for(rows * row_len, 256)), dim3(256), 0, s, x0, noise, t, sqrt_ac, sqrt_1mac, rows, row_len, out);
  ECHO_LAUNCH_CHECK();
}

void mse_rows(const float* pred, const float* target, int64_t rows, int64_t row_len, const int* ranges_dev, int n_ranges, float* out, cudaStream_t s) {
  if (rows == 0) return;
  mse_rows_kernel<<<cdiv(rows * 32, 256), 256, 0, s>>>(pred, target, rows, row_len, ranges_dev, n_ranges, out);
  ECHO_LAUNCH_CHECK();
}

}  // namespace echo

// ---- C handle of the optimizer ---------------------------------------------------------------------------------------------
struct echo_optimizer {
  echo::DevPool pool;
  echo::OptTensor* d_tensors = nullptr;
  echo::OptChunk* d_chunks = nullptr;
  double* d_partial = nullptr;
  float* d_clip = nullptr;               // [norm, coefficient]
  unsigned long long* d_nans = nullptr;
  std::vector<echo::OptTensor> host_tensors;
  int n_tensors = 0, n_chunks = 0;
  bool any_clip = false;
  int64_t n_params = 0, step = 0;
};

namespace echo {

echo_optimizer* optimizer_create(const echo_opt_tensor_t* tensors, int n) {
  ECHO_CHECK(tensors && n > 0, "optimizer_create: no tensors");
  echo_optimizer* h = new echo_optimizer();
  try {
    std::vector<OptTensor> ht(n);
    std::vector<OptChunk> hc;
    for (int i = 0; i < n; ++i) {
      const echo_opt_tensor_t& t = tensors[i];
      ECHO_CHECK(t.param && t.grad && t.exp_avg && t.exp_avg_sq && t.numel > 0, "optimizer_create: tensor %d has a null pointer or no elements", i);
      ht[i].p = t.param; ht[i].g = t.grad; ht[i].m = t.exp_avg; ht[i].v = t.exp_avg_sq; ht[i].n = t.numel; ht[i].clip = t.clip_group ? 1 : 0;
      ht[i].pad = 0;
      h->any_clip |= t.clip_group != 0;
      h->n_params += t.numel;
      for (int64_t b = 0; b < t.numel; b += OPT_CHUNK) hc.push_back({i, 0, b});
    }
    h->host_tensors = ht;
    h->n_tensors = n;
    h->n_chunks = (int)hc.size();
    h->d_tensors = (OptTensor*)h->pool.alloc(sizeof(OptTensor) * ht.size());
    h->d_chunks = (OptChunk*)h->pool.alloc(sizeof(OptChunk) * hc.size());
    h->d_partial = (double*)h->pool.alloc(sizeof(double) * hc.size());
    h->d_clip = (float*)h->pool.alloc(sizeof(float) * 4);
    h->d_nans = (unsigned long long*)h->pool.alloc(sizeof(unsigned long long) * 2);
    ECHO_CUDA(cudaMemcpy(h->d_tensors, ht.data(), sizeof(OptTensor) * ht.size(), cudaMemcpyHostToDevice));
    ECHO_CUDA(cudaMemcpy(h->d_chunks, hc.data(), sizeof(OptChunk) * hc.size(), cudaMemcpyHostToDevice));
    ECHO_CUDA(cudaMemset(h->d_nans, 0, sizeof(unsigned long long) * 2));
    ECHO_CUDA(cudaMemset(h->d_clip, 0, sizeof(float) * 4));
    return h;
  } catch (...) {
    h->pool.destroy();
    delete h;
    throw;
  }
}

void optimizer_destroy(echo_optimizer* h) {
  if (!h) return;
  h->pool.destroy();
  delete h;
}

// New parameter / gradient / state pointers for the same tensor list (torch's zero_grad(set_to_none=True) gives every step
// fresh gradient tensors): the chunk table stays, the pointer table is re-uploaded stream-ordered.
void optimizer_set_tensors(echo_optimizer* h, const echo_opt_tensor_t* tensors, int n, cudaStream_t s) {
  ECHO_CHECK(tensors && n == h->n_tensors, "optimizer_set_tensors: %d tensors, the optimizer was created for %d", n, h->n_tensors);
  for (int i = 0; i < n; ++i) {
    const echo_opt_tensor_t& t = tensors[i];
    ECHO_CHECK(t.param && t.grad && t.exp_avg && t.exp_avg_sq, "optimizer_set_tensors: tensor %d has a null pointer", i);
    ECHO_CHECK(t.numel == h->host_tensors[i].n && (t.clip_group != 0) == (h->host_tensors[i].clip != 0),
               "optimizer_set_tensors: tensor %d changed its size or clip group", i);
    h->host_tensors[i].p = t.param; h->host_tensors[i].g = t.grad; h->host_tensors[i].m = t.exp_avg; h->host_tensors[i].v = t.exp_avg_sq;
  }
  ECHO_CUDA(cudaMemcpyAsync(h->d_tensors, h->host_tensors.data(), sizeof(OptTensor) * n, cudaMemcpyHostToDevice, s));
  ECHO_CUDA(cudaStreamSynchronize(s));   // host_tensors may be rewritten by the next call before a pageable copy has been staged
}

// hyper-parameters arrive as doubles: torch evaluates 1 - beta2, lr / bias_correction1, ... on python floats and rounds the RESULT to
// fp32 (1 - 0.999f would be off by 1.3e-5)
void optimizer_step(echo_optimizer* h, int64_t step, double lr, double beta1, double beta2, double eps, double weight_decay, double clip_max_norm,
                    cudaStream_t s) {
  ECHO_CHECK(lr >= 0. && beta1 >= 0. && beta1 < 1. && beta2 >= 0. && beta2 < 1. && eps >= 0., "optimizer_step: bad hyper-parameters");
  ECHO_CHECK(step >= 1, "optimizer_step: step numbers start at 1 (torch's state['step'] after the update)");
  h->step = step;
  const bool clip = h->any_clip && clip_max_norm > 0.;
  if (clip) {
    opt_sumsq_kernel<<<h->n_chunks, 256, 0, s>>>(h->d_tensors, h->d_chunks, h->n_chunks, h->d_partial);
    ECHO_LAUNCH_CHECK();
    opt_clip_coef_kernel<<<1, 256, 0, s>>>(h->d_partial, h->n_chunks, (float)clip_max_norm, h->d_clip);
    ECHO_LAUNCH_CHECK();
  }
  AdamW a;
  // torch: bias_correction1 = 1 - beta1 ** step; step_size = lr / bias_correction1; bias_correction2_sqrt = sqrt(1 - beta2 ** step)
  a.decay = (float)(1.0 - lr * weight_decay);
  a.w1 = (float)(1.0 - beta1);
  a.w2 = (float)(1.0 - beta2);
  a.beta2 = (float)beta2;
  a.eps = (float)eps;
  a.step_size = (float)(lr / (1.0 - pow(beta1, (double)h->step)));
  a.bias2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)h->step));
  opt_step_kernel<<<h->n_chunks, 256, 0, s>>>(h->d_tensors, h->d_chunks, h->n_chunks, clip ? h->d_clip : nullptr, a, h->d_nans);
  ECHO_LAUNCH_CHECK();
}

void optimizer_info(const echo_optimizer* h, int64_t* out4, float* clip2, cudaStream_t s) {
  unsigned long long nans = 0;
  ECHO_CUDA(cudaMemcpyAsync(&nans, h->d_nans, sizeof(nans), cudaMemcpyDeviceToHost, s));
  if (clip2) ECHO_CUDA(cudaMemcpyAsync(clip2, h->d_clip, sizeof(float) * 2, cudaMemcpyDeviceToHost, s));
  ECHO_CUDA(cudaStreamSynchronize(s));
  out4[0] = h->step; out4[1] = h->n_params; out4[2] = h->n_chunks; out4[3] = (int64_t)nans;
}

}  // namespace echo
