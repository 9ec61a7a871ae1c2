// L0 hard-concrete gate kernels (xvlm_l0_module.py:174-271), flat-arena AdamW / grad-norm clip (optim.py:23-69,
// apex_ddp_accelerator.py:98-101) and the small fp32 SIMT GEMM used for the precision-critical ITC similarity.
#include "evlm_common.cuh"
#include "../../include/evlm.h"
#include <atomic>
#include <math.h>

namespace evlm {
extern std::atomic<unsigned long long> g_launch_count;

constexpr float L0_L = -0.1f, L0_R = 1.1f, L0_EPS = 1e-6f;

// z = hardtanh(sigmoid((log u - log(1-u) + loga) / beta) * (r - l) + l, 0, 1)     (:180-182,239-250)
__global__ void l0_sample_fwd_kernel(const float* __restrict__ loga, const float* __restrict__ u, float* __restrict__ z, int64_t n, float beta) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = (logf(u[i]) - logf(1.f - u[i]) + loga[i]) / beta;
  const float y = 1.f / (1.f + expf(-x));
  z[i] = fminf(fmaxf(y * (L0_R - L0_L) + L0_L, 0.f), 1.f);
}
__global__ void l0_sample_bwd_kernel(const float* __restrict__ loga, const float* __restrict__ u, const float* __restrict__ dz,
                                     float* __restrict__ dloga, int64_t n, float beta) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = (logf(u[i]) - logf(1.f - u[i]) + loga[i]) / beta;
  const float y = 1.f / (1.f + expf(-x));
  const float zz = y * (L0_R - L0_L) + L0_L;
  const float pass = (zz > 0.f && zz < 1.f) ? 1.f : 0.f;  // hardtanh gradient (zero where it clamps)
  dloga[i] = dz[i] * pass * (L0_R - L0_L) * y * (1.f - y) / beta;
}
// score = 1 - clamp(sigmoid(beta * log(xn/(1-xn)) - loga), eps, 1-eps), xn = (0 - l)/(r - l)            (:174-178)
__device__ __forceinline__ float l0_cdf0(float loga, float beta, float logit0, float& dcdf) {
  const float s = 1.f / (1.f + expf(-(logit0 * beta - loga)));
  const bool inside = (s > L0_EPS) && (s < 1.f - L0_EPS);
  dcdf = inside ? -s * (1.f - s) : 0.f;  // d cdf / d loga
  return fminf(fmaxf(s, L0_EPS), 1.f - L0_EPS);
}
__global__ void __launch_bounds__(256) l0_expected_fwd_kernel(const float* __restrict__ loga, int64_t n, float beta, float logit0, float weight,
                                                              float* __restrict__ out) {
  __shared__ float red[32];
  float acc = 0.f, d;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc += 1.f - l0_cdf0(loga[i], beta, logit0, d);
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, acc * weight);
}
__global__ void l0_expected_bwd_kernel(const float* __restrict__ loga, int64_t n, float beta, float logit0, float weight,
                                       const float* __restrict__ g, float* __restrict__ dloga) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float d;
  l0_cdf0(loga[i], beta, logit0, d);
  dloga[i] += g[0] * weight * (-d);
}

// Deterministic eval mask.  n0 = round_half_even(size - sum(score)); the n0 entries with the smallest soft score
// sigmoid(loga / beta * magic) are zeroed; ties -> lower index first.  Rank by counting:
// rank_i = #{j : soft_j < soft_i or (soft_j == soft_i and j < i)}; mask_i = rank_i >= n0.
// grid = (layer rows, slices of 256 entries): every block stages the whole row in shared memory and recomputes n0 with the same
// summation order (bit-identical across the blocks of a row), then each thread ranks ONE entry — 3072 broadcast reads instead of
// the 12 x 3072 a single block per row needed (0.2 ms per gate type at 3072 columns, 13% of a VQA inference step).
__global__ void __launch_bounds__(256) l0_deterministic_kernel(const float* __restrict__ loga, float* __restrict__ mask, int32_t* __restrict__ kept,
                                                               int size, float beta, float logit0, float magic) {
  extern __shared__ float soft[];
  __shared__ float red[32];
  __shared__ int n0_s;
  const float* la = loga + (int64_t)blockIdx.x * size;
  float acc = 0.f, d;
  for (int i = threadIdx.x; i < size; i += blockDim.x) {
    acc += 1.f - l0_cdf0(la[i], beta, logit0, d);
    soft[i] = 1.f / (1.f + expf(-(la[i] / beta * magic)));
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    const float ez = (float)size - acc;          // expected number of zeros (fp32, as the reference's .item())
    n0_s = (int)rintf(ez);                       // round-half-even == python round()
  }
  __syncthreads();
  const int n0 = n0_s;
  int kept_local = 0;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < size; i += gridDim.y * blockDim.x) {
    float m = 1.f;
    if (n0 > 0) {
      const float si = soft[i];
      int rank = 0;
      for (int j = 0; j < size; ++j) {
        const float sj = soft[j];
        rank += (sj < si || (sj == si && j < i)) ? 1 : 0;
      }
      m = rank >= n0 ? 1.f : 0.f;
    }
    mask[(int64_t)blockIdx.x * size + i] = m;
    kept_local += m > 0.f;
  }
  if (kept) {
    const float tot = block_sum((float)kept_local, red);
    if (threadIdx.x == 0) atomicAdd(kept + blockIdx.x, (int)(tot + 0.5f));     // zeroed by the host wrapper
  }
}

__global__ void clamp_kernel(float* __restrict__ x, int64_t n, float lo, float hi) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = fminf(fmaxf(x[i], lo), hi);
}

// ------------------------------------------------------------------------------------------------ optimizer
// Sum of squares, DETERMINISTIC: every data-parallel rank must get bit-identical clip coefficients from bit-identical (all-reduced)
// gradients, or the replicas drift apart (tests/test_gpu_distributed.py).  Block partials go to a scratch array; the block that
// finishes last adds them in a fixed order (no float atomics whose order varies from run to run) and accumulates into `out`.
// The scratch is per process (= per GPU); calls are ordered on one stream (the optimizer's), never concurrent.
constexpr int SUMSQ_MAX_BLOCKS = 148 * 4;
__device__ float g_sumsq_partial[SUMSQ_MAX_BLOCKS];
__device__ unsigned int g_sumsq_done = 0;
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  __shared__ float red[32];
  __shared__ bool last;
  float acc = 0.f;
  const int64_t n4 = n >> 2;
  if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
      const float4 v = *reinterpret_cast<const float4*>(x + i * 4);
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc += x[i] * x[i];
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc += x[i] * x[i];
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    g_sumsq_partial[blockIdx.x] = acc;
    __threadfence();
    last = atomicAdd(&g_sumsq_done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float t = 0.f;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += __ldcg(&g_sumsq_partial[i]);
  __syncthreads();
  t = block_sum(t, red);
  if (threadIdx.x == 0) {
    out[0] += t;
    g_sumsq_done = 0;
  }
}
__global__ void clip_coef_kernel(const float* __restrict__ sumsq, float max_norm, float* __restrict__ coef) {
  const float c = max_norm / (sqrtf(sumsq[0]) + 1e-6f);
  coef[0] = c < 1.f ? c : 1.f;
}
// HF AdamW (transformers.optimization.AdamW, call site optim.py:67): m,v EMA; step_size = lr*sqrt(1-b2^t)/(1-b1^t);
// p -= step_size * m / (sqrt(v) + eps); then decoupled decay p -= lr * wd * p (applied AFTER the Adam update).
// `hyper` (optional, device): {step_size, lr * weight_decay} of this group, refreshed by the host before a graph replay.
__global__ void __launch_bounds__(256) adamw_kernel(evlm_adamw_group G, const float* __restrict__ grad_scale, float step_size,
                                                    const float* __restrict__ hyper) {
  const float gs = grad_scale ? grad_scale[0] : 1.f;
  float decay = G.lr * G.weight_decay;
  if (hyper) {
    step_size = hyper[0];
    decay = hyper[1];
  }
  auto upd = [&](float g, float& m, float& v, float& p) {
    g *= gs;
    m = G.beta1 * m + (1.f - G.beta1) * g;
    v = G.beta2 * v + (1.f - G.beta2) * g * g;
    p -= step_size * m / (sqrtf(v) + G.eps);
    if (G.weight_decay != 0.f) p -= decay * p;
  };
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  const bool vec = (G.n & 3) == 0 && ((reinterpret_cast<uintptr_t>(G.p) | reinterpret_cast<uintptr_t>(G.g) | reinterpret_cast<uintptr_t>(G.m) |
                                       reinterpret_cast<uintptr_t>(G.v)) & 15) == 0 && (reinterpret_cast<uintptr_t>(G.p_bf16) & 7) == 0;
  if (vec) {   // 16-byte accesses: the arenas are 256-byte aligned and padded to 64 floats (optim.py: ARENA_ALIGN)
    const int64_t n4 = G.n >> 2;
    for (int64_t i = tid; i < n4; i += stride) {
      const float4 g4 = reinterpret_cast<const float4*>(G.g)[i];
      float4 m4 = reinterpret_cast<float4*>(G.m)[i], v4 = reinterpret_cast<float4*>(G.v)[i], p4 = reinterpret_cast<float4*>(G.p)[i];
      upd(g4.x, m4.x, v4.x, p4.x);
      upd(g4.y, m4.y, v4.y, p4.y);
      upd(g4.z, m4.z, v4.z, p4.z);
      upd(g4.w, m4.w, v4.w, p4.w);
      reinterpret_cast<float4*>(G.m)[i] = m4;
      reinterpret_cast<float4*>(G.v)[i] = v4;
      reinterpret_cast<float4*>(G.p)[i] = p4;
      if (G.p_bf16) reinterpret_cast<uint2*>(G.p_bf16)[i] = make_uint2(pack_bf16x2(p4.x, p4.y), pack_bf16x2(p4.z, p4.w));
    }
    return;
  }
  for (int64_t i = tid; i < G.n; i += stride) {
    float m = G.m[i], v = G.v[i], p = G.p[i];
    upd(G.g[i], m, v, p);
    G.m[i] = m;
    G.v[i] = v;
    G.p[i] = p;
    if (G.p_bf16) reinterpret_cast<__nv_bfloat16*>(G.p_bf16)[i] = __float2bfloat16(p);
  }
}

// ------------------------------------------------------------------------------------------------ fp32 SIMT GEMM (tiny problems)
// D[M,N] = alpha * opA(A) * opB(B) + beta * D ; 16x16 tiles, one output per thread.
__global__ void sgemm_kernel(int M, int N, int K, float alpha, const float* __restrict__ A, int64_t lda, int a_trans,
                             const float* __restrict__ B, int64_t ldb, int b_trans, float beta, float* __restrict__ D, int64_t ldd,
                             const float* __restrict__ alpha_dev, int alpha_dev_inv) {
  if (alpha_dev) alpha = alpha_dev_inv ? alpha / alpha_dev[0] : alpha * alpha_dev[0];
  __shared__ float sA[16][17], sB[16][17];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int m = blockIdx.y * 16 + ty, n = blockIdx.x * 16 + tx;
  float acc = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    const int ka = k0 + tx, kb = k0 + ty;
    sA[ty][tx] = (m < M && ka < K) ? (a_trans ? A[(int64_t)ka * lda + m] : A[(int64_t)m * lda + ka]) : 0.f;
    const int nb = blockIdx.x * 16 + tx;
    sB[ty][tx] = (nb < N && kb < K) ? (b_trans ? B[(int64_t)nb * ldb + kb] : B[(int64_t)kb * ldb + nb]) : 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc += sA[ty][k] * sB[k][tx];
    __syncthreads();
  }
  if (m < M && n < N) {
    float* d = D + (int64_t)m * ldd + n;
    *d = alpha * acc + (beta != 0.f ? beta * (*d) : 0.f);
  }
}
__global__ void __launch_bounds__(256) dot_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t n, float scale,
                                                  float* __restrict__ out) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc += x[i] * y[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, acc * scale);
}
}  // namespace evlm
using namespace evlm;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define COUNT(n) g_launch_count.fetch_add(n, std::memory_order_relaxed)
static inline float l0_logit0() {
  const double xn = (0.0 - (-0.1)) / (1.1 - (-0.1));
  return (float)(log(xn) - log(1.0 - xn));
}

extern "C" int evlm_l0_sample_fwd(const float* loga, const float* u, float* z, int64_t n, float temperature, void* stream) {
  if (!loga || !u || !z || n < 0 || temperature <= 0.f) return EVLM_EINVAL;
  if (n == 0) return EVLM_OK;
  l0_sample_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(loga, u, z, n, temperature);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
extern "C" int evlm_l0_sample_bwd(const float* loga, const float* u, const float* dz, float* dloga, int64_t n, float temperature, void* stream) {
  if (!loga || !u || !dz || !dloga || n < 0 || temperature <= 0.f) return EVLM_EINVAL;
  if (n == 0) return EVLM_OK;
  l0_sample_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(loga, u, dz, dloga, n, temperature);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
extern "C" int evlm_l0_expected_fwd(const float* loga, int64_t n, float temperature, float weight, float* out, int32_t accumulate, void* stream) {
  if (!loga || !out || n < 0) return EVLM_EINVAL;
  if (!accumulate) cudaMemsetAsync(out, 0, sizeof(float), ST(stream));
  if (n == 0) return EVLM_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148) blocks = 148;
  l0_expected_fwd_kernel<<<(unsigned)blocks, 256, 0, ST(stream)>>>(loga, n, temperature, l0_logit0(), weight, out);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
extern "C" int evlm_l0_expected_bwd(const float* loga, int64_t n, float temperature, float weight, const float* g, float* dloga, void* stream) {
  if (!loga || !g || !dloga || n < 0) return EVLM_EINVAL;
  if (n == 0) return EVLM_OK;
  l0_expected_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(loga, n, temperature, l0_logit0(), weight, g, dloga);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
extern "C" int evlm_l0_deterministic(const float* loga, float* mask, int32_t* kept_count, int layers, int size, float temperature,
                                     float magical_number, void* stream) {
  if (!loga || !mask || layers <= 0 || size <= 0 || size > 12000) return EVLM_EINVAL;
  if (kept_count) {
    cudaError_t e = cudaMemsetAsync(kept_count, 0, (size_t)layers * sizeof(int32_t), ST(stream));
    if (e != cudaSuccess) return (int)e;
  }
  dim3 grid(layers, (size + 255) / 256);
  l0_deterministic_kernel<<<grid, 256, size * sizeof(float), ST(stream)>>>(loga, mask, kept_count, size, temperature, l0_logit0(), magical_number);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
extern "C" int evlm_clamp_(float* x, int64_t n, float lo, float hi, void* stream) {
  if (!x || n < 0) return EVLM_EINVAL;
  if (n == 0) return EVLM_OK;
  clamp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(x, n, lo, hi);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
extern "C" int evlm_sumsq(const float* x, int64_t n, float* out, void* stream) {
  if (!x || !out || n < 0) return EVLM_EINVAL;
  if (n == 0) return EVLM_OK;
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > SUMSQ_MAX_BLOCKS) blocks = SUMSQ_MAX_BLOCKS;
  if (blocks < 1) blocks = 1;
  sumsq_kernel<<<(unsigned)blocks, 256, 0, ST(stream)>>>(x, n, out);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
extern "C" int evlm_clip_coef(const float* sumsq, float max_norm, float* coef, void* stream) {
  if (!sumsq || !coef) return EVLM_EINVAL;
  clip_coef_kernel<<<1, 1, 0, ST(stream)>>>(sumsq, max_norm, coef);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
static int adamw_launch(const evlm_adamw_group* groups_host, int ngroups, const float* grad_scale_dev, const float* hyper_dev, void* stream) {
  if (!groups_host || ngroups <= 0) return EVLM_EINVAL;
  for (int i = 0; i < ngroups; ++i) {
    const evlm_adamw_group& G = groups_host[i];
    if (!G.p || !G.g || !G.m || !G.v || G.n < 0 || G.step <= 0) return EVLM_EINVAL;
    if (G.n == 0) continue;
    const double bc1 = 1.0 - pow((double)G.beta1, (double)G.step);
    const double bc2 = 1.0 - pow((double)G.beta2, (double)G.step);
    const float step_size = (float)((double)G.lr * sqrt(bc2) / bc1);
    int64_t blocks = (G.n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    adamw_kernel<<<(unsigned)blocks, 256, 0, ST(stream)>>>(G, grad_scale_dev, step_size, hyper_dev ? hyper_dev + 2 * i : nullptr);
    COUNT(1);
  }
  EVLM_CUDA_RETURN();
}
extern "C" int evlm_adamw_step(const evlm_adamw_group* groups_host, int ngroups, const float* grad_scale_dev, void* stream) {
  return adamw_launch(groups_host, ngroups, grad_scale_dev, nullptr, stream);
}
extern "C" int evlm_adamw_step_dev(const evlm_adamw_group* groups_host, int ngroups, const float* grad_scale_dev, const float* hyper_dev,
                                   void* stream) {
  if (!hyper_dev) return EVLM_EINVAL;
  return adamw_launch(groups_host, ngroups, grad_scale_dev, hyper_dev, stream);
}
extern "C" int evlm_sgemm(int M, int N, int K, float alpha, const float* A, int64_t lda, int a_trans, const float* B, int64_t ldb, int b_trans,
                          float beta, float* D, int64_t ldd, const float* alpha_dev, int alpha_dev_inv, void* stream) {
  if (!A || !B || !D || M <= 0 || N <= 0 || K <= 0) return EVLM_EINVAL;
  dim3 grid((N + 15) / 16, (M + 15) / 16), blk(16, 16);
  sgemm_kernel<<<grid, blk, 0, ST(stream)>>>(M, N, K, alpha, A, lda, a_trans, B, ldb, b_trans, beta, D, ldd, alpha_dev, alpha_dev_inv);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
extern "C" int evlm_dot(const float* x, const float* y, int64_t n, float scale, float* out, int32_t accumulate, void* stream) {
  if (!x || !y || !out || n < 0) return EVLM_EINVAL;
  if (!accumulate) cudaMemsetAsync(out, 0, sizeof(float), ST(stream));
  if (n == 0) return EVLM_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  dot_kernel<<<(unsigned)blocks, 256, 0, ST(stream)>>>(x, y, n, scale, out);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
