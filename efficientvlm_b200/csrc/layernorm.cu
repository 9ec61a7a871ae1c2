// LayerNorm forward / backward (HBM-bound), one warp per row, row cached in registers.
// Replaces nn.LayerNorm call sites eff_vit.py:252,264,452,467 and eff_bert.py:213,380,461,725 and the
// residual-gradient adds autograd would otherwise issue as separate kernels (dres is fused into dx).
// NV = float4 chunks per lane is a template parameter (H <= 128*NV) so the 768-wide rows of the hot path use half the
// registers of the 1536-wide ITM-head rows.
#include "evlm_common.cuh"
#include "../../include/evlm.h"
#include <atomic>
#include <cstdlib>

namespace evlm {
extern std::atomic<unsigned long long> g_launch_count;

constexpr int LN_MAX_H = 1536;

template <bool X_BF16>
__device__ __forceinline__ float4 ld4(const void* base, int64_t off) {
  if (X_BF16) {
    const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + off);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
}
__device__ __forceinline__ void st4_bf16(void* base, int64_t off, float4 v) {
  *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + off) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

template <bool X_BF16, int NV>
__global__ void __launch_bounds__(128) ln_fwd_kernel(const void* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     float eps, float* __restrict__ y32, void* __restrict__ y16, float* __restrict__ mean_o,
                                                     float* __restrict__ rstd_o, int64_t rows, int H, float p, uint64_t seed, uint32_t sid) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nv = H >> 2;  // float4 chunks in the row
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      v[i] = ld4<X_BF16>(x, row * H + c * 4);
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  const float mean = warp_sum(s) / H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      q += a * a + b * b + cc * cc + d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / H + eps);
  if (lane == 0) {
    if (mean_o) mean_o[row] = mean;
    if (rstd_o) rstd_o[row] = rstd;
  }
  const float keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c * 4));
      const float4 bt = __ldg(reinterpret_cast<const float4*>(beta + c * 4));
      float4 o;
      o.x = (v[i].x - mean) * rstd * gm.x + bt.x;
      o.y = (v[i].y - mean) * rstd * gm.y + bt.y;
      o.z = (v[i].z - mean) * rstd * gm.z + bt.z;
      o.w = (v[i].w - mean) * rstd * gm.w + bt.w;
      if (p > 0.f) {
        const float4 u = dropout_uniform4(seed + rng_offset(), sid, (uint64_t)(row * H + c * 4) >> 2);
        o.x = u.x >= p ? o.x * keep : 0.f;
        o.y = u.y >= p ? o.y * keep : 0.f;
        o.z = u.z >= p ? o.z * keep : 0.f;
        o.w = u.w >= p ? o.w * keep : 0.f;
      }
      if (y32) *reinterpret_cast<float4*>(y32 + row * H + c * 4) = o;
      if (y16) st4_bf16(y16, row * H + c * 4, o);
    }
  }
}

// Backward. Each block handles a strided set of rows with 4 warps; per-lane column partials of dgamma / dbeta are
// reduced across the block in shared memory and added to global memory with one atomic per column per block.
template <bool DY_BF16, bool X_BF16, int NV>
__global__ void __launch_bounds__(128) ln_bwd_kernel(const void* __restrict__ dy, const void* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const float* __restrict__ dres, float* __restrict__ dx32, void* __restrict__ dx16,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows, int H, float p,
                                                     uint64_t seed, uint32_t sid) {
  extern __shared__ float sred[];  // [2][4 warps][H]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = H >> 2;
  float4 dg[NV], db[NV], gm[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    dg[i] = db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int c = lane + i * 32;
    gm[i] = c < nv ? __ldg(reinterpret_cast<const float4*>(gamma + c * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  for (int64_t row = (int64_t)blockIdx.x * 4 + warp; row < rows; row += (int64_t)gridDim.x * 4) {
    const float mu = mean[row], rs = rstd[row];
    float4 g[NV], xh[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nv) {
        float4 d = ld4<DY_BF16>(dy, row * H + c * 4);
        if (p > 0.f) {
          const float4 u = dropout_uniform4(seed + rng_offset(), sid, (uint64_t)(row * H + c * 4) >> 2);
          d.x = u.x >= p ? d.x * keep : 0.f;
          d.y = u.y >= p ? d.y * keep : 0.f;
          d.z = u.z >= p ? d.z * keep : 0.f;
          d.w = u.w >= p ? d.w * keep : 0.f;
        }
        const float4 xv = ld4<X_BF16>(x, row * H + c * 4);
        xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        g[i] = make_float4(d.x * gm[i].x, d.y * gm[i].y, d.z * gm[i].z, d.w * gm[i].w);
        s1 += g[i].x + g[i].y + g[i].z + g[i].w;
        s2 += g[i].x * xh[i].x + g[i].y * xh[i].y + g[i].z * xh[i].z + g[i].w * xh[i].w;
        dg[i].x += d.x * xh[i].x; dg[i].y += d.y * xh[i].y; dg[i].z += d.z * xh[i].z; dg[i].w += d.w * xh[i].w;
        db[i].x += d.x; db[i].y += d.y; db[i].z += d.z; db[i].w += d.w;
      }
    }
    s1 = warp_sum(s1) / H;
    s2 = warp_sum(s2) / H;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nv) {
        float4 o;
        o.x = rs * (g[i].x - s1 - xh[i].x * s2);
        o.y = rs * (g[i].y - s1 - xh[i].y * s2);
        o.z = rs * (g[i].z - s1 - xh[i].z * s2);
        o.w = rs * (g[i].w - s1 - xh[i].w * s2);
        if (dres) {
          const float4 r = *reinterpret_cast<const float4*>(dres + row * H + c * 4);
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        if (dx32) *reinterpret_cast<float4*>(dx32 + row * H + c * 4) = o;
        if (dx16) st4_bf16(dx16, row * H + c * 4, o);
      }
    }
  }
  if (dgamma == nullptr && dbeta == nullptr) return;
  float* sg = sred + warp * H;
  float* sb = sred + (4 + warp) * H;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      *reinterpret_cast<float4*>(sg + c * 4) = dg[i];
      *reinterpret_cast<float4*>(sb + c * 4) = db[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    const float a = sred[c] + sred[H + c] + sred[2 * H + c] + sred[3 * H + c];
    const float b = sred[4 * H + c] + sred[5 * H + c] + sred[6 * H + c] + sred[7 * H + c];
    if (dgamma) atomicAdd(dgamma + c, a);
    if (dbeta) atomicAdd(dbeta + c, b);
  }
}

// Column-owner backward for rows whose width is a multiple of 128 (the hot path's 768): one block of H/4 threads, thread t owns
// columns [4t, 4t+4) of EVERY row the block visits, four rows per iteration.  Compared with the warp-per-row kernel above:
//   * all of an iteration's loads (dy, x, dres of four rows: up to 160 bytes per thread) are in flight together, one memory
//     round trip per four rows instead of two per row, at half the registers (no per-lane copy of the whole row);
//   * dgamma / dbeta partials are 8 registers per thread and leave as one atomicAdd per column and block.
// The two row sums of the four rows are reduced by warp shuffles + one shared-memory exchange per iteration.
constexpr int LNC_R = 4;
template <bool DY_BF16, bool X_BF16>
__global__ void __launch_bounds__(256, 3) ln_bwd_cols_kernel(const void* __restrict__ dy, const void* __restrict__ x, const float* __restrict__ gamma,
                                                          const float* __restrict__ mean, const float* __restrict__ rstd,
                                                          const float* __restrict__ dres, float* __restrict__ dx32, void* __restrict__ dx16,
                                                          float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows, int H, float p,
                                                          uint64_t seed, uint32_t sid, float p_out, uint32_t sid_out,
                                                          float* __restrict__ dcolsum) {
  // p_out / sid_out / dcolsum (evlm_layernorm_bwd_ex): the bf16 copy of dx is what the Linear in front of this LayerNorm receives as its
  // output gradient.  That Linear's output went through dropout (stream sid_out) in the forward, so the copy gets the same mask
  // replayed here, and its column sums ARE that Linear's bias gradient: one pass instead of a cast kernel and a colsum kernel.
  __shared__ float sred[2][8][2 * LNC_R];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nw = blockDim.x >> 5;
  const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + t);
  const float keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  const float inv_h = 1.f / (float)H;
  const uint64_t sd = seed + rng_offset();
  float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = dg, dcs = dg;
  const float keep_out = p_out > 0.f ? 1.f / (1.f - p_out) : 1.f;
  int buf = 0;
  for (int64_t row0 = (int64_t)blockIdx.x * LNC_R; row0 < rows; row0 += (int64_t)gridDim.x * LNC_R) {
    float4 d[LNC_R], xh[LNC_R], rd[LNC_R];
    float rs[LNC_R], mu[LNC_R];
#pragma unroll
    for (int r = 0; r < LNC_R; ++r) {
      const int64_t row = row0 + r;
      const bool ok = row < rows;
      const int64_t off = row * H + t * 4;
      d[r] = ok ? ld4<DY_BF16>(dy, off) : make_float4(0.f, 0.f, 0.f, 0.f);
      xh[r] = ok ? ld4<X_BF16>(x, off) : make_float4(0.f, 0.f, 0.f, 0.f);
      rd[r] = (ok && dres) ? *reinterpret_cast<const float4*>(dres + off) : make_float4(0.f, 0.f, 0.f, 0.f);
      mu[r] = ok ? mean[row] : 0.f;
      rs[r] = ok ? rstd[row] : 0.f;
    }
    float part[2 * LNC_R];
#pragma unroll
    for (int r = 0; r < LNC_R; ++r) {
      if (p > 0.f) {
        const float4 u = dropout_uniform4(sd, sid, (uint64_t)((row0 + r) * H + t * 4) >> 2);
        d[r].x = u.x >= p ? d[r].x * keep : 0.f;
        d[r].y = u.y >= p ? d[r].y * keep : 0.f;
        d[r].z = u.z >= p ? d[r].z * keep : 0.f;
        d[r].w = u.w >= p ? d[r].w * keep : 0.f;
      }
      xh[r] = make_float4((xh[r].x - mu[r]) * rs[r], (xh[r].y - mu[r]) * rs[r], (xh[r].z - mu[r]) * rs[r], (xh[r].w - mu[r]) * rs[r]);
      const float4 g = make_float4(d[r].x * gm.x, d[r].y * gm.y, d[r].z * gm.z, d[r].w * gm.w);
      part[r] = g.x + g.y + g.z + g.w;
      part[LNC_R + r] = g.x * xh[r].x + g.y * xh[r].y + g.z * xh[r].z + g.w * xh[r].w;
      dg.x += d[r].x * xh[r].x; dg.y += d[r].y * xh[r].y; dg.z += d[r].z * xh[r].z; dg.w += d[r].w * xh[r].w;
      db.x += d[r].x; db.y += d[r].y; db.z += d[r].z; db.w += d[r].w;
    }
    // Eight warp sums at once: at every butterfly step a lane keeps half of its remaining values and hands the other half to its
    // partner (4 + 2 + 1 shuffles), then the single survivor is summed over the last two steps (2 shuffles): 9 shuffles instead of
    // 8 x 5.  Lane l ends up with the total of value index ((l >> 4) & 1) * 4 + ((l >> 3) & 1) * 2 + ((l >> 2) & 1).
    {
      static_assert(LNC_R == 4, "the transposed reduction below is written for 8 values");
      const bool hi16 = lane & 16, hi8 = lane & 8, hi4 = lane & 4;
      float q4[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float keep = hi16 ? part[4 + k] : part[k], give = hi16 ? part[k] : part[4 + k];
        q4[k] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
      }
      float q2[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float keep = hi8 ? q4[2 + k] : q4[k], give = hi8 ? q4[k] : q4[2 + k];
        q2[k] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
      }
      float q1 = (hi4 ? q2[1] : q2[0]) + __shfl_xor_sync(0xffffffffu, hi4 ? q2[0] : q2[1], 4);
      q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
      q1 += __shfl_xor_sync(0xffffffffu, q1, 1);
      if ((lane & 3) == 0) sred[buf][warp][(hi16 ? 4 : 0) + (hi8 ? 2 : 0) + (hi4 ? 1 : 0)] = q1;
    }
    __syncthreads();
    float tot[2 * LNC_R];
#pragma unroll
    for (int k = 0; k < 2 * LNC_R; ++k) tot[k] = 0.f;
    for (int w = 0; w < nw; ++w) {
      const float4 a = *reinterpret_cast<const float4*>(&sred[buf][w][0]);
      const float4 b = *reinterpret_cast<const float4*>(&sred[buf][w][4]);
      tot[0] += a.x; tot[1] += a.y; tot[2] += a.z; tot[3] += a.w;
      tot[4] += b.x; tot[5] += b.y; tot[6] += b.z; tot[7] += b.w;
    }
    buf ^= 1;
#pragma unroll
    for (int r = 0; r < LNC_R; ++r) {
      const int64_t row = row0 + r;
      if (row >= rows) break;
      const float s1 = tot[r] * inv_h, s2 = tot[LNC_R + r] * inv_h;
      float4 o;
      o.x = rs[r] * (d[r].x * gm.x - s1 - xh[r].x * s2) + rd[r].x;
      o.y = rs[r] * (d[r].y * gm.y - s1 - xh[r].y * s2) + rd[r].y;
      o.z = rs[r] * (d[r].z * gm.z - s1 - xh[r].z * s2) + rd[r].z;
      o.w = rs[r] * (d[r].w * gm.w - s1 - xh[r].w * s2) + rd[r].w;
      const int64_t off = row * H + t * 4;
      if (dx32) *reinterpret_cast<float4*>(dx32 + off) = o;
      if (dx16) {
        if (p_out > 0.f) {
          const float4 u = dropout_uniform4(sd, sid_out, (uint64_t)off >> 2);
          o.x = u.x >= p_out ? o.x * keep_out : 0.f;
          o.y = u.y >= p_out ? o.y * keep_out : 0.f;
          o.z = u.z >= p_out ? o.z * keep_out : 0.f;
          o.w = u.w >= p_out ? o.w * keep_out : 0.f;
        }
        st4_bf16(dx16, off, o);
        if (dcolsum) {   // sums of the ROUNDED values (what a colsum over the bf16 tensor would add up)
          dcs.x += __bfloat162float(__float2bfloat16(o.x)); dcs.y += __bfloat162float(__float2bfloat16(o.y));
          dcs.z += __bfloat162float(__float2bfloat16(o.z)); dcs.w += __bfloat162float(__float2bfloat16(o.w));
        }
      }
    }
  }
  if (dcolsum && dx16) {
    atomicAdd(dcolsum + t * 4, dcs.x); atomicAdd(dcolsum + t * 4 + 1, dcs.y); atomicAdd(dcolsum + t * 4 + 2, dcs.z); atomicAdd(dcolsum + t * 4 + 3, dcs.w);
  }
  if (dgamma) {
    atomicAdd(dgamma + t * 4, dg.x); atomicAdd(dgamma + t * 4 + 1, dg.y); atomicAdd(dgamma + t * 4 + 2, dg.z); atomicAdd(dgamma + t * 4 + 3, dg.w);
  }
  if (dbeta) {
    atomicAdd(dbeta + t * 4, db.x); atomicAdd(dbeta + t * 4 + 1, db.y); atomicAdd(dbeta + t * 4 + 2, db.z); atomicAdd(dbeta + t * 4 + 3, db.w);
  }
}

template <int NV>
static void launch_fwd(bool xb, unsigned grid, cudaStream_t st, const void* x, const float* gamma, const float* beta, float eps, float* y_f32,
                       void* y_bf16, float* mean, float* rstd, int64_t rows, int H, float p, uint64_t seed, uint32_t sid) {
  if (xb) ln_fwd_kernel<true, NV><<<grid, 128, 0, st>>>(x, gamma, beta, eps, y_f32, y_bf16, mean, rstd, rows, H, p, seed, sid);
  else ln_fwd_kernel<false, NV><<<grid, 128, 0, st>>>(x, gamma, beta, eps, y_f32, y_bf16, mean, rstd, rows, H, p, seed, sid);
}
template <bool DB, bool XB, int NV>
static void launch_bwd1(unsigned grid, size_t smem, cudaStream_t st, const void* dy, const void* x, const float* gamma, const float* mean,
                        const float* rstd, const float* dres, float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta, int64_t rows, int H,
                        float p, uint64_t seed, uint32_t sid) {
  if (smem > 48 * 1024) cudaFuncSetAttribute(ln_bwd_kernel<DB, XB, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ln_bwd_kernel<DB, XB, NV><<<grid, 128, smem, st>>>(dy, x, gamma, mean, rstd, dres, dx_f32, dx_bf16, dgamma, dbeta, rows, H, p, seed, sid);
}
template <int NV>
static void launch_bwd(bool db, bool xb, unsigned grid, size_t smem, cudaStream_t st, const void* dy, const void* x, const float* gamma,
                       const float* mean, const float* rstd, const float* dres, float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta,
                       int64_t rows, int H, float p, uint64_t seed, uint32_t sid) {
#define ARGS grid, smem, st, dy, x, gamma, mean, rstd, dres, dx_f32, dx_bf16, dgamma, dbeta, rows, H, p, seed, sid
  if (db) {
    if (xb) launch_bwd1<true, true, NV>(ARGS); else launch_bwd1<true, false, NV>(ARGS);
  } else {
    if (xb) launch_bwd1<false, true, NV>(ARGS); else launch_bwd1<false, false, NV>(ARGS);
  }
#undef ARGS
}

}  // namespace evlm

extern "C" int evlm_layernorm_fwd(const void* x, int32_t x_dtype, const float* gamma, const float* beta, float eps, float* y_f32, void* y_bf16,
                                  float* mean, float* rstd, int64_t rows, int H, float dropout_p, uint64_t seed, uint32_t stream_id,
                                  void* stream) {
  using namespace evlm;
  if (!x || !gamma || !beta || (!y_f32 && !y_bf16) || rows < 0) return EVLM_EINVAL;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) |
       reinterpret_cast<uintptr_t>(y_f32) | reinterpret_cast<uintptr_t>(y_bf16)) & 15)
    return EVLM_EINVAL;  // 16-byte vector accesses
  if (H <= 0 || (H & 3) || H > LN_MAX_H) return EVLM_EUNSUPPORTED;
  if (rows == 0) return EVLM_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((rows + 3) / 4);
  const bool xb = x_dtype == EVLM_BF16;
  if (H <= 256) launch_fwd<2>(xb, grid, st, x, gamma, beta, eps, y_f32, y_bf16, mean, rstd, rows, H, dropout_p, seed, stream_id);
  else if (H <= 768) launch_fwd<6>(xb, grid, st, x, gamma, beta, eps, y_f32, y_bf16, mean, rstd, rows, H, dropout_p, seed, stream_id);
  else launch_fwd<12>(xb, grid, st, x, gamma, beta, eps, y_f32, y_bf16, mean, rstd, rows, H, dropout_p, seed, stream_id);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  EVLM_CUDA_RETURN();
}

static int layernorm_bwd_impl(const void* dy, int32_t dy_dtype, const void* x, int32_t x_dtype, const float* gamma, const float* mean,
                              const float* rstd, const float* dres, float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta,
                              int64_t rows, int H, float dropout_p, uint64_t seed, uint32_t stream_id, float out_dropout_p,
                              uint32_t out_stream_id, float* dcolsum, void* stream) {
  using namespace evlm;
  const bool extras = out_dropout_p > 0.f || dcolsum != nullptr;
  if (extras && (!dx_bf16 || (H % 128) != 0 || H > 1024)) return EVLM_EUNSUPPORTED;   // column-owner kernel only
  if (!dy || !x || !gamma || !mean || !rstd || (!dx_f32 && !dx_bf16) || rows < 0) return EVLM_EINVAL;
  if ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gamma) |
       reinterpret_cast<uintptr_t>(dres) | reinterpret_cast<uintptr_t>(dx_f32) | reinterpret_cast<uintptr_t>(dx_bf16)) & 15)
    return EVLM_EINVAL;  // 16-byte vector accesses
  if (H <= 0 || (H & 3) || H > LN_MAX_H) return EVLM_EUNSUPPORTED;
  if (rows == 0) return EVLM_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int64_t want = (rows + 3) / 4;
  const unsigned grid = (unsigned)(want < 148 * 8 ? want : 148 * 8);
  const size_t smem = (size_t)8 * H * sizeof(float);
  const bool db = dy_dtype == EVLM_BF16, xb = x_dtype == EVLM_BF16;
  if ((H % 128) == 0 && H <= 1024) {   // column-owner kernel: H/4 threads per block, 4 rows per iteration
    const int64_t groups = (rows + LNC_R - 1) / LNC_R;
    // resident blocks per SM at 80 registers and H / 4 threads per block (H = 768: 4 x 192 threads x 80 = 61 440 registers);
    // EVLM_LN_BWD_BLOCKS_PER_SM overrides (profiling knob)
    static const int per_sm_env = getenv("EVLM_LN_BWD_BLOCKS_PER_SM") ? atoi(getenv("EVLM_LN_BWD_BLOCKS_PER_SM")) : 0;
    const int per_sm = per_sm_env > 0 ? per_sm_env : (H <= 256 ? 8 : (H <= 512 ? 6 : (H <= 768 ? 4 : 3)));
    const unsigned g2 = (unsigned)(groups < 148 * per_sm ? groups : 148 * per_sm);
    const unsigned th = (unsigned)(H / 4);
#define CARGS dy, x, gamma, mean, rstd, dres, dx_f32, dx_bf16, dgamma, dbeta, rows, H, dropout_p, seed, stream_id, out_dropout_p, out_stream_id, dcolsum
    if (db) {
      if (xb) ln_bwd_cols_kernel<true, true><<<g2, th, 0, st>>>(CARGS); else ln_bwd_cols_kernel<true, false><<<g2, th, 0, st>>>(CARGS);
    } else {
      if (xb) ln_bwd_cols_kernel<false, true><<<g2, th, 0, st>>>(CARGS); else ln_bwd_cols_kernel<false, false><<<g2, th, 0, st>>>(CARGS);
    }
#undef CARGS
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    EVLM_CUDA_RETURN();
  }
#define ARGS db, xb, grid, smem, st, dy, x, gamma, mean, rstd, dres, dx_f32, dx_bf16, dgamma, dbeta, rows, H, dropout_p, seed, stream_id
  if (H <= 256) launch_bwd<2>(ARGS);
  else if (H <= 768) launch_bwd<6>(ARGS);
  else launch_bwd<12>(ARGS);
#undef ARGS
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  EVLM_CUDA_RETURN();
}

extern "C" int evlm_layernorm_bwd(const void* dy, int32_t dy_dtype, const void* x, int32_t x_dtype, const float* gamma, const float* mean,
                                  const float* rstd, const float* dres, float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta,
                                  int64_t rows, int H, float dropout_p, uint64_t seed, uint32_t stream_id, void* stream) {
  return layernorm_bwd_impl(dy, dy_dtype, x, x_dtype, gamma, mean, rstd, dres, dx_f32, dx_bf16, dgamma, dbeta, rows, H, dropout_p, seed,
                            stream_id, 0.f, 0u, nullptr, stream);
}
extern "C" int evlm_layernorm_bwd_ex(const void* dy, int32_t dy_dtype, const void* x, int32_t x_dtype, const float* gamma, const float* mean,
                                     const float* rstd, const float* dres, float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta,
                                     int64_t rows, int H, float dropout_p, uint64_t seed, uint32_t stream_id, float out_dropout_p,
                                     uint32_t out_stream_id, float* dcolsum, void* stream) {
  return layernorm_bwd_impl(dy, dy_dtype, x, x_dtype, gamma, mean, rstd, dres, dx_f32, dx_bf16, dgamma, dbeta, rows, H, dropout_p, seed,
                            stream_id, out_dropout_p, out_stream_id, dcolsum, stream);
}

// evlm_rng_bind() reaches the per-translation-unit seed-offset pointer through this hook (evlm_common.cuh).
namespace evlm { cudaError_t rng_bind_layernorm(const void* state_dev) { return tu_rng_bind(state_dev); } }
