// Fused multi-head attention (head_dim 64) for the X-VLM encoders, forward and backward.
//
// Replaces eff_vit.py:141-197 (bmm / softmax / dropout / bmm / head_z) and eff_bert.py:297-359
// (matmul / scale / +mask / Softmax / Dropout / matmul / head_z), including the *materialised* attention
// probabilities the KD losses consume (GeneralDistill.py:62-69).
//
// KD-mode attention is HBM-bound on the fp32 P write, so this kernel keeps everything else
// on chip: one CTA per (batch, head, 64-query tile); K/V stream through shared memory in 64-key tiles;
// QK^T and PV run on the tensor cores (mma.sync m16n8k16 bf16, fp32 accumulate) with an online softmax;
// P is written exactly once, normalised, in a second sweep that only recomputes QK^T.
// The backward never reads P: it recomputes it from the saved row log-sum-exp, FlashAttention-2 style,
// one CTA per (batch, head), key tiles outer (dK/dV in registers), query tiles inner.
#include "evlm_common.cuh"
#include "../../include/evlm.h"
#include <atomic>
#include <cstdlib>

namespace evlm {
extern std::atomic<unsigned long long> g_launch_count;
int attention_fwd_decode(const evlm_attn_args* a, cudaStream_t st);   // attention_decode.cu (Lq == 1, no map)
int attention_fwd_tc(const evlm_attn_args* a, cudaStream_t st);   // attention_tc.cu
int attention_bwd_tc(const evlm_attn_args* a, cudaStream_t st);   // attention_tc_bwd.cu
int attention_fwd_tc_long(const evlm_attn_args* a, cudaStream_t st);   // attention_tc_long.cu (256 < Lk <= 1024)

constexpr int HD = 64;    // head dim
constexpr int TS = 64;    // tile size (queries / keys)
constexpr int LDS = 72;   // padded smem row stride in bf16 elements (144 B: conflict-free ldmatrix)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const __nv_bfloat16* p) {
  const uint32_t a = smem_u32(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const __nv_bfloat16* p) {
  const uint32_t a = smem_u32(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A fragment (16 rows x 16 k) of a row-major [rows][k] tile at (r0, k0).
__device__ __forceinline__ void load_a_frag(uint32_t (&a)[4], const __nv_bfloat16* s, int r0, int k0, int lane) {
  ldsm_x4(a, s + (r0 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + k0 + (lane >> 4) * 8);
}
// B fragments for two adjacent n-tiles (16 n x 16 k) where B[k][n] = X[n][k], X row-major [n][k]:
//   b[0],b[1] -> n-tile n0 ; b[2],b[3] -> n-tile n0+8
__device__ __forceinline__ void load_b_frag_nk(uint32_t (&b)[4], const __nv_bfloat16* s, int n0, int k0, int lane) {
  ldsm_x4(b, s + (n0 + (lane & 7) + (lane >> 4) * 8) * LDS + k0 + ((lane >> 3) & 1) * 8);
}
// B fragments for two adjacent n-tiles where B[k][n] = X[k][n], X row-major [k][n] (transposing load).
__device__ __forceinline__ void load_b_frag_kn(uint32_t (&b)[4], const __nv_bfloat16* s, int k0, int n0, int lane) {
  ldsm_x4_trans(b, s + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + n0 + (lane >> 4) * 8);
}

// Cooperative load of a [64 x 64] bf16 tile (rows row0.. of a matrix with `nrows` valid rows) into padded smem.
__device__ __forceinline__ void load_tile(__nv_bfloat16* s, const __nv_bfloat16* g, int64_t ld, int row0, int nrows) {
  for (int c = threadIdx.x; c < TS * 8; c += blockDim.x) {
    const int r = c >> 3, cc = c & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row0 + r < nrows) v = *reinterpret_cast<const uint4*>(g + (int64_t)(row0 + r) * ld + cc * 8);
    *reinterpret_cast<uint4*>(s + r * LDS + cc * 8) = v;
  }
}

struct MaskCtx {
  const float* key_mask;   // [Lk] for this batch or null
  const float* full_mask;  // [Lq, Lk] for this batch or null
  int causal, causal_offset, Lq, Lk;
  float scale;
};
// score in natural units (before softmax) for query i, key j, raw dot product s
__device__ __forceinline__ float masked_score(const MaskCtx& m, float s, int i, int j) {
  if (j >= m.Lk) return -INFINITY;
  float v = s * m.scale;
  if (m.key_mask) v += __ldg(m.key_mask + j);
  if (m.full_mask && i < m.Lq) v += __ldg(m.full_mask + (int64_t)i * m.Lk + j);
  if (m.causal && j > i + m.causal_offset) v += -10000.0f;
  return v;
}

// dropout stream index: rows padded to a multiple of 4 so that (j even, j+1) share one Philox call
__device__ __forceinline__ uint64_t drop_index(int b, int h, int H, int Lq, int Lk, int i, int j) {
  const uint64_t lkp = (uint64_t)((Lk + 3) & ~3);
  return (((uint64_t)b * H + h) * Lq + i) * lkp + j;
}

// =============================================================================================
// forward
// =============================================================================================
__global__ void __launch_bounds__(128) attn_fwd_kernel(const evlm_attn_args a) {
  __shared__ __align__(16) __nv_bfloat16 sQ[TS * LDS];
  __shared__ __align__(16) __nv_bfloat16 sK[TS * LDS];
  __shared__ __align__(16) __nv_bfloat16 sV[TS * LDS];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * TS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const __nv_bfloat16* qg = reinterpret_cast<const __nv_bfloat16*>(a.q) + (int64_t)b * a.Lq * a.ldq + h * HD;
  const int kvb = a.kv_index ? a.kv_index[b] : b;   // K/V batch item of this query item
  const __nv_bfloat16* kg = reinterpret_cast<const __nv_bfloat16*>(a.k) + (int64_t)kvb * a.Lk * a.ldk + h * HD;
  const __nv_bfloat16* vg = reinterpret_cast<const __nv_bfloat16*>(a.v) + (int64_t)kvb * a.Lk * a.ldv + h * HD;
  MaskCtx mc;
  mc.key_mask = a.key_mask ? a.key_mask + (int64_t)b * a.Lk : nullptr;
  mc.full_mask = a.full_mask ? a.full_mask + (int64_t)b * a.Lq * a.Lk : nullptr;
  mc.causal = a.causal; mc.causal_offset = a.causal_offset; mc.Lq = a.Lq; mc.Lk = a.Lk; mc.scale = a.scale;

  load_tile(sQ, qg, a.ldq, q0, a.Lq);
  __syncthreads();
  uint32_t qf[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) load_a_frag(qf[ks], sQ, warp * 16, ks * 16, lane);

  const int nkt = (a.Lk + TS - 1) / TS;
  const int row_lo = q0 + warp * 16 + g, row_hi = row_lo + 8;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  const float keep_inv = a.dropout_p > 0.f ? 1.f / (1.f - a.dropout_p) : 1.f;

  for (int kt = 0; kt < nkt; ++kt) {
    __syncthreads();
    load_tile(sK, kg, a.ldk, kt * TS, a.Lk);
    load_tile(sV, vg, a.ldv, kt * TS, a.Lk);
    __syncthreads();
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bf[4];
        load_b_frag_nk(bf, sK, np * 16, ks * 16, lane);
        mma16816(s[2 * np], qf[ks], bf[0], bf[1]);
        mma16816(s[2 * np + 1], qf[ks], bf[2], bf[3]);
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = kt * TS + nt * 8 + 2 * t + (c & 1);
        const int i = (c < 2) ? row_lo : row_hi;
        s[nt][c] = masked_score(mc, s[nt][c], i, j);
        mx[c >> 1] = fmaxf(mx[c >> 1], s[nt][c]);
      }
    }
    float corr[2], m_new[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      m_new[r] = fmaxf(m_run[r], mx[r]);
      corr[r] = (m_run[r] == -INFINITY) ? 0.f : exp2f((m_run[r] - m_new[r]) * LOG2E);
      m_run[r] = m_new[r];
      l_run[r] *= corr[r];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float p = (s[nt][c] == -INFINITY) ? 0.f : exp2f((s[nt][c] - m_new[c >> 1]) * LOG2E);
        l_run[c >> 1] += p;
        s[nt][c] = p;
      }
      if (a.dropout_p > 0.f) {
        const int j = kt * TS + nt * 8 + 2 * t;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int i = r ? row_hi : row_lo;
          const uint64_t e = drop_index(b, h, a.H, a.Lq, a.Lk, i, j);
          const float4 u = dropout_uniform4(a.dropout_seed + rng_offset(), a.dropout_stream, e >> 2);
          const float u0 = (e & 2) ? u.z : u.x, u1 = (e & 2) ? u.w : u.y;
          s[nt][2 * r] = u0 >= a.dropout_p ? s[nt][2 * r] * keep_inv : 0.f;
          s[nt][2 * r + 1] = u1 >= a.dropout_p ? s[nt][2 * r + 1] * keep_inv : 0.f;
        }
      }
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t bf[4];
        load_b_frag_kn(bf, sV, kk * 16, dp * 16, lane);
        mma16816(o[2 * dp], pa, bf[0], bf[1]);
        mma16816(o[2 * dp + 1], pa, bf[2], bf[3]);
      }
    }
  }
  float lse[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    lse[r] = m_run[r] + logf(l_run[r]);
  }
  const float z = a.head_z ? __ldg(a.head_z + h) : 1.f;
  __nv_bfloat16* cg = reinterpret_cast<__nv_bfloat16*>(a.ctx) + (int64_t)b * a.Lq * a.ldc + h * HD;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = r ? row_hi : row_lo;
    if (i < a.Lq) {
      const float sc = z / l_run[r];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
        *reinterpret_cast<uint32_t*>(cg + (int64_t)i * a.ldc + nt * 8 + 2 * t) = pack_bf16x2(o[nt][2 * r] * sc, o[nt][2 * r + 1] * sc);
      if (a.lse && t == 0) a.lse[((int64_t)b * a.H + h) * a.Lq + i] = lse[r];
    }
  }
  if (a.probs == nullptr) return;

  // ---- second sweep: write the normalised probabilities once ----
  float* pg = a.probs + ((int64_t)b * a.H + h) * a.Lq * (int64_t)(a.ldp ? a.ldp : a.Lk);
  for (int kt = 0; kt < nkt; ++kt) {
    __syncthreads();
    load_tile(sK, kg, a.ldk, kt * TS, a.Lk);
    __syncthreads();
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bf[4];
        load_b_frag_nk(bf, sK, np * 16, ks * 16, lane);
        mma16816(s[2 * np], qf[ks], bf[0], bf[1]);
        mma16816(s[2 * np + 1], qf[ks], bf[2], bf[3]);
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int j = kt * TS + nt * 8 + 2 * t;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int i = r ? row_hi : row_lo;
        if (i >= a.Lq || j >= a.Lk) continue;
        const float p0 = exp2f((masked_score(mc, s[nt][2 * r], i, j) - lse[r]) * LOG2E);
        float* dst = pg + (int64_t)i * (a.ldp ? a.ldp : a.Lk) + j;
        if (j + 1 < a.Lk) {
          const float p1 = exp2f((masked_score(mc, s[nt][2 * r + 1], i, j + 1) - lse[r]) * LOG2E);
          if ((reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
            *reinterpret_cast<float2*>(dst) = make_float2(p0, p1);
          } else {
            dst[0] = p0; dst[1] = p1;
          }
        } else {
          dst[0] = p0;
        }
      }
    }
  }
}

// =============================================================================================
// backward
// =============================================================================================
// delta[b,h,i] = sum_d dctx[b,i,h,d]*ctx[b,i,h,d] + sum_j dP_ext[b,h,i,j]*P[b,h,i,j]
// Eight lanes per row (16 bytes = 8 head dims of dctx and of ctx each), four rows per warp and pass, two passes in flight: the kernel is
// a latency-bound stream of 128-byte row segments, so what matters is loads in flight per lane (was: one warp per row, 4-byte loads).
constexpr int DELTA_ROWS_PER_WARP = 8;
__global__ void __launch_bounds__(256) attn_bwd_delta_kernel(const evlm_attn_args a, float* delta) {
  const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nrows = (int64_t)a.B * a.H * a.Lq;
  const float coef = a.dp_kd_coef ? __ldg(a.dp_kd_coef) : 1.f;
  int64_t rows[2];
  uint4 x[2], y[2];
  bool ok[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    rows[u] = warp * DELTA_ROWS_PER_WARP + u * 4 + grp;
    ok[u] = rows[u] < nrows;
    if (ok[u]) {
      const int i = (int)(rows[u] % a.Lq);
      const int h = (int)((rows[u] / a.Lq) % a.H);
      const int b = (int)(rows[u] / ((int64_t)a.Lq * a.H));
      x[u] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.dctx) + ((int64_t)b * a.Lq + i) * a.lddc + h * HD + sub * 8));
      y[u] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.ctx) + ((int64_t)b * a.Lq + i) * a.ldc + h * HD + sub * 8));
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    float acc = 0.f;
    if (ok[u]) {
      const float2 x0 = unpack_bf16x2(x[u].x), x1 = unpack_bf16x2(x[u].y), x2 = unpack_bf16x2(x[u].z), x3 = unpack_bf16x2(x[u].w);
      const float2 y0 = unpack_bf16x2(y[u].x), y1 = unpack_bf16x2(y[u].y), y2 = unpack_bf16x2(y[u].z), y3 = unpack_bf16x2(y[u].w);
      acc = x0.x * y0.x + x0.y * y0.y + x1.x * y1.x + x1.y * y1.y + x2.x * y2.x + x2.y * y2.y + x3.x * y3.x + x3.y * y3.y;
      if (a.dprobs_ext != nullptr && a.dp_rowdot != nullptr) {
        // supplied by the producer of dP (KD MSE backward, or — unscaled — the MSE forward when dP itself is formed in the attention backward)
        if (sub == 0) acc += a.dp_rowdot[rows[u]] * coef;
      } else if (a.dprobs_ext != nullptr) {
        const int64_t ldp = a.ldp ? a.ldp : a.Lk;
        const float* dp = a.dprobs_ext + rows[u] * ldp;
        const float* p = a.probs + rows[u] * ldp;
        for (int j = sub; j < a.Lk; j += 8) acc += dp[j] * p[j];
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (ok[u] && sub == 0) delta[rows[u]] = acc;
  }
}

struct BwdSmem {
  __nv_bfloat16 K[TS * LDS];
  __nv_bfloat16 V[TS * LDS];
  __nv_bfloat16 Q[TS * LDS];
  __nv_bfloat16 dO[TS * LDS];
  __nv_bfloat16 dS[TS * LDS];  // [q][key]
  float dPe[TS * (TS + 1)];    // [q][key] tile of the external dP (fp32), padded
  float lse[TS];
  float delta[TS];
  float red[32];
};

__global__ void __launch_bounds__(128) attn_bwd_kernel(const evlm_attn_args a, const float* __restrict__ delta_g, float* __restrict__ dq_acc) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const __nv_bfloat16* qg = reinterpret_cast<const __nv_bfloat16*>(a.q) + (int64_t)b * a.Lq * a.ldq + h * HD;
  const int kvb = a.kv_index ? a.kv_index[b] : b;   // K/V batch item of this query item
  const __nv_bfloat16* kg = reinterpret_cast<const __nv_bfloat16*>(a.k) + (int64_t)kvb * a.Lk * a.ldk + h * HD;
  const __nv_bfloat16* vg = reinterpret_cast<const __nv_bfloat16*>(a.v) + (int64_t)kvb * a.Lk * a.ldv + h * HD;
  const __nv_bfloat16* dog = reinterpret_cast<const __nv_bfloat16*>(a.dctx) + (int64_t)b * a.Lq * a.lddc + h * HD;
  const float* lse_g = a.lse + ((int64_t)b * a.H + h) * a.Lq;
  const float* dlt_g = delta_g + ((int64_t)b * a.H + h) * a.Lq;
  const float* dpe_g = a.dprobs_ext ? a.dprobs_ext + ((int64_t)b * a.H + h) * a.Lq * (int64_t)(a.ldp ? a.ldp : a.Lk) : nullptr;
  MaskCtx mc;
  mc.key_mask = a.key_mask ? a.key_mask + (int64_t)b * a.Lk : nullptr;
  mc.full_mask = a.full_mask ? a.full_mask + (int64_t)b * a.Lq * a.Lk : nullptr;
  mc.causal = a.causal; mc.causal_offset = a.causal_offset; mc.Lq = a.Lq; mc.Lk = a.Lk; mc.scale = a.scale;
  const float z = a.head_z ? __ldg(a.head_z + h) : 1.f;
  const float keep_inv = a.dropout_p > 0.f ? 1.f / (1.f - a.dropout_p) : 1.f;
  const int nkt = (a.Lk + TS - 1) / TS, nqt = (a.Lq + TS - 1) / TS;
  float dz_part = 0.f;

  for (int kt = 0; kt < nkt; ++kt) {
    __syncthreads();
    load_tile(sm.K, kg, a.ldk, kt * TS, a.Lk);
    load_tile(sm.V, vg, a.ldv, kt * TS, a.Lk);
    __syncthreads();
    uint32_t kf[4][4], vf[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      load_a_frag(kf[ks], sm.K, warp * 16, ks * 16, lane);
      load_a_frag(vf[ks], sm.V, warp * 16, ks * 16, lane);
    }
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
      dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    const int key_lo = kt * TS + warp * 16 + g, key_hi = key_lo + 8;

    for (int qt = 0; qt < nqt; ++qt) {
      const int q0 = qt * TS;
      __syncthreads();  // previous iteration's readers of Q / dO / dS / dPe are done
      load_tile(sm.Q, qg, a.ldq, q0, a.Lq);
      load_tile(sm.dO, dog, a.lddc, q0, a.Lq);
      if (threadIdx.x < TS) {
        const int i = q0 + threadIdx.x;
        sm.lse[threadIdx.x] = i < a.Lq ? lse_g[i] : 0.f;
        sm.delta[threadIdx.x] = i < a.Lq ? dlt_g[i] : 0.f;
      }
      if (dpe_g) {
        for (int c = threadIdx.x; c < TS * TS; c += blockDim.x) {
          const int r = c >> 6, cc = c & 63;
          const int i = q0 + r, j = kt * TS + cc;
          sm.dPe[r * (TS + 1) + cc] = (i < a.Lq && j < a.Lk) ? dpe_g[(int64_t)i * (a.ldp ? a.ldp : a.Lk) + j] : 0.f;
        }
      }
      __syncthreads();
      // S^T and G^T: [16 keys of this warp] x [64 queries]
      float st[8][4], gt[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
        gt[i][0] = gt[i][1] = gt[i][2] = gt[i][3] = 0.f;
      }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t bf[4];
          load_b_frag_nk(bf, sm.Q, np * 16, ks * 16, lane);
          mma16816(st[2 * np], kf[ks], bf[0], bf[1]);
          mma16816(st[2 * np + 1], kf[ks], bf[2], bf[3]);
          load_b_frag_nk(bf, sm.dO, np * 16, ks * 16, lane);
          mma16816(gt[2 * np], vf[ks], bf[0], bf[1]);
          mma16816(gt[2 * np + 1], vf[ks], bf[2], bf[3]);
        }
      }
      // elementwise: st -> dS^T (bf16 A operand for dK and smem dS), gt -> (D o P)^T (A operand for dV)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int ql = nt * 8 + 2 * t + (c & 1);  // query index within tile
          const int i = q0 + ql;
          const int j = (c < 2) ? key_lo : key_hi;
          float p = 0.f;
          if (i < a.Lq && j < a.Lk) p = exp2f((masked_score(mc, st[nt][c], i, j) - sm.lse[ql]) * LOG2E);
          float dmask = 1.f;
          if (a.dropout_p > 0.f) {
            const float u = dropout_uniform(a.dropout_seed + rng_offset(), a.dropout_stream, drop_index(b, h, a.H, a.Lq, a.Lk, i, j));
            dmask = u >= a.dropout_p ? keep_inv : 0.f;
          }
          const float gval = gt[nt][c];
          const float pd = p * dmask;
          dz_part += pd * gval;
          float dp = z * dmask * gval;
          if (dpe_g) dp += sm.dPe[ql * (TS + 1) + ((c < 2) ? (warp * 16 + g) : (warp * 16 + g + 8))];
          st[nt][c] = p * (dp - sm.delta[ql]);  // dS^T
          gt[nt][c] = pd;                        // (D o P)^T
        }
      }
      // dS to smem as [q][key] for the dQ product
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int ql = nt * 8 + 2 * t + (c & 1);
          const int kl = warp * 16 + g + ((c < 2) ? 0 : 8);
          sm.dS[ql * LDS + kl] = __float2bfloat16(st[nt][c]);
        }
      }
      // dV += (D o P)^T dO ; dK += dS^T Q          (k dimension = queries, 4 steps of 16)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t pa[4], sa[4];
        pa[0] = pack_bf16x2(gt[2 * kk][0], gt[2 * kk][1]);
        pa[1] = pack_bf16x2(gt[2 * kk][2], gt[2 * kk][3]);
        pa[2] = pack_bf16x2(gt[2 * kk + 1][0], gt[2 * kk + 1][1]);
        pa[3] = pack_bf16x2(gt[2 * kk + 1][2], gt[2 * kk + 1][3]);
        sa[0] = pack_bf16x2(st[2 * kk][0], st[2 * kk][1]);
        sa[1] = pack_bf16x2(st[2 * kk][2], st[2 * kk][3]);
        sa[2] = pack_bf16x2(st[2 * kk + 1][0], st[2 * kk + 1][1]);
        sa[3] = pack_bf16x2(st[2 * kk + 1][2], st[2 * kk + 1][3]);
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t bf[4];
          load_b_frag_kn(bf, sm.dO, kk * 16, dp * 16, lane);
          mma16816(dv[2 * dp], pa, bf[0], bf[1]);
          mma16816(dv[2 * dp + 1], pa, bf[2], bf[3]);
          load_b_frag_kn(bf, sm.Q, kk * 16, dp * 16, lane);
          mma16816(dk[2 * dp], sa, bf[0], bf[1]);
          mma16816(dk[2 * dp + 1], sa, bf[2], bf[3]);
        }
      }
      __syncthreads();  // dS tile complete
      // dQ[16 queries of this warp][64 d] = dS[q][keys] K[keys][d]
      float dq[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t af[4];
        load_a_frag(af, sm.dS, warp * 16, kk * 16, lane);
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t bf[4];
          load_b_frag_kn(bf, sm.K, kk * 16, dp * 16, lane);
          mma16816(dq[2 * dp], af, bf[0], bf[1]);
          mma16816(dq[2 * dp + 1], af, bf[2], bf[3]);
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int i = q0 + warp * 16 + g + 8 * r;
        if (i < a.Lq) {
          float* dst = dq_acc + ((int64_t)b * a.Lq + i) * (a.H * HD) + h * HD;
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) {
            float2* p2 = reinterpret_cast<float2*>(dst + nt * 8 + 2 * t);
            float2 v = make_float2(dq[nt][2 * r] * a.scale, dq[nt][2 * r + 1] * a.scale);
            if (kt > 0) {
              const float2 old = *p2;
              v.x += old.x; v.y += old.y;
            }
            *p2 = v;
          }
        }
      }
    }
    // write dK, dV for this warp's 16 keys
    __nv_bfloat16* dkg = reinterpret_cast<__nv_bfloat16*>(a.dk) + (int64_t)b * a.Lk * a.lddk + h * HD;
    __nv_bfloat16* dvg = reinterpret_cast<__nv_bfloat16*>(a.dv) + (int64_t)b * a.Lk * a.lddv + h * HD;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int j = r ? key_hi : key_lo;
      if (j < a.Lk) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          *reinterpret_cast<uint32_t*>(dkg + (int64_t)j * a.lddk + nt * 8 + 2 * t) =
              pack_bf16x2(dk[nt][2 * r] * a.scale, dk[nt][2 * r + 1] * a.scale);
          *reinterpret_cast<uint32_t*>(dvg + (int64_t)j * a.lddv + nt * 8 + 2 * t) = pack_bf16x2(dv[nt][2 * r] * z, dv[nt][2 * r + 1] * z);
        }
      }
    }
  }
  if (a.dhead_z != nullptr) {
    const float tot = block_sum(dz_part, sm.red);
    if (threadIdx.x == 0) atomicAdd(a.dhead_z + h, tot);
  }
}

// =============================================================================================
// forward for TINY sequences (Lq, Lk <= 16): answer decoders (4-token answers, single-token decode steps with a short cache).
// A tensor-core CTA per (item, head) would be almost all setup for a 4 x 4 problem and there are tens of thousands of them
// (VQA rank_answer: 24 x 128 candidates x 12 heads); here a WARP owns one (item, head): Q / K / V rows staged in shared memory
// (row pitch 66 bf16: conflict-free when every lane reads its own row), lane i computes query row i in fp32.
// =============================================================================================
constexpr int SM_L = 16;          // max rows
constexpr int SM_WARPS = 4;
// Lane layout (round 2): 8 lanes share a query row (16 bytes = 8 head dims each), 4 query rows per warp — every lane works even for
// 4-token answers (before, lane = query row left 28 of 32 lanes idle and staged Q / K / V through shared memory).  A warp owns one
// (item, head, group of 4 query rows); the <= 16 keys are walked with all K / V loads issued up front, scores are reduced over the
// 8 lanes by shuffles, the softmax runs on the register-resident score row.
template <int LKMAX>      // 4 / 8 / 16: the K / V rows of the problem live in registers
__global__ void __launch_bounds__(SM_WARPS * 32) attn_fwd_small_kernel(const evlm_attn_args a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane & 7, rsel = lane >> 3;
  const int rgroups = (a.Lq + 3) >> 2;
  const int64_t unit = (int64_t)blockIdx.x * SM_WARPS + warp;     // (item, head, row group)
  if (unit >= (int64_t)a.B * a.H * rgroups) return;
  const int rg = (int)(unit % rgroups);
  const int64_t pair = unit / rgroups;
  const int b = (int)(pair / a.H), h = (int)(pair % a.H);
  const int i = rg * 4 + rsel;
  const bool rowok = i < a.Lq;
  const __nv_bfloat16* qg = reinterpret_cast<const __nv_bfloat16*>(a.q) + ((int64_t)b * a.Lq + (rowok ? i : 0)) * a.ldq + h * HD + sub * 8;
  const __nv_bfloat16* kg = reinterpret_cast<const __nv_bfloat16*>(a.k) + (int64_t)b * a.Lk * a.ldk + h * HD + sub * 8;
  const __nv_bfloat16* vg = reinterpret_cast<const __nv_bfloat16*>(a.v) + (int64_t)b * a.Lk * a.ldv + h * HD + sub * 8;
  float q[8];
  {
    const uint4 qv = __ldg(reinterpret_cast<const uint4*>(qg));
    const float2 x0 = unpack_bf16x2(qv.x), x1 = unpack_bf16x2(qv.y), x2 = unpack_bf16x2(qv.z), x3 = unpack_bf16x2(qv.w);
    q[0] = x0.x; q[1] = x0.y; q[2] = x1.x; q[3] = x1.y; q[4] = x2.x; q[5] = x2.y; q[6] = x3.x; q[7] = x3.y;
  }
  MaskCtx mc{a.key_mask ? a.key_mask + (int64_t)b * a.Lk : nullptr, nullptr, a.causal, a.causal_offset, a.Lq, a.Lk, a.scale};
  uint4 kv[LKMAX], vv[LKMAX];
#pragma unroll
  for (int j = 0; j < LKMAX; ++j) {
    if (j < a.Lk) {
      kv[j] = __ldg(reinterpret_cast<const uint4*>(kg + (int64_t)j * a.ldk));
      vv[j] = __ldg(reinterpret_cast<const uint4*>(vg + (int64_t)j * a.ldv));
    }
  }
  float sc[LKMAX];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < LKMAX; ++j) {
    sc[j] = -INFINITY;
    if (j < a.Lk) {      // (warp-uniform)
      const float2 k0 = unpack_bf16x2(kv[j].x), k1 = unpack_bf16x2(kv[j].y), k2 = unpack_bf16x2(kv[j].z), k3 = unpack_bf16x2(kv[j].w);
      float acc = q[0] * k0.x + q[1] * k0.y + q[2] * k1.x + q[3] * k1.y + q[4] * k2.x + q[5] * k2.y + q[6] * k3.x + q[7] * k3.y;
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      sc[j] = masked_score(mc, acc, i, j);
      mx = fmaxf(mx, sc[j]);
    }
  }
  float l = 0.f;
#pragma unroll
  for (int j = 0; j < LKMAX; ++j) {
    sc[j] = j < a.Lk ? exp2f((sc[j] - mx) * LOG2E) : 0.f;
    l += sc[j];
  }
  if (!rowok) return;
  const float inv_l = 1.f / l;
  const int64_t row = ((int64_t)b * a.H + h) * a.Lq + i;
  if (sub == 0) {
    if (a.lse) a.lse[row] = mx + logf(l);
    if (a.probs) {
      float* pg = a.probs + row * (a.ldp ? a.ldp : a.Lk);
#pragma unroll
      for (int j = 0; j < LKMAX; ++j)
        if (j < a.Lk) pg[j] = sc[j] * inv_l;
    }
  }
  // the P V product sees bf16 probabilities, like the tensor-core kernels
  float o[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) o[d] = 0.f;
#pragma unroll
  for (int j = 0; j < LKMAX; ++j) {
    if (j < a.Lk) {
      const float pj = __bfloat162float(__float2bfloat16(sc[j]));
      const float2 v0 = unpack_bf16x2(vv[j].x), v1 = unpack_bf16x2(vv[j].y), v2 = unpack_bf16x2(vv[j].z), v3 = unpack_bf16x2(vv[j].w);
      o[0] = fmaf(pj, v0.x, o[0]); o[1] = fmaf(pj, v0.y, o[1]); o[2] = fmaf(pj, v1.x, o[2]); o[3] = fmaf(pj, v1.y, o[3]);
      o[4] = fmaf(pj, v2.x, o[4]); o[5] = fmaf(pj, v2.y, o[5]); o[6] = fmaf(pj, v3.x, o[6]); o[7] = fmaf(pj, v3.y, o[7]);
    }
  }
  const float z = (a.head_z ? __ldg(a.head_z + h) : 1.f) * inv_l;
  __nv_bfloat16* cg = reinterpret_cast<__nv_bfloat16*>(a.ctx) + ((int64_t)b * a.Lq + i) * a.ldc + h * HD + sub * 8;
  *reinterpret_cast<uint4*>(cg) = make_uint4(pack_bf16x2(o[0] * z, o[1] * z), pack_bf16x2(o[2] * z, o[3] * z), pack_bf16x2(o[4] * z, o[5] * z),
                                             pack_bf16x2(o[6] * z, o[7] * z));
}

// dq (bf16, strided) = dq_acc (fp32 [B*Lq, H*64])
__global__ void attn_dq_cast_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq, int64_t rows, int cols, int64_t ld) {
  const int64_t idx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (idx >= rows * cols) return;
  const int64_t r = idx / cols;
  const int c = (int)(idx % cols);
  const float4 v = *reinterpret_cast<const float4*>(acc + idx);
  uint2 o = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  *reinterpret_cast<uint2*>(dq + r * ld + c) = o;
}

static int check_common(const evlm_attn_args* a) {
  if (!a || !a->q || !a->k || !a->v) return EVLM_EINVAL;
  if (a->B <= 0 || a->H <= 0 || a->Lq <= 0 || a->Lk <= 0) return EVLM_EINVAL;
  if ((a->ldq % 8) || (a->ldk % 8) || (a->ldv % 8)) return EVLM_EINVAL;
  if ((reinterpret_cast<uintptr_t>(a->q) & 15) || (reinterpret_cast<uintptr_t>(a->k) & 15) || (reinterpret_cast<uintptr_t>(a->v) & 15))
    return EVLM_EINVAL;
  if (a->dropout_p < 0.f || a->dropout_p >= 1.f) return EVLM_EINVAL;
  if (a->kv_index && a->kv_batches <= 0) return EVLM_EINVAL;
  return 0;
}

}  // namespace evlm

extern "C" int evlm_attention_fwd(const evlm_attn_args* a, void* stream) {
  using namespace evlm;
  int rc = check_common(a);
  if (rc) return rc;
  if (!a->ctx || (a->ldc % 8) || (reinterpret_cast<uintptr_t>(a->ctx) & 15)) return EVLM_EINVAL;
  // single-query decode steps (Lq == 1, no map): K/V stream kernel
  rc = attention_fwd_decode(a, reinterpret_cast<cudaStream_t>(stream));
  if (rc != EVLM_EUNSUPPORTED) return rc;
  if (a->kv_item_rows && a->kv_item_rows != a->Lk) return EVLM_EUNSUPPORTED;   // only the decode kernel reads a KV cache with spare rows
  // tiny problems (answer decoders: 4-token answers, single-token decode steps): one warp per (item, head)
  static const bool no_small = getenv("EVLM_ATTN_NO_SMALL") != nullptr;          // profiling knob
  if (!no_small && a->Lq <= SM_L && a->Lk <= SM_L && !a->full_mask && !a->pack_items && !a->kv_index && a->dropout_p == 0.f &&
      (int64_t)a->B * a->H >= 1024) {
    const int64_t units = (int64_t)a->B * a->H * ((a->Lq + 3) / 4);     // a warp per (item, head, 4 query rows)
    if (a->probs && a->ldp > a->Lk)
      cudaMemsetAsync(a->probs, 0, (size_t)a->B * a->H * a->Lq * (size_t)a->ldp * sizeof(float), reinterpret_cast<cudaStream_t>(stream));
    const unsigned grid = (unsigned)((units + SM_WARPS - 1) / SM_WARPS);
    cudaStream_t st_small = reinterpret_cast<cudaStream_t>(stream);
    if (a->Lk <= 4) attn_fwd_small_kernel<4><<<grid, SM_WARPS * 32, 0, st_small>>>(*a);
    else if (a->Lk <= 8) attn_fwd_small_kernel<8><<<grid, SM_WARPS * 32, 0, st_small>>>(*a);
    else attn_fwd_small_kernel<16><<<grid, SM_WARPS * 32, 0, st_small>>>(*a);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    EVLM_CUDA_RETURN();
  }
  // key lengths <= 256 (ViT-224, BERT, text->image cross attention): tcgen05 / TMEM kernel; longer: tiled kernel below
  static const bool force_tiled = getenv("EVLM_ATTN_FORCE_TILED") != nullptr;   // profiling knob: bypass the tcgen05 kernels
  rc = force_tiled ? EVLM_EUNSUPPORTED : attention_fwd_tc(a, reinterpret_cast<cudaStream_t>(stream));
  if (rc != EVLM_EUNSUPPORTED) return rc;
  if (!force_tiled) {   // 256 < Lk <= 1024 (ViT-384 / ViT-480, question -> image cross attention): two-sweep tcgen05 kernel
    rc = attention_fwd_tc_long(a, reinterpret_cast<cudaStream_t>(stream));
    if (rc != EVLM_EUNSUPPORTED) return rc;
  }
  if (a->pack_items) return EVLM_EUNSUPPORTED;   // packed query items exist in the tcgen05 kernels only
  if (a->probs && a->ldp > a->Lk)   // the tiled kernel writes columns < Lk only: the pad columns of a pitched map must still be 0
    cudaMemsetAsync(a->probs, 0, (size_t)a->B * a->H * a->Lq * (size_t)a->ldp * sizeof(float), reinterpret_cast<cudaStream_t>(stream));
  dim3 grid((a->Lq + TS - 1) / TS, a->H, a->B);
  attn_fwd_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*a);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  EVLM_CUDA_RETURN();
}

extern "C" size_t evlm_attention_bwd_workspace(const evlm_attn_args* a) {
  // delta [B,H,Lq] fp32 + dq accumulator [B*Lq, H*64] fp32
  const size_t n_delta = (((size_t)a->B * a->H * a->Lq) + 3) & ~(size_t)3;   // keeps the dq accumulator 16-byte aligned
  return (n_delta + (size_t)a->B * a->Lq * a->H * 64) * sizeof(float);
}

extern "C" int evlm_attention_bwd(const evlm_attn_args* a, void* stream) {
  using namespace evlm;
  int rc = check_common(a);
  if (rc) return rc;
  if (!a->dctx || !a->ctx || !a->lse || !a->dq || !a->dk || !a->dv || !a->dkv_accum) return EVLM_EINVAL;
  if (a->kv_item_rows && a->kv_item_rows != a->Lk) return EVLM_EUNSUPPORTED;   // KV caches are inference-only
  if ((a->lddc % 8) || (a->ldc % 8) || (a->lddq % 4) || (a->lddk % 2) || (a->lddv % 2)) return EVLM_EINVAL;
  if ((reinterpret_cast<uintptr_t>(a->dctx) | reinterpret_cast<uintptr_t>(a->ctx)) & 15) return EVLM_EINVAL;   // 16-byte row segments
  if (a->dprobs_ext && !a->probs && !a->dp_kd_coef) return EVLM_EINVAL;
  if (a->dp_kd_coef && (!a->dprobs_ext || !a->dp_rowdot || a->dropout_p > 0.f)) return EVLM_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* delta = a->dkv_accum;
  float* dq_acc = delta + ((((size_t)a->B * a->H * a->Lq) + 3) & ~(size_t)3);
  if ((reinterpret_cast<uintptr_t>(dq_acc) & 15)) return EVLM_EINVAL;
  const int64_t nrows = (int64_t)a->B * a->H * a->Lq;
  attn_bwd_delta_kernel<<<(unsigned)((nrows + 8 * DELTA_ROWS_PER_WARP - 1) / (8 * DELTA_ROWS_PER_WARP)), 256, 0, st>>>(*a, delta);
  {  // Lq, Lk <= 256: tcgen05 / TMEM kernel writes dq / dk / dv directly
    static const bool force_tiled = getenv("EVLM_ATTN_FORCE_TILED") != nullptr;
    int rc_tc = force_tiled ? EVLM_EUNSUPPORTED : attention_bwd_tc(a, st);
    if (rc_tc == EVLM_EUNSUPPORTED && a->pack_items) return EVLM_EUNSUPPORTED;
    if (rc_tc == (1 << 30)) {   // LONG variant: dq was accumulated in fp32 (red.add) -> bf16
      const int64_t n = (int64_t)a->B * a->Lq * a->H * 64;
      attn_dq_cast_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(dq_acc, reinterpret_cast<__nv_bfloat16*>(a->dq),
                                                                          (int64_t)a->B * a->Lq, a->H * 64, a->lddq);
      g_launch_count.fetch_add(2, std::memory_order_relaxed);
      EVLM_CUDA_RETURN();
    }
    if (rc_tc != EVLM_EUNSUPPORTED) {
      g_launch_count.fetch_add(1, std::memory_order_relaxed);
      return rc_tc;
    }
  }
  if (a->dp_kd_coef) return EVLM_EUNSUPPORTED;   // the tiled kernel reads a materialised dP only
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem));
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  attn_bwd_kernel<<<a->B * a->H, 128, sizeof(BwdSmem), st>>>(*a, delta, dq_acc);
  const int64_t n = (int64_t)a->B * a->Lq * a->H * 64;
  attn_dq_cast_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(dq_acc, reinterpret_cast<__nv_bfloat16*>(a->dq), (int64_t)a->B * a->Lq,
                                                                      a->H * 64, a->lddq);
  g_launch_count.fetch_add(3, std::memory_order_relaxed);
  EVLM_CUDA_RETURN();
}

// evlm_rng_bind() reaches the per-translation-unit seed-offset pointer through this hook (evlm_common.cuh).
namespace evlm { cudaError_t rng_bind_attention(const void* state_dev) { return tu_rng_bind(state_dev); } }
