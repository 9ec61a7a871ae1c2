// Single-query attention for the decode steps of the caption / answer generators (model_generation.py:233-300 through
// eff_bert.py:297-359 with a KV cache): Lq == 1, no attention map returned.
//
// One new token per sequence attends to Lk cached keys: per (item, head) that is 2 * Lk * 128 bytes of K / V against 128 bytes of q —
// a pure K/V STREAM (caption decoder, 577 image tokens: 57 MB per cross-attention layer and step).  The tcgen05 kernels spend a 128-row
// query tile, five key tiles and two sweeps on it (41 us per launch in profiles/r02_launch_summary_caption.txt); here
//   * a CTA of 4 warps owns one (item, head); 8 lanes share a key (16 bytes of the 128-byte K and V rows each), so a warp covers 4 keys
//     per step and the CTA 16, two steps in flight;
//   * every 8-lane group keeps an online softmax (running max, sum, 8 output dims per lane): ONE pass over K and V, no score buffer,
//     any Lk; the 16 groups are merged by shuffles and one shared-memory round.
// K / V rows of item b start at row kv_item(b) * kv_item_rows (evlm_attn_args.kv_item_rows: a pre-allocated KV cache holds more rows
// per item than are valid).  Bound: HBM (K/V bytes / duration).
#include "evlm_common.cuh"
#include "../../include/evlm.h"
#include <atomic>
#include <cstdlib>

namespace evlm {
extern std::atomic<unsigned long long> g_launch_count;

constexpr int DEC_WARPS = 4;

__device__ __forceinline__ float dot8_bf16(const uint4& a, const float (&q)[8]) {
  const float2 x0 = unpack_bf16x2(a.x), x1 = unpack_bf16x2(a.y), x2 = unpack_bf16x2(a.z), x3 = unpack_bf16x2(a.w);
  return x0.x * q[0] + x0.y * q[1] + x1.x * q[2] + x1.y * q[3] + x2.x * q[4] + x2.y * q[5] + x3.x * q[6] + x3.y * q[7];
}
// (m, l, acc) <- merge of two online-softmax states
__device__ __forceinline__ void merge_state(float& m, float& l, float (&acc)[8], float m2, float l2, const float (&acc2)[8]) {
  const float mn = fmaxf(m, m2);
  const float c1 = m == -INFINITY ? 0.f : __expf(m - mn), c2 = m2 == -INFINITY ? 0.f : __expf(m2 - mn);
  l = l * c1 + l2 * c2;
#pragma unroll
  for (int d = 0; d < 8; ++d) acc[d] = acc[d] * c1 + acc2[d] * c2;
  m = mn;
}

__global__ void __launch_bounds__(DEC_WARPS * 32) attn_fwd_decode_kernel(const evlm_attn_args a) {
  __shared__ float s_m[DEC_WARPS], s_l[DEC_WARPS], s_acc[DEC_WARPS][64];
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane & 7, kg = lane >> 3;
  const int kvb = a.kv_index ? __ldg(a.kv_index + b) : b;
  const int64_t kv_rows = a.kv_item_rows > 0 ? a.kv_item_rows : a.Lk;
  const __nv_bfloat16* K = reinterpret_cast<const __nv_bfloat16*>(a.k) + (int64_t)kvb * kv_rows * a.ldk + h * 64 + sub * 8;
  const __nv_bfloat16* V = reinterpret_cast<const __nv_bfloat16*>(a.v) + (int64_t)kvb * kv_rows * a.ldv + h * 64 + sub * 8;
  float q[8];
  {
    const uint4 qv = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.q) + (int64_t)b * a.ldq + h * 64 + sub * 8));
    const float2 x0 = unpack_bf16x2(qv.x), x1 = unpack_bf16x2(qv.y), x2 = unpack_bf16x2(qv.z), x3 = unpack_bf16x2(qv.w);
    q[0] = x0.x * a.scale; q[1] = x0.y * a.scale; q[2] = x1.x * a.scale; q[3] = x1.y * a.scale;
    q[4] = x2.x * a.scale; q[5] = x2.y * a.scale; q[6] = x3.x * a.scale; q[7] = x3.y * a.scale;
  }
  const float* mask = a.key_mask ? a.key_mask + (int64_t)b * a.Lk : nullptr;
  float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) acc[d] = 0.f;
  constexpr int STEP = DEC_WARPS * 4;
  auto absorb = [&](int j, const uint4& kx, const uint4& vx) {
    float s = dot8_bf16(kx, q);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (j < a.Lk) {      // (uniform inside the 8-lane group; the shuffles above are executed by the whole warp)
      if (mask) s += __ldg(mask + j);
      if (a.causal && j > a.causal_offset) s += -10000.0f;
      const float mn = fmaxf(m, s);
      const float c = m == -INFINITY ? 0.f : __expf(m - mn), p = __expf(s - mn);
      const float2 v0 = unpack_bf16x2(vx.x), v1 = unpack_bf16x2(vx.y), v2 = unpack_bf16x2(vx.z), v3 = unpack_bf16x2(vx.w);
      l = l * c + p;
      acc[0] = acc[0] * c + p * v0.x; acc[1] = acc[1] * c + p * v0.y; acc[2] = acc[2] * c + p * v1.x; acc[3] = acc[3] * c + p * v1.y;
      acc[4] = acc[4] * c + p * v2.x; acc[5] = acc[5] * c + p * v2.y; acc[6] = acc[6] * c + p * v3.x; acc[7] = acc[7] * c + p * v3.y;
      m = mn;
    }
  };
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  for (int jb = 0; jb < a.Lk; jb += 2 * STEP) {
    // (the whole CTA iterates together — the shuffles in absorb() need full warps; keys beyond Lk load nothing and are skipped)
    const int j0 = jb + warp * 4 + kg;
    const int j1 = j0 + STEP;
    const bool ok0 = j0 < a.Lk, ok1 = j1 < a.Lk;
    const uint4 k0 = ok0 ? __ldg(reinterpret_cast<const uint4*>(K + (int64_t)j0 * a.ldk)) : zero4;
    const uint4 v0 = ok0 ? __ldg(reinterpret_cast<const uint4*>(V + (int64_t)j0 * a.ldv)) : zero4;
    const uint4 k1 = ok1 ? __ldg(reinterpret_cast<const uint4*>(K + (int64_t)j1 * a.ldk)) : zero4;
    const uint4 v1 = ok1 ? __ldg(reinterpret_cast<const uint4*>(V + (int64_t)j1 * a.ldv)) : zero4;
    absorb(j0, k0, v0);
    absorb(j1, k1, v1);
  }
  // merge the 4 key groups of the warp (lanes with the same `sub` hold the same 8 output dims)
#pragma unroll
  for (int o = 8; o <= 16; o <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), l2 = __shfl_xor_sync(0xffffffffu, l, o);
    float acc2[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) acc2[d] = __shfl_xor_sync(0xffffffffu, acc[d], o);
    merge_state(m, l, acc, m2, l2, acc2);
  }
  if (kg == 0) {
    if (sub == 0) { s_m[warp] = m; s_l[warp] = l; }
#pragma unroll
    for (int d = 0; d < 8; ++d) s_acc[warp][sub * 8 + d] = acc[d];
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int d = threadIdx.x;
    float M = s_m[0];
#pragma unroll
    for (int w = 1; w < DEC_WARPS; ++w) M = fmaxf(M, s_m[w]);
    float L = 0.f, o = 0.f;
#pragma unroll
    for (int w = 0; w < DEC_WARPS; ++w) {
      const float c = s_m[w] == -INFINITY ? 0.f : __expf(s_m[w] - M);
      L += s_l[w] * c;
      o += s_acc[w][d] * c;
    }
    const float z = a.head_z ? __ldg(a.head_z + h) : 1.f;
    reinterpret_cast<__nv_bfloat16*>(a.ctx)[(int64_t)b * a.ldc + h * 64 + d] = __float2bfloat16(o / L * z);
    if (d == 0 && a.lse) a.lse[(int64_t)b * a.H + h] = M + __logf(L);
  }
}

// Lq == 1 without a returned map: EVLM_EUNSUPPORTED otherwise (the caller falls through to the tile kernels).
int attention_fwd_decode(const evlm_attn_args* a, cudaStream_t st) {
  static const bool off = getenv("EVLM_ATTN_NO_DECODE") != nullptr;   // profiling knob
  if (off || a->Lq != 1 || a->probs || a->full_mask || a->pack_items || a->dropout_p > 0.f) return EVLM_EUNSUPPORTED;
  attn_fwd_decode_kernel<<<dim3((unsigned)a->H, (unsigned)a->B), DEC_WARPS * 32, 0, st>>>(*a);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  EVLM_CUDA_RETURN();
}

}  // namespace evlm
