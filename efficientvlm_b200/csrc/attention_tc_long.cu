// tcgen05 / TMEM forward attention for LONG key sequences, 256 < Lk <= 1024: ViT-384 (577 tokens), ViT-480 (901 tokens) and the
// question -> image cross attention of the VQA / retrieval models (eff_vit.py:141-197, eff_bert.py:297-359).
//
// One CTA per (batch item, head, 128-query tile); the keys stream through shared memory in tiles of 128.  The row statistics
// are not known until every key has been seen, and KD mode has to write NORMALISED probabilities exactly once, so the CTA makes
// two sweeps over the key tiles (the score MMA is cheap next to the softmax and the fp32 P write):
//   sweep 1   S = Q K_j^T (tcgen05, fp32 in TMEM)  ->  running row max (and, in KD mode, the running row sum)
//   sweep 2   S = Q K_j^T again -> p = 2^(s*c + mask - max): (KD) p / sum -> global, coalesced through a per-warp transpose stage;
//             bf16(dropout(p)) -> 128B-swizzled smem tile -> O += P V_j (tcgen05, V used in place as an MN-major B operand)
//   epilogue  ctx = O * head_z / sum -> bf16 ; lse for the backward.
// Without a P output, sweep 1 only takes maxima (no exponentials) and sweep 2 accumulates the sum.
// Q stays in shared memory for the whole CTA; K_j / V_j are single-buffered TMA tiles, the next one is requested as soon as the
// MMA that read the buffer has completed.  ~103 KB smem and 256 TMEM columns per CTA: two CTAs per SM, so one CTA's softmax
// overlaps the other's loads and MMAs.  Every query row is shared by 2 threads (column groups), as in attention_tc.cu.
#include "evlm_common.cuh"
#include "evlm_tma.cuh"
#include "../../include/evlm.h"
#include <atomic>
#include <cstdlib>

namespace evlm {
extern std::atomic<unsigned long long> g_launch_count;

constexpr float TL_LOG2E = 1.4426950408889634f;
constexpr float TL_LN2 = 0.6931471805599453f;
constexpr int TL_KT = 128;                         // keys per tile
constexpr int TL_MAX_LK = 1024;
constexpr int TL_SPLIT = 2;                        // column groups per query row
constexpr int TL_SM_WARPS = 4 * TL_SPLIT;          // (TMEM lane quadrant) x (column group)
constexpr int TL_SM_THREADS = TL_SM_WARPS * 32;
constexpr int TL_THREADS = TL_SM_THREADS + 32;     // + control warp (TMA + MMA issue)
constexpr int TL_STAGE_LD = 17;
constexpr int TL_OFF_Q = 0, TL_OFF_K = 16384, TL_OFF_V = 32768, TL_OFF_P = 49152;     // Q | K | V tiles 16 KB each, P 32 KB
constexpr int TL_OFF_STAGE = TL_OFF_P + 32768;
constexpr int TL_STAGE_BYTES = TL_SM_WARPS * 32 * TL_STAGE_LD * 4;
constexpr int TL_OFF_RED = TL_OFF_STAGE + TL_STAGE_BYTES;
constexpr int TL_RED_BYTES = 2 * TL_SPLIT * 128 * 4;
constexpr int TL_OFF_MASK = TL_OFF_RED + TL_RED_BYTES;
constexpr int TL_OFF_BAR = TL_OFF_MASK + TL_MAX_LK * 4;
constexpr int TL_SMEM = TL_OFF_BAR + 64;
constexpr int TL_TMEM_COLS = 256;                  // S: columns [0,128), O: columns [128,192)

struct AttnTlParams {
  CUtensorMap tq, tk, tv;
  CUtensorMap tp32;   // probability map [B*H][Lq][ldp], boxes of 32 rows x 16 columns (valid when p_tma)
  int p_tma;
  evlm_attn_args a;
  int nt;   // key tiles
};

__device__ __forceinline__ void tl_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  tmem_ld_32x32b_x16(taddr, r);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void tl_named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__global__ void __launch_bounds__(TL_THREADS, 2) attn_fwd_tc_long_kernel(const __grid_constant__ AttnTlParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const evlm_attn_args& a = p.a;
  const uint32_t sbase = smem_u32(smem_raw);
  if ((sbase & 1023u) != 0) __trap();
  uint8_t* sptr = smem_raw;
  const uint32_t sQ = sbase + TL_OFF_Q, sK = sbase + TL_OFF_K, sV = sbase + TL_OFF_V, sP = sbase + TL_OFF_P;
  float* stage = reinterpret_cast<float*>(sptr + TL_OFF_STAGE);
  float* red_a = reinterpret_cast<float*>(sptr + TL_OFF_RED);      // [group][row]
  float* red_b = red_a + TL_SPLIT * 128;
  float* smask = reinterpret_cast<float*>(sptr + TL_OFF_MASK);
  const uint32_t bar0 = sbase + TL_OFF_BAR;
  const uint32_t bar_k = bar0, bar_v = bar0 + 8, bar_s = bar0 + 16, bar_p = bar0 + 24, bar_o = bar0 + 32;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(sptr + TL_OFF_BAR + 48);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * 128;
  const int nt = p.nt;
  const int total_it = 2 * nt;

  if (warp == TL_SM_WARPS) {
    if (lane == 0) {
      tma_prefetch_desc(&p.tq);
      tma_prefetch_desc(&p.tk);
      tma_prefetch_desc(&p.tv);
      mbar_init(bar_k, 1);
      mbar_init(bar_v, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, TL_SM_THREADS);
      mbar_init(bar_o, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32((const void*)tmem_ptr_smem), (uint32_t)TL_TMEM_COLS);
    tmem_relinquish();
  }
  // additive key mask in log2 units; keys beyond Lk (tile padding / the next item's rows) are excluded with -inf
  for (int j = threadIdx.x; j < nt * TL_KT; j += TL_THREADS)
    smask[j] = j < a.Lk ? (a.key_mask ? a.key_mask[(int64_t)b * a.Lk + j] * TL_LOG2E : 0.f) : -INFINITY;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_smem;
  const uint32_t tmem_s = tmem, tmem_o = tmem + 128;

  if (warp == TL_SM_WARPS) {
    if (lane == 0) {
      const int kvb = a.kv_index ? __ldg(a.kv_index + b) : b;   // K/V batch item of this query item
      const int kv_row0 = kvb * a.Lk;
      mbar_expect_tx(bar_k, 2 * 16384);
      tma_load_2d(sQ, &p.tq, h * 64, b * a.Lq + q0, bar_k);
      tma_load_2d(sK, &p.tk, h * 64, kv_row0, bar_k);
      mbar_expect_tx(bar_v, 16384);
      tma_load_2d(sV, &p.tv, h * 64, kv_row0, bar_v);
      const uint32_t idesc_o = make_idesc_bf16(128, 64, false, true);
      // S = Q K_j^T of iteration `it` (tile j of sweep 1 or 2); its K tile must have landed
      auto issue_s = [&](int it) {
        const int j = it >= nt ? it - nt : it;
        const int Np = (min(TL_KT, a.Lk - j * TL_KT) + 15) & ~15;
        mbar_wait(bar_k, it & 1);
        tc_fence_after();
        const uint32_t idesc_s = make_idesc_bf16(128, Np, false, false);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_s, make_desc_kmajor(sQ + k * 32), make_desc_kmajor(sK + k * 32), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(bar_s);
      };
      issue_s(0);
      for (int it = 0; it < total_it; ++it) {
        const bool pass2 = it >= nt;
        const int j = pass2 ? it - nt : it;
        const int Np = (min(TL_KT, a.Lk - j * TL_KT) + 15) & ~15;
        // the K buffer is free once S(it) has completed: request the tile of iteration it + 1 (sweep 2 starts over at tile 0)
        mbar_wait(bar_s, it & 1);
        if (it + 1 < total_it) {
          const int jn = (it + 1 >= nt) ? it + 1 - nt : it + 1;
          mbar_expect_tx(bar_k, 16384);
          tma_load_2d(sK, &p.tk, h * 64, kv_row0 + jn * TL_KT, bar_k);
        }
        // ---- softmax warps: done with S (sweep 2: the P tile is in shared memory) ----
        mbar_wait(bar_p, it & 1);
        tc_fence_after();
        if (pass2) {
          mbar_wait(bar_v, j & 1);
          tc_fence_after();
          const int ksteps = Np >> 4;
          for (int k = 0; k < ksteps; ++k)
            umma_bf16(tmem_o, make_desc_kmajor(sP + (k >> 2) * 16384 + (k & 3) * 32), make_desc_mnmajor(sV + k * 2048), idesc_o,
                      (j > 0 || k > 0) ? 1u : 0u);
          umma_commit(bar_o);
        }
        // the S columns are free (bar_p): the next tile's score MMA goes in right behind the P V MMAs instead of waiting for them
        if (it + 1 < total_it) issue_s(it + 1);
        if (pass2) {
          mbar_wait(bar_o, j & 1);       // V and P buffers free
          if (j + 1 < nt) {
            mbar_expect_tx(bar_v, 16384);
            tma_load_2d(sV, &p.tv, h * 64, kv_row0 + (j + 1) * TL_KT, bar_v);
          }
        }
      }
    }
  } else {
    const int quad = warp & 3, grp = warp >> 2;
    const int r = quad * 32 + lane;
    const int qi = q0 + r;
    const bool row_valid = qi < a.Lq;
    const int warp_rows = min(32, a.Lq - q0 - quad * 32);   // valid query rows of this warp (<= 0: none)
    const bool dead = warp_rows <= 0;                        // (warp-uniform) only the barrier protocol matters
    const uint32_t trow = tmem_s + ((uint32_t)(quad * 32) << 16);
    const uint32_t trow_o = tmem_o + ((uint32_t)(quad * 32) << 16);
    const float sc2 = a.scale * TL_LOG2E;
    const bool want_probs = a.probs != nullptr;
    const int64_t ldp = a.ldp ? a.ldp : a.Lk;   // row pitch of the probability map
    const float keep_inv = a.dropout_p > 0.f ? 1.f / (1.f - a.dropout_p) : 1.f;
    const uint64_t seed = a.dropout_seed + rng_offset();
    const uint64_t lkp4 = (uint64_t)((a.Lk + 3) & ~3);
    const uint64_t erow = (((uint64_t)b * a.H + h) * a.Lq + (row_valid ? qi : 0)) * lkp4;
    uint8_t* prow = sptr + TL_OFF_P + r * 128;               // row r inside each 16 KB atom of the P tile
    float* st = stage + warp * 32 * TL_STAGE_LD;
    const int cj = lane & 15, rh = lane >> 4;
    const int64_t grow0 = ((int64_t)b * a.H + h) * a.Lq + q0 + quad * 32;   // global P row of this warp's first tile row
    float m2 = -1e30f, l = 0.f, inv_l = 1.f;

    for (int it = 0; it < total_it; ++it) {
      const bool pass2 = it >= nt;
      const int j = pass2 ? it - nt : it;
      const int nkeys = min(TL_KT, a.Lk - j * TL_KT);
      const int n16 = (nkeys + 15) >> 4;
      const float* mrow = smask + j * TL_KT;
      if (it == nt) {
        // ---- between the sweeps: the two column groups of a row agree on the row max (KD mode: and on the row sum) ----
        red_a[grp * 128 + r] = m2;
        red_b[grp * 128 + r] = l;
        tl_named_bar(1 + quad, 32 * TL_SPLIT);
        float mm = -1e30f;
#pragma unroll
        for (int g2 = 0; g2 < TL_SPLIT; ++g2) mm = fmaxf(mm, red_a[g2 * 128 + r]);
        if (want_probs) {
          float ll = 0.f;
#pragma unroll
          for (int g2 = 0; g2 < TL_SPLIT; ++g2) ll += red_b[g2 * 128 + r] * fast_ex2(red_a[g2 * 128 + r] - mm);
          l = ll;
          inv_l = 1.f / ll;
        } else {
          l = 0.f;
        }
        m2 = mm;
      }
      mbar_wait(bar_s, it & 1);
      tc_fence_after();
      if (!pass2) {
        if (!dead) {
          for (int cc = grp; cc < n16; cc += TL_SPLIT) {
            float v[16];
            tl_ld16(trow + cc * 16, v);
            float cm = -1e30f;
#pragma unroll
            for (int j4 = 0; j4 < 16; j4 += 4) {
              const float4 m4 = *reinterpret_cast<const float4*>(mrow + cc * 16 + j4);
              v[j4] = fmaf(v[j4], sc2, m4.x);
              v[j4 + 1] = fmaf(v[j4 + 1], sc2, m4.y);
              v[j4 + 2] = fmaf(v[j4 + 2], sc2, m4.z);
              v[j4 + 3] = fmaf(v[j4 + 3], sc2, m4.w);
              cm = fmaxf(cm, fmaxf(fmaxf(v[j4], v[j4 + 1]), fmaxf(v[j4 + 2], v[j4 + 3])));
            }
            if (want_probs) {
              const float mn = fmaxf(m2, cm);
              float s = 0.f;
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) s += fast_ex2(v[jj] - mn);
              l = l * fast_ex2(m2 - mn) + s;
              m2 = mn;
            } else {
              m2 = fmaxf(m2, cm);
            }
          }
        }
        tc_fence_before();
        mbar_arrive(bar_p);
      } else {
        if (j > 0) mbar_wait(bar_o, (j - 1) & 1);      // O += P V of the previous tile has consumed the P tile
        if (!dead) {
          for (int cc = grp; cc < n16; cc += TL_SPLIT) {
            float v[16];
            tl_ld16(trow + cc * 16, v);
#pragma unroll
            for (int j4 = 0; j4 < 16; j4 += 4) {
              const float4 m4 = *reinterpret_cast<const float4*>(mrow + cc * 16 + j4);
              v[j4] = fast_ex2(fmaf(v[j4], sc2, m4.x) - m2);
              v[j4 + 1] = fast_ex2(fmaf(v[j4 + 1], sc2, m4.y) - m2);
              v[j4 + 2] = fast_ex2(fmaf(v[j4 + 2], sc2, m4.z) - m2);
              v[j4 + 3] = fast_ex2(fmaf(v[j4 + 3], sc2, m4.w) - m2);
            }
            if (want_probs && p.p_tma) {
              // normalised probabilities leave as one TMA box store per warp and chunk (see attention_tc.cu): the thread parks its
              // row in the warp's swizzled [32][16] stage, lane 0 hands the box to the TMA unit; rows >= Lq / columns >= ldp are clipped
              const int col0 = j * TL_KT + cc * 16;
              if (col0 < (int)ldp) {
                float* st2 = stage + warp * 512;
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous box has left the stage
                __syncwarp();
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4)
                  *reinterpret_cast<float4*>(st2 + lane * 16 + ((q4 ^ ((lane >> 1) & 3)) << 2)) =
                      make_float4(v[4 * q4] * inv_l, v[4 * q4 + 1] * inv_l, v[4 * q4 + 2] * inv_l, v[4 * q4 + 3] * inv_l);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                                   reinterpret_cast<uint64_t>(&p.tp32)),
                               "r"(smem_u32(st2)), "r"(col0), "r"(q0 + quad * 32), "r"(b * a.H + h)
                               : "memory");
                  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
              }
            } else if (want_probs) {
              // normalised probabilities -> global through the per-warp transpose stage: one store instruction covers
              // 2 rows x 16 keys (64 contiguous bytes each)
              __syncwarp();
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) st[lane * TL_STAGE_LD + jj] = v[jj] * inv_l;
              __syncwarp();
              const int col = j * TL_KT + cc * 16 + cj;
              if (col < ldp) {                                       // (pad columns Lk .. ldp-1 receive their exact zeros)
                float* pb = a.probs + (grow0 + rh) * ldp + col;     // rows 2u + rh: fixed 32-bit offsets from the first one
                const float* sb = st + rh * TL_STAGE_LD + cj;
                const int step2 = 2 * (int)ldp;
                if (warp_rows >= 32) {                               // (warp-uniform) full warp: branch-free, fully unrolled
#pragma unroll
                  for (int u = 0; u < 16; ++u) pb[u * step2] = sb[u * 2 * TL_STAGE_LD];
                } else {
#pragma unroll 4
                  for (int u = 0; u < 16; ++u)
                    if (2 * u + rh < warp_rows) pb[u * step2] = sb[u * 2 * TL_STAGE_LD];
                }
              }
            } else {
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) l += v[jj];
            }
            if (a.dropout_p > 0.f) {
#pragma unroll
              for (int jj = 0; jj < 16; jj += 4) {   // one Philox call per 4 consecutive keys (row stride padded to a multiple of 4)
                const float4 u = dropout_uniform4(seed, a.dropout_stream, (erow + (uint64_t)(j * TL_KT + cc * 16 + jj)) >> 2);
                v[jj] = u.x >= a.dropout_p ? v[jj] * keep_inv : 0.f;
                v[jj + 1] = u.y >= a.dropout_p ? v[jj + 1] * keep_inv : 0.f;
                v[jj + 2] = u.z >= a.dropout_p ? v[jj + 2] * keep_inv : 0.f;
                v[jj + 3] = u.w >= a.dropout_p ? v[jj + 3] * keep_inv : 0.f;
              }
            }
            // bf16, 128B-swizzled K-major tile: atom = 64 keys; 16-byte chunk index XOR (row % 8)
#pragma unroll
            for (int j8 = 0; j8 < 2; ++j8) {
              const int key = cc * 16 + j8 * 8;
              const int atom = key >> 6, chunk = (key & 63) >> 3;
              uint4 o = make_uint4(pack_bf16x2(v[j8 * 8], v[j8 * 8 + 1]), pack_bf16x2(v[j8 * 8 + 2], v[j8 * 8 + 3]),
                                   pack_bf16x2(v[j8 * 8 + 4], v[j8 * 8 + 5]), pack_bf16x2(v[j8 * 8 + 6], v[j8 * 8 + 7]));
              *reinterpret_cast<uint4*>(prow + atom * 16384 + ((chunk ^ (r & 7)) << 4)) = o;
            }
          }
        }
        // hand the P tile to the tensor core: generic-proxy smem writes -> async proxy, TMEM reads done
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(bar_p);
      }
    }
    if (!want_probs) {
      // the row sum was accumulated in sweep 2: add the two column groups (red_b is idle since the exchange between the sweeps)
      red_b[grp * 128 + r] = l;
      tl_named_bar(1 + quad, 32 * TL_SPLIT);
      l = 0.f;
#pragma unroll
      for (int g2 = 0; g2 < TL_SPLIT; ++g2) l += red_b[g2 * 128 + r];
      inv_l = 1.f / l;
    }
    // ---- epilogue: O -> ctx; thread (r, grp) stores 32 columns of its row ----
    mbar_wait(bar_o, (nt - 1) & 1);
    tc_fence_after();
    if (!dead) {
      const float z = a.head_z ? __ldg(a.head_z + h) : 1.f;
      const float osc = z * inv_l;
      constexpr int OC = 64 / TL_SPLIT;
      __nv_bfloat16* cg = reinterpret_cast<__nv_bfloat16*>(a.ctx) + ((int64_t)b * a.Lq + (row_valid ? qi : 0)) * a.ldc + h * 64 + grp * OC;
#pragma unroll
      for (int c = 0; c < OC / 16; ++c) {
        float v[16];
        tl_ld16(trow_o + grp * OC + c * 16, v);
        if (row_valid) {
#pragma unroll
          for (int jj = 0; jj < 16; jj += 8) {
            uint4 o = make_uint4(pack_bf16x2(v[jj] * osc, v[jj + 1] * osc), pack_bf16x2(v[jj + 2] * osc, v[jj + 3] * osc),
                                 pack_bf16x2(v[jj + 4] * osc, v[jj + 5] * osc), pack_bf16x2(v[jj + 6] * osc, v[jj + 7] * osc));
            *reinterpret_cast<uint4*>(cg + c * 16 + jj) = o;
          }
        }
      }
      if (grp == 0 && row_valid && a.lse) a.lse[((int64_t)b * a.H + h) * a.Lq + qi] = (m2 + log2f(l)) * TL_LN2;
    }
  }
  if (p.p_tma && warp < TL_SM_WARPS) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // bulk stores complete before the CTA exits
  tc_fence_before();
  __syncthreads();
  if (warp == TL_SM_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem, (uint32_t)TL_TMEM_COLS);
  }
}

// Returns EVLM_EUNSUPPORTED when the shape is outside this kernel's envelope (the caller then uses the tiled mma.sync kernel).
int attention_fwd_tc_long(const evlm_attn_args* a, cudaStream_t st) {
  static const bool disabled = getenv("EVLM_ATTN_NO_LONG") != nullptr;    // profiling knob: A/B against the tiled kernel
  if (disabled || a->Lk <= 256 || a->Lk > TL_MAX_LK || a->full_mask != nullptr || a->causal || a->pack_items) return EVLM_EUNSUPPORTED;
  if ((a->ldc % 8) || (reinterpret_cast<uintptr_t>(a->ctx) & 15)) return EVLM_EUNSUPPORTED;
  AttnTlParams p;
  p.a = *a;
  p.nt = (a->Lk + TL_KT - 1) / TL_KT;
  int rc = make_tmap_bf16(&p.tq, a->q, (int64_t)a->B * a->Lq, (int64_t)a->H * 64, a->ldq, 128);
  if (rc) return rc;
  const int64_t kv_items = a->kv_index ? a->kv_batches : a->B;
  rc = make_tmap_bf16(&p.tk, a->k, kv_items * a->Lk, (int64_t)a->H * 64, a->ldk, TL_KT);
  if (rc) return rc;
  rc = make_tmap_bf16(&p.tv, a->v, kv_items * a->Lk, (int64_t)a->H * 64, a->ldv, TL_KT);
  if (rc) return rc;
  p.p_tma = 0;
  {
    static const bool no_tma_p = getenv("EVLM_ATTN_NO_TMA_P") != nullptr;
    const int64_t ldp = a->ldp ? a->ldp : a->Lk;
    if (a->probs && !no_tma_p && (ldp % 4) == 0 && (reinterpret_cast<uintptr_t>(a->probs) & 15) == 0)
      p.p_tma = make_tmap_probs(&p.tp32, a->probs, ldp, a->Lq, (int64_t)a->B * a->H, 32) == 0 ? 1 : 0;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TL_SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid((a->Lq + 127) / 128, a->H, a->B);
  attn_fwd_tc_long_kernel<<<grid, TL_THREADS, TL_SMEM, st>>>(p);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? EVLM_OK : (int)e;
}

cudaError_t rng_bind_attention_tc_long(const void* state_dev) { return tu_rng_bind(state_dev); }

}  // namespace evlm
