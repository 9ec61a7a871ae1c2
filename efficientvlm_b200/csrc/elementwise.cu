// HBM-bound layout / elementwise kernels: dtype casts (with dropout-mask replay), column reductions
// (bias / gate gradients), ViT patchify + token assembly, BERT embedding gather / scatter.
#include "evlm_common.cuh"
#include "../../include/evlm.h"
#include <atomic>

namespace evlm {
extern std::atomic<unsigned long long> g_launch_count;

// ------------------------------------------------------------------------------------------------ casts
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, int64_t lds, __nv_bfloat16* __restrict__ dst, int64_t ldd, int64_t rows,
                                     int64_t cols, float p, uint64_t seed, uint32_t sid) {
  const int64_t c4 = (cols + 3) >> 2;
  const float keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  const bool vec = ((lds & 3) == 0) && ((ldd & 3) == 0) && ((cols & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(dst) & 7) == 0);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < rows * c4; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / c4, c = (idx % c4) * 4;
    float v[4];
    if (vec) {
      const float4 q = *reinterpret_cast<const float4*>(src + r * lds + c);
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (c + j < cols) ? src[r * lds + c + j] : 0.f;
    }
    if (p > 0.f) {
      // dropout stream index = r * cols + c  (the GEMM epilogue's convention)
      if ((cols & 3) == 0) {
        const float4 u = dropout_uniform4(seed + rng_offset(), sid, (uint64_t)(r * cols + c) >> 2);
        v[0] = u.x >= p ? v[0] * keep : 0.f;
        v[1] = u.y >= p ? v[1] * keep : 0.f;
        v[2] = u.z >= p ? v[2] * keep : 0.f;
        v[3] = u.w >= p ? v[3] * keep : 0.f;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float u = dropout_uniform(seed + rng_offset(), sid, (uint64_t)(r * cols + c + j));
          v[j] = u >= p ? v[j] * keep : 0.f;
        }
      }
    }
    if (vec) {
      *reinterpret_cast<uint2*>(dst + r * ldd + c) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c + j < cols) dst[r * ldd + c + j] = __float2bfloat16(v[j]);
    }
  }
}

// Contiguous fast path (lds == ldd == cols, 16-byte aligned, n % 8 == 0): no row / column arithmetic (the generic kernel pays a
// 64-bit division per 4 elements), 8 elements per thread and iteration (two 16-byte loads, one 16-byte store), 4 iterations in flight.
__global__ void __launch_bounds__(256) cast_f32_bf16_flat_kernel(const float4* __restrict__ src, uint4* __restrict__ dst, int64_t n8, float p,
                                                                 uint64_t seed, uint32_t sid) {
  const float keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  auto convert = [&](int64_t i, const float4& a, const float4& b) {
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (p > 0.f) {   // dropout stream index = flat element index (== r * cols + c of the generic kernel), one Philox call per 4
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 u = dropout_uniform4(seed + rng_offset(), sid, (uint64_t)(2 * i + h));
        v[4 * h] = u.x >= p ? v[4 * h] * keep : 0.f;
        v[4 * h + 1] = u.y >= p ? v[4 * h + 1] * keep : 0.f;
        v[4 * h + 2] = u.z >= p ? v[4 * h + 2] * keep : 0.f;
        v[4 * h + 3] = u.w >= p ? v[4 * h + 3] * keep : 0.f;
      }
    }
    dst[i] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  };
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n8; i += 4 * stride) {
    float4 a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      a[u] = src[2 * (i + u * stride)];
      b[u] = src[2 * (i + u * stride) + 1];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) convert(i + u * stride, a[u], b[u]);
  }
  for (; i < n8; i += stride) convert(i, src[2 * i], src[2 * i + 1]);
}

__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd, int64_t rows,
                                     int64_t cols) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < rows * cols; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / cols, c = idx % cols;
    dst[r * ldd + c] = __bfloat162float(src[r * lds + c]);
  }
}

// ------------------------------------------------------------------------------------------------ column sums
// out[n] (+)= sum_m X[m,n].  Block = 32 x 8 threads; each block owns 32*VEC columns and a slab of rows.
template <bool BF16>
__global__ void colsum_kernel(const void* __restrict__ X, int64_t ldx, int64_t rows, int64_t cols, float* __restrict__ out, int rows_per_block) {
  __shared__ float sm[8][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t col = (int64_t)blockIdx.x * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  float acc = 0.f;
  if (col < cols) {
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      acc += BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(X)[r * ldx + col])
                  : reinterpret_cast<const float*>(X)[r * ldx + col];
    }
  }
  sm[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && col < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][tx];
    atomicAdd(out + col, t);
  }
}

// bf16 fast path: each thread reads 8 consecutive columns (16 bytes) of a row, a block covers 256 columns x a slab of rows.
__global__ void __launch_bounds__(256) colsum_bf16x8_kernel(const __nv_bfloat16* __restrict__ X, int64_t ldx, int64_t rows, int64_t cols,
                                                            float* __restrict__ out, int rows_per_block) {
  __shared__ float sm[8][256 + 8];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t col = (int64_t)blockIdx.x * 256 + tx * 8;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < cols) {   // cols % 8 == 0 on this path
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      const uint4 q = *reinterpret_cast<const uint4*>(X + r * ldx + col);
      const float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), c = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
      acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y; acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[ty][tx * 8 + j] = acc[j];
  __syncthreads();
  const int c = threadIdx.x;
  if ((int64_t)blockIdx.x * 256 + c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][c];
    atomicAdd(out + (int64_t)blockIdx.x * 256 + c, t);
  }
}

__global__ void coldot_kernel(const __nv_bfloat16* __restrict__ X, const __nv_bfloat16* __restrict__ Y, int64_t ld, int64_t rows, int64_t cols,
                              float* __restrict__ out, int rows_per_block) {
  __shared__ float sm[8][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t col = (int64_t)blockIdx.x * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  float acc = 0.f;
  if (col < cols)
    for (int64_t r = r0 + ty; r < r1; r += 8) acc += __bfloat162float(X[r * ld + col]) * __bfloat162float(Y[r * ld + col]);
  sm[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && col < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][tx];
    atomicAdd(out + col, t);
  }
}

// ------------------------------------------------------------------------------------------------ ViT patchify / assemble
// patches[(b*G*G + gy*G + gx), (c*P + ky)*P + kx] = image[b, c, gy*P + ky, gx*P + kx]
__global__ void im2col_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int C, int R, int P) {
  const int G = R / P;
  const int64_t K = (int64_t)C * P * P;
  const int64_t total4 = (int64_t)B * G * G * K / 4;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total4; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = idx * 4;
    const int64_t row = e / K;
    const int k = (int)(e % K);
    const int kx = k % P, ky = (k / P) % P, c = k / (P * P);
    const int gx = (int)(row % G), gy = (int)((row / G) % G), b = (int)(row / ((int64_t)G * G));
    const float4 v = *reinterpret_cast<const float4*>(img + (((int64_t)b * C + c) * R + gy * P + ky) * R + gx * P + kx);
    *reinterpret_cast<uint2*>(out + e) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

// out[b, 0, :] = cls + pos[0];  out[b, n, :] = patch_emb[b, n - 1, :] + pos[n]      (eff_vit.py:448-450)
// One thread per 4 columns (H % 4 == 0: 8-byte bf16 load, 16-byte fp32 load / store); rows are walked with 32-bit arithmetic.
__global__ void __launch_bounds__(256) vit_assemble_fwd_kernel(const __nv_bfloat16* __restrict__ pe, const float* __restrict__ cls,
                                                               const float* __restrict__ pos, float* __restrict__ out, int B, int N, int H) {
  const int h4 = H >> 2;
  const int64_t total4 = (int64_t)B * N * h4;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total4; idx += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t row = (uint32_t)(idx / h4);          // b * N + n  (< 2^32 rows)
    const int c = (int)(idx - (int64_t)row * h4) << 2;
    const uint32_t b = row / (uint32_t)N, n = row - b * (uint32_t)N;
    float4 v;
    if (n == 0) {
      v = *reinterpret_cast<const float4*>(cls + c);
    } else {
      const uint2 u = *reinterpret_cast<const uint2*>(pe + ((int64_t)b * (N - 1) + (n - 1)) * H + c);
      const float2 lo = unpack_bf16x2(u.x), hi = unpack_bf16x2(u.y);
      v = make_float4(lo.x, lo.y, hi.x, hi.y);
    }
    const float4 p = *reinterpret_cast<const float4*>(pos + (int64_t)n * H + c);
    *reinterpret_cast<float4*>(out + (int64_t)row * H + c) = make_float4(v.x + p.x, v.y + p.y, v.z + p.z, v.w + p.w);
  }
}

// one block per token position n: dpos[n,:] += sum_b dh[b,n,:]; n==0 also feeds dcls; n>0 writes dpatch (bf16)
__global__ void vit_assemble_bwd_kernel(const float* __restrict__ dh, __nv_bfloat16* __restrict__ dpatch, float* __restrict__ dcls,
                                        float* __restrict__ dpos, int B, int N, int H) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) {
      const float v = dh[((int64_t)b * N + n) * H + c];
      acc += v;
      if (n > 0 && dpatch) dpatch[((int64_t)b * (N - 1) + (n - 1)) * H + c] = __float2bfloat16(v);
    }
    if (dpos) dpos[(int64_t)n * H + c] += acc;
    if (n == 0 && dcls) dcls[c] += acc;
  }
}

// ------------------------------------------------------------------------------------------------ BERT embeddings
__global__ void bert_embed_fwd_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ tt, const int64_t* __restrict__ pids,
                                      const float* __restrict__ word, const float* __restrict__ type, const float* __restrict__ pos,
                                      float* __restrict__ out, int64_t rows, int L, int H, int past) {
  const int64_t row = blockIdx.x;
  if (row >= rows) return;
  const int64_t w = ids[row];
  const int64_t ty = tt ? tt[row] : 0;
  const int64_t ps = pids ? pids[row] : (row % L) + past;
  for (int c = threadIdx.x * 4; c < H; c += blockDim.x * 4) {
    const float4 a = *reinterpret_cast<const float4*>(word + w * H + c);
    const float4 b = *reinterpret_cast<const float4*>(type + ty * H + c);
    const float4 d = *reinterpret_cast<const float4*>(pos + ps * H + c);
    *reinterpret_cast<float4*>(out + row * H + c) = make_float4(a.x + b.x + d.x, a.y + b.y + d.y, a.z + b.z + d.z, a.w + b.w + d.w);
  }
}
__global__ void bert_embed_bwd_kernel(const float* __restrict__ dout, const int64_t* __restrict__ ids, const int64_t* __restrict__ tt,
                                      const int64_t* __restrict__ pids, float* __restrict__ dword, float* __restrict__ dtype,
                                      float* __restrict__ dpos, int64_t rows, int L, int H, int past, int64_t padding_idx) {
  const int64_t row = blockIdx.x;
  if (row >= rows) return;
  const int64_t w = ids[row];
  if (w == padding_idx) dword = nullptr;   // nn.Embedding(padding_idx=...): the padding row receives no gradient from look-ups
  const int64_t ty = tt ? tt[row] : 0;
  const int64_t ps = pids ? pids[row] : (row % L) + past;
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    const float g = dout[row * H + c];
    if (dword) atomicAdd(dword + w * H + c, g);
    if (dtype) atomicAdd(dtype + ty * H + c, g);
    if (dpos) atomicAdd(dpos + ps * H + c, g);
  }
}

static inline unsigned grid_for(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = 148 * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace evlm
using namespace evlm;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define COUNT(n) g_launch_count.fetch_add(n, std::memory_order_relaxed)

extern "C" int evlm_cast_f32_to_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int64_t cols, float dropout_p,
                                     uint64_t seed, uint32_t stream_id, void* stream) {
  if (!src || !dst || rows < 0 || cols < 0 || dropout_p < 0.f || dropout_p >= 1.f) return EVLM_EINVAL;
  if (rows == 0 || cols == 0) return EVLM_OK;
  const int64_t n = rows * cols;
  if (lds == cols && ldd == cols && (n % 8) == 0 && (cols % 4) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    cast_f32_bf16_flat_kernel<<<grid_for(n / 8, 256), 256, 0, ST(stream)>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<uint4*>(dst),
                                                                           n / 8, dropout_p, seed, stream_id);
    COUNT(1);
    EVLM_CUDA_RETURN();
  }
  cast_f32_bf16_kernel<<<grid_for(rows * ((cols + 3) / 4), 256), 256, 0, ST(stream)>>>(src, lds, reinterpret_cast<__nv_bfloat16*>(dst), ldd, rows,
                                                                                     cols, dropout_p, seed, stream_id);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
extern "C" int evlm_cast_bf16_to_f32(const void* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int64_t cols, void* stream) {
  if (!src || !dst || rows < 0 || cols < 0) return EVLM_EINVAL;
  if (rows == 0 || cols == 0) return EVLM_OK;
  cast_bf16_f32_kernel<<<grid_for(rows * cols, 256), 256, 0, ST(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(src), lds, dst, ldd, rows, cols);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
extern "C" int evlm_colsum(const void* X, int32_t x_dtype, int64_t ldx, int64_t rows, int64_t cols, float* out, int32_t accumulate, void* stream) {
  if (!X || !out || rows < 0 || cols <= 0) return EVLM_EINVAL;
  cudaStream_t st = ST(stream);
  if (!accumulate) cudaMemsetAsync(out, 0, cols * sizeof(float), st);
  if (rows == 0) return EVLM_OK;
  if (x_dtype == EVLM_BF16 && (cols % 8) == 0 && (ldx % 8) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0) {
    const int cb = (int)((cols + 255) / 256);
    int rb = (148 * 4 + cb - 1) / cb;
    if (rb > (rows + 63) / 64) rb = (int)((rows + 63) / 64);
    if (rb < 1) rb = 1;
    const int rpb8 = (int)((rows + rb - 1) / rb);
    colsum_bf16x8_kernel<<<dim3(cb, rb), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(X), ldx, rows, cols, out, rpb8);
    COUNT(1);
    EVLM_CUDA_RETURN();
  }
  const int col_blocks = (int)((cols + 31) / 32);
  int row_blocks = (148 * 4 + col_blocks - 1) / col_blocks;
  if (row_blocks > (rows + 63) / 64) row_blocks = (int)((rows + 63) / 64);
  if (row_blocks < 1) row_blocks = 1;
  const int rpb = (int)((rows + row_blocks - 1) / row_blocks);
  dim3 grid(col_blocks, row_blocks), blk(32, 8);
  if (x_dtype == EVLM_BF16) colsum_kernel<true><<<grid, blk, 0, st>>>(X, ldx, rows, cols, out, rpb);
  else colsum_kernel<false><<<grid, blk, 0, st>>>(X, ldx, rows, cols, out, rpb);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
extern "C" int evlm_coldot(const void* X, const void* Y, int64_t ld, int64_t rows, int64_t cols, float* out, void* stream) {
  if (!X || !Y || !out || rows < 0 || cols <= 0) return EVLM_EINVAL;
  cudaStream_t st = ST(stream);
  cudaMemsetAsync(out, 0, cols * sizeof(float), st);
  if (rows == 0) return EVLM_OK;
  const int col_blocks = (int)((cols + 31) / 32);
  int row_blocks = (148 * 4 + col_blocks - 1) / col_blocks;
  if (row_blocks > (rows + 63) / 64) row_blocks = (int)((rows + 63) / 64);
  if (row_blocks < 1) row_blocks = 1;
  const int rpb = (int)((rows + row_blocks - 1) / row_blocks);
  dim3 grid(col_blocks, row_blocks), blk(32, 8);
  coldot_kernel<<<grid, blk, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(X), reinterpret_cast<const __nv_bfloat16*>(Y), ld, rows, cols, out, rpb);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
extern "C" int evlm_im2col_patch(const float* image, void* patches, int B, int C, int R, int P, void* stream) {
  if (!image || !patches || B <= 0 || C <= 0 || R <= 0 || P <= 0 || (R % P) || (P % 4)) return EVLM_EINVAL;
  const int64_t total4 = (int64_t)B * R * R * C / 4;
  im2col_kernel<<<grid_for(total4, 256), 256, 0, ST(stream)>>>(image, reinterpret_cast<__nv_bfloat16*>(patches), B, C, R, P);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
extern "C" int evlm_vit_assemble_fwd(const void* patch_emb, const float* cls, const float* pos, float* out, int B, int N, int H, void* stream) {
  if (!patch_emb || !cls || !pos || !out || B <= 0 || N <= 1 || H <= 0 || (H & 3)) return EVLM_EINVAL;
  if ((reinterpret_cast<uintptr_t>(patch_emb) & 7) || ((reinterpret_cast<uintptr_t>(cls) | reinterpret_cast<uintptr_t>(pos) | reinterpret_cast<uintptr_t>(out)) & 15))
    return EVLM_EINVAL;
  vit_assemble_fwd_kernel<<<grid_for((int64_t)B * N * H / 4, 256), 256, 0, ST(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(patch_emb), cls, pos, out,
                                                                                   B, N, H);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
extern "C" int evlm_vit_assemble_bwd(const float* dh, void* dpatch, float* dcls, float* dpos, int B, int N, int H, void* stream) {
  if (!dh || B <= 0 || N <= 1 || H <= 0) return EVLM_EINVAL;
  vit_assemble_bwd_kernel<<<N, 256, 0, ST(stream)>>>(dh, reinterpret_cast<__nv_bfloat16*>(dpatch), dcls, dpos, B, N, H);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
extern "C" int evlm_bert_embed_fwd(const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids, const float* word, const float* type,
                                   const float* pos, float* out, int64_t rows, int L, int H, int past_len, int64_t vocab, void* stream) {
  if (!ids || !word || !type || !pos || !out || rows < 0 || L <= 0 || H <= 0 || (H & 3)) return EVLM_EINVAL;
  (void)vocab;
  if (rows == 0) return EVLM_OK;
  bert_embed_fwd_kernel<<<(unsigned)rows, 192, 0, ST(stream)>>>(ids, type_ids, pos_ids, word, type, pos, out, rows, L, H, past_len);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
extern "C" int evlm_bert_embed_bwd(const float* dout, const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids, float* dword,
                                   float* dtype, float* dpos, int64_t rows, int L, int H, int past_len, int64_t padding_idx, void* stream) {
  if (!dout || !ids || rows < 0 || L <= 0 || H <= 0) return EVLM_EINVAL;
  if (rows == 0) return EVLM_OK;
  bert_embed_bwd_kernel<<<(unsigned)rows, 256, 0, ST(stream)>>>(dout, ids, type_ids, pos_ids, dword, dtype, dpos, rows, L, H, past_len, padding_idx);
  COUNT(1);
  EVLM_CUDA_RETURN();
}

// ------------------------------------------------------------------------------------------------ standalone activations
namespace evlm {
__device__ __forceinline__ float ld_any(const void* p, int dt, int64_t i) {
  return dt == EVLM_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]) : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void st_any(void* p, int dt, int64_t i, float v) {
  if (dt == EVLM_BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16(v);
  else reinterpret_cast<float*>(p)[i] = v;
}
__global__ void act_fwd_kernel(const void* __restrict__ x, int xdt, void* __restrict__ y, int ydt, int64_t n, int act) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = ld_any(x, xdt, i);
    st_any(y, ydt, i, act == EVLM_ACT_QUICK_GELU ? quick_gelu(v) : act == EVLM_ACT_GELU_ERF ? gelu_erf(v) : v);
  }
}
__global__ void act_bwd_kernel(const void* __restrict__ dy, int dydt, const void* __restrict__ x, int xdt, void* __restrict__ dx, int dxdt,
                               int64_t n, int act) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = ld_any(x, xdt, i);
    const float d = act == EVLM_ACT_QUICK_GELU ? quick_gelu_grad(v) : act == EVLM_ACT_GELU_ERF ? gelu_erf_grad(v) : 1.f;
    st_any(dx, dxdt, i, ld_any(dy, dydt, i) * d);
  }
}
}  // namespace evlm
extern "C" int evlm_act_fwd(const void* x, int32_t x_dtype, void* y, int32_t y_dtype, int64_t n, int32_t act, void* stream) {
  if (!x || !y || n < 0) return EVLM_EINVAL;
  if (n == 0) return EVLM_OK;
  act_fwd_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(x, x_dtype, y, y_dtype, n, act);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
extern "C" int evlm_act_bwd(const void* dy, int32_t dy_dtype, const void* x, int32_t x_dtype, void* dx, int32_t dx_dtype, int64_t n,
                            int32_t act, void* stream) {
  if (!dy || !x || !dx || n < 0) return EVLM_EINVAL;
  if (n == 0) return EVLM_OK;
  act_bwd_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(dy, dy_dtype, x, x_dtype, dx, dx_dtype, n, act);
  COUNT(1);
  EVLM_CUDA_RETURN();
}

// evlm_rng_bind() reaches the per-translation-unit seed-offset pointer through this hook (evlm_common.cuh).
namespace evlm { cudaError_t rng_bind_elementwise(const void* state_dev) { return tu_rng_bind(state_dev); } }

// ------------------------------------------------------------------------------------------------ shared K/V gradient fold
namespace evlm {
// dst[index[i], :] += src[i, :]: 8 bf16 per thread, two fp32 vector reductions (items that share a K/V item collide rarely
// in time; the L2 resolves them)
__global__ void __launch_bounds__(256) index_add_rows_kernel(const __nv_bfloat16* __restrict__ src, const int32_t* __restrict__ index,
                                                             float* __restrict__ dst, int64_t n_src, int64_t row_elems) {
  const int64_t chunks = row_elems >> 3;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_src * chunks; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / chunks, c = t - i * chunks;
    const uint4 q = *reinterpret_cast<const uint4*>(src + i * row_elems + c * 8);
    const float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), d = unpack_bf16x2(q.z), e = unpack_bf16x2(q.w);
    float* o = dst + (int64_t)index[i] * row_elems + c * 8;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y) : "memory");
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4), "f"(d.x), "f"(d.y), "f"(e.x), "f"(e.y) : "memory");
  }
}
}  // namespace evlm
extern "C" int evlm_index_add_rows(const void* src_bf16, const int32_t* index, float* dst, int64_t n_src, int64_t row_elems, void* stream) {
  if (!src_bf16 || !index || !dst || n_src < 0 || row_elems <= 0 || (row_elems & 7)) return EVLM_EINVAL;
  if ((reinterpret_cast<uintptr_t>(src_bf16) | reinterpret_cast<uintptr_t>(dst)) & 15) return EVLM_EINVAL;
  if (n_src == 0) return EVLM_OK;
  index_add_rows_kernel<<<148 * 8, 256, 0, ST(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(src_bf16), index, dst, n_src, row_elems);
  COUNT(1);
  EVLM_CUDA_RETURN();
}

// Deterministic fold: dst[u, :] = sum over {i : index[i] == u} src[i, :]  (bf16 in, fp32 accumulate, bf16 out), no atomics on the
// data: a one-block kernel builds the CSR lists of the (few thousand) source rows, the fold kernel then streams every row once.
namespace evlm {
__global__ void __launch_bounds__(1024) fold_build_kernel(const int32_t* __restrict__ index, int n_src, int n_dst, int32_t* __restrict__ offsets,
                                                          int32_t* __restrict__ cursor, int32_t* __restrict__ list) {
  __shared__ int part[1024];
  const int t = threadIdx.x;
  for (int u = t; u < n_dst; u += 1024) cursor[u] = 0;
  __syncthreads();
  for (int i = t; i < n_src; i += 1024) {
    const int u = index[i];
    if (u >= 0 && u < n_dst) atomicAdd(&cursor[u], 1);
  }
  __syncthreads();
  // exclusive scan of the counts: each thread owns a contiguous segment
  const int seg = (n_dst + 1023) / 1024;
  const int lo = min(t * seg, n_dst), hi = min(lo + seg, n_dst);
  int sum = 0;
  for (int u = lo; u < hi; ++u) sum += cursor[u];
  part[t] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int v = t >= o ? part[t - o] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  int run = part[t] - sum;
  for (int u = lo; u < hi; ++u) {
    const int c = cursor[u];
    offsets[u] = run;
    cursor[u] = run;
    run += c;
  }
  if (t == 1023) offsets[n_dst] = part[1023];
  __syncthreads();
  for (int i = t; i < n_src; i += 1024) {
    const int u = index[i];
    if (u >= 0 && u < n_dst) list[atomicAdd(&cursor[u], 1)] = i;
  }
  __syncthreads();
  // fixed summation order: sort every (short) list
  for (int u = t; u < n_dst; u += 1024) {
    const int a0 = offsets[u], a1 = offsets[u + 1];
    for (int x = a0 + 1; x < a1; ++x) {
      const int v = list[x];
      int y = x - 1;
      while (y >= a0 && list[y] > v) { list[y + 1] = list[y]; --y; }
      list[y + 1] = v;
    }
  }
}
__global__ void __launch_bounds__(256) fold_rows_kernel(const __nv_bfloat16* __restrict__ src, const int32_t* __restrict__ offsets,
                                                        const int32_t* __restrict__ list, __nv_bfloat16* __restrict__ dst, int64_t n_dst,
                                                        int64_t row_elems) {
  const int64_t chunks = row_elems >> 3;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_dst * chunks; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = t / chunks, c = t - u * chunks;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = offsets[u]; k < offsets[u + 1]; ++k) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(src + (int64_t)list[k] * row_elems + c * 8));
      const float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), d = unpack_bf16x2(q.z), e = unpack_bf16x2(q.w);
      acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y; acc[4] += d.x; acc[5] += d.y; acc[6] += e.x; acc[7] += e.y;
    }
    *reinterpret_cast<uint4*>(dst + u * row_elems + c * 8) =
        make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
  }
}
}  // namespace evlm
extern "C" int evlm_index_fold_rows(const void* src_bf16, const int32_t* index, int64_t n_src, int64_t n_dst, int64_t row_elems, void* dst_bf16,
                                    int32_t* workspace, void* stream) {
  if (!src_bf16 || !index || !dst_bf16 || !workspace || n_src < 0 || n_dst <= 0 || row_elems <= 0 || (row_elems & 7)) return EVLM_EINVAL;
  if (n_src > (1 << 30) || n_dst > (1 << 20)) return EVLM_EUNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(src_bf16) | reinterpret_cast<uintptr_t>(dst_bf16)) & 15) return EVLM_EINVAL;
  int32_t* offsets = workspace;                 // [n_dst + 1]
  int32_t* cursor = workspace + n_dst + 1;      // [n_dst]
  int32_t* list = cursor + n_dst;               // [n_src]
  fold_build_kernel<<<1, 1024, 0, ST(stream)>>>(index, (int)n_src, (int)n_dst, offsets, cursor, list);
  fold_rows_kernel<<<148 * 8, 256, 0, ST(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(src_bf16), offsets, list,
                                                    reinterpret_cast<__nv_bfloat16*>(dst_bf16), n_dst, row_elems);
  COUNT(2);
  EVLM_CUDA_RETURN();
}

// ------------------------------------------------------------------------------------------------ CUDA-graph support
namespace evlm {
cudaError_t rng_bind_attention(const void*);
cudaError_t rng_bind_attention_tc(const void*);
cudaError_t rng_bind_attention_tc_bwd(const void*);
cudaError_t rng_bind_attention_tc_long(const void*);
cudaError_t rng_bind_gemm_tcgen05(const void*);
cudaError_t rng_bind_layernorm(const void*);
struct F32Payload { float v[32]; };
__global__ void store_f32_kernel(float* __restrict__ dst, F32Payload p, int n) {
  if ((int)threadIdx.x < n) dst[threadIdx.x] = p.v[threadIdx.x];
}
__global__ void rng_advance_kernel(unsigned long long* state, unsigned long long delta, int set) {
  *state = set ? delta : *state + delta;
}
}  // namespace evlm
extern "C" int evlm_rng_bind(const uint64_t* state_dev) {
  cudaError_t e;
  if ((e = rng_bind_attention(state_dev)) != cudaSuccess) return (int)e;
  if ((e = rng_bind_attention_tc(state_dev)) != cudaSuccess) return (int)e;
  if ((e = rng_bind_attention_tc_bwd(state_dev)) != cudaSuccess) return (int)e;
  if ((e = rng_bind_attention_tc_long(state_dev)) != cudaSuccess) return (int)e;
  if ((e = rng_bind_gemm_tcgen05(state_dev)) != cudaSuccess) return (int)e;
  if ((e = rng_bind_layernorm(state_dev)) != cudaSuccess) return (int)e;
  if ((e = rng_bind_elementwise(state_dev)) != cudaSuccess) return (int)e;
  return EVLM_OK;
}
extern "C" int evlm_rng_advance(uint64_t* state_dev, uint64_t delta, int32_t set, void* stream) {
  if (!state_dev) return EVLM_EINVAL;
  rng_advance_kernel<<<1, 1, 0, ST(stream)>>>(reinterpret_cast<unsigned long long*>(state_dev), (unsigned long long)delta, set);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
extern "C" int evlm_store_f32(float* dst_dev, const float* values_host, int32_t n, void* stream) {
  if (!dst_dev || !values_host || n < 0 || n > 32) return EVLM_EINVAL;
  if (n == 0) return EVLM_OK;
  F32Payload p;
  for (int i = 0; i < 32; ++i) p.v[i] = i < n ? values_host[i] : 0.f;
  store_f32_kernel<<<1, 32, 0, ST(stream)>>>(dst_dev, p, n);
  COUNT(1);
  EVLM_CUDA_RETURN();
}

// ------------------------------------------------------------------------------------------------ zero-skip index work
// (include/evlm.h: evlm_compact_index / evlm_gather_* / evlm_scatter_*).  The gate vectors are at most a few thousand entries: one
// block, one scan.  Everything downstream reads `count` from device memory, so a captured step graph replays with fresh masks.
namespace evlm {
__global__ void __launch_bounds__(1024) compact_index_kernel(const float* __restrict__ z, int n, int* __restrict__ idx, int* __restrict__ count) {
  __shared__ int warp_tot[32];
  __shared__ int base_kept, base_drop;
  if (threadIdx.x == 0) { base_kept = 0; base_drop = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // pass 1: number kept
  int mine = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) mine += z[i] != 0.f;
  mine = __reduce_add_sync(0xffffffffu, mine);
  if (lane == 0) warp_tot[w] = mine;
  __syncthreads();
  int total = 0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) total += warp_tot[i];
  __syncthreads();
  // pass 2: stable positions, chunk by chunk (kept entries first, the others after them, both ascending)
  for (int c0 = 0; c0 < n; c0 += blockDim.x) {
    const int i = c0 + threadIdx.x;
    const bool valid = i < n;
    const bool k = valid && z[i] != 0.f;
    const unsigned bk = __ballot_sync(0xffffffffu, k), bd = __ballot_sync(0xffffffffu, valid && !k);
    if (lane == 0) warp_tot[w] = __popc(bk) | (__popc(bd) << 16);
    __syncthreads();
    int pk = 0, pd = 0;
    for (int j = 0; j < w; ++j) { pk += warp_tot[j] & 0xffff; pd += warp_tot[j] >> 16; }
    const unsigned below = (1u << lane) - 1u;
    if (k) idx[base_kept + pk + __popc(bk & below)] = i;
    else if (valid) idx[total + base_drop + pd + __popc(bd & below)] = i;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tk = 0, td = 0;
      for (int j = 0; j < (int)(blockDim.x >> 5); ++j) { tk += warp_tot[j] & 0xffff; td += warp_tot[j] >> 16; }
      base_kept += tk;
      base_drop += td;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) count[0] = total;
}
// dst[j, :] = j < count ? src[idx[j], :] : 0 ; element size ES bytes, 16-byte vectors when the pitches allow
template <int ES>
__global__ void gather_rows_kernel(const uint8_t* __restrict__ src, int64_t lds, const int* __restrict__ idx, const int* __restrict__ count,
                                   uint8_t* __restrict__ dst, int64_t ldd, int64_t rows, int64_t cols, bool vec) {
  const int cnt = count[0];
  const int64_t row_bytes = cols * ES;
  for (int64_t j = blockIdx.x; j < rows; j += gridDim.x) {
    const bool keep = j < cnt;
    const uint8_t* s = src + (keep ? (int64_t)idx[j] : 0) * lds * ES;
    uint8_t* d = dst + j * ldd * ES;
    if (vec) {
      for (int64_t b = (int64_t)threadIdx.x * 16; b < row_bytes; b += (int64_t)blockDim.x * 16)
        *reinterpret_cast<uint4*>(d + b) = keep ? *reinterpret_cast<const uint4*>(s + b) : make_uint4(0u, 0u, 0u, 0u);
    } else {
      for (int64_t b = (int64_t)threadIdx.x * ES; b < row_bytes; b += (int64_t)blockDim.x * ES) {
        if (ES == 2) *reinterpret_cast<uint16_t*>(d + b) = keep ? *reinterpret_cast<const uint16_t*>(s + b) : (uint16_t)0;
        else *reinterpret_cast<uint32_t*>(d + b) = keep ? *reinterpret_cast<const uint32_t*>(s + b) : 0u;
      }
    }
  }
}
// dst[r, j] = j < count ? src[r, idx[j]] : 0   (bf16)
__global__ void gather_cols_bf16_kernel(const uint16_t* __restrict__ src, int64_t lds, const int* __restrict__ idx, const int* __restrict__ count,
                                        uint16_t* __restrict__ dst, int64_t ldd, int64_t rows, int64_t cols) {
  const int cnt = count[0];
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x)
    for (int64_t j = threadIdx.x; j < cols; j += blockDim.x) dst[r * ldd + j] = j < cnt ? src[r * lds + idx[j]] : (uint16_t)0;
}
// dst[idx[j], :] (+)= src[j, :] for j < count; without `accumulate` the rows idx[count..rows) are zeroed
__global__ void scatter_rows_kernel(const float* __restrict__ src, int64_t lds, const int* __restrict__ idx, const int* __restrict__ count,
                                    float* __restrict__ dst, int64_t ldd, int64_t rows, int64_t cols, int accumulate) {
  const int cnt = count[0];
  for (int64_t j = blockIdx.x; j < rows; j += gridDim.x) {
    const bool keep = j < cnt;
    if (!keep && accumulate) continue;
    float* d = dst + (int64_t)idx[j] * ldd;
    const float* s = src + j * lds;
    for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) d[c] = keep ? (accumulate ? d[c] + s[c] : s[c]) : 0.f;
  }
}
// dst[r, idx[j]] (+)= src[r, j] for j < count; without `accumulate` the columns idx[count..cols) are zeroed
__global__ void scatter_cols_kernel(const float* __restrict__ src, int64_t lds, const int* __restrict__ idx, const int* __restrict__ count,
                                    float* __restrict__ dst, int64_t ldd, int64_t rows, int64_t cols, int accumulate) {
  const int cnt = count[0];
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x)
    for (int64_t j = threadIdx.x; j < cols; j += blockDim.x) {
      const bool keep = j < cnt;
      if (!keep && accumulate) continue;
      float* d = dst + r * ldd + idx[j];
      *d = keep ? (accumulate ? *d + src[r * lds + j] : src[r * lds + j]) : 0.f;
    }
}
}  // namespace evlm

extern "C" int evlm_compact_index(const float* z, int32_t n, int32_t* idx, int32_t* count, void* stream) {
  using namespace evlm;
  if (!z || !idx || !count || n <= 0 || n > 65536) return EVLM_EINVAL;
  compact_index_kernel<<<1, 1024, 0, ST(stream)>>>(z, n, idx, count);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
extern "C" int evlm_gather_rows(const void* src, int64_t lds, int32_t dtype, const int32_t* idx, const int32_t* count, void* dst, int64_t ldd,
                                int64_t rows, int64_t cols, void* stream) {
  using namespace evlm;
  if (!src || !idx || !count || !dst || rows <= 0 || cols <= 0) return EVLM_EINVAL;
  const int es = dtype == EVLM_F32 ? 4 : 2;
  const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | (uintptr_t)(lds * es) | (uintptr_t)(ldd * es) |
                     (uintptr_t)(cols * es)) & 15) == 0;
  const int threads = cols * es >= 4096 ? 256 : (cols * es >= 512 ? 64 : 32);
  const unsigned blocks = (unsigned)(rows < 148 * 16 ? rows : 148 * 16);
  if (es == 4) gather_rows_kernel<4><<<blocks, threads, 0, ST(stream)>>>((const uint8_t*)src, lds, idx, count, (uint8_t*)dst, ldd, rows, cols, vec);
  else gather_rows_kernel<2><<<blocks, threads, 0, ST(stream)>>>((const uint8_t*)src, lds, idx, count, (uint8_t*)dst, ldd, rows, cols, vec);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
extern "C" int evlm_gather_cols_bf16(const void* src, int64_t lds, const int32_t* idx, const int32_t* count, void* dst, int64_t ldd, int64_t rows,
                                     int64_t cols, void* stream) {
  using namespace evlm;
  if (!src || !idx || !count || !dst || rows <= 0 || cols <= 0) return EVLM_EINVAL;
  gather_cols_bf16_kernel<<<(unsigned)(rows < 148 * 8 ? rows : 148 * 8), 256, 0, ST(stream)>>>((const uint16_t*)src, lds, idx, count, (uint16_t*)dst,
                                                                                                ldd, rows, cols);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
extern "C" int evlm_scatter_rows_add(const float* src, int64_t lds, const int32_t* idx, const int32_t* count, float* dst, int64_t ldd, int64_t rows,
                                     int64_t cols, int32_t accumulate, void* stream) {
  using namespace evlm;
  if (!src || !idx || !count || !dst || rows <= 0 || cols <= 0) return EVLM_EINVAL;
  scatter_rows_kernel<<<(unsigned)(rows < 148 * 16 ? rows : 148 * 16), cols >= 512 ? 256 : 32, 0, ST(stream)>>>(src, lds, idx, count, dst, ldd, rows,
                                                                                                                  cols, accumulate);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
extern "C" int evlm_scatter_cols_add(const float* src, int64_t lds, const int32_t* idx, const int32_t* count, float* dst, int64_t ldd, int64_t rows,
                                     int64_t cols, int32_t accumulate, void* stream) {
  using namespace evlm;
  if (!src || !idx || !count || !dst || rows <= 0 || cols <= 0) return EVLM_EINVAL;
  scatter_cols_kernel<<<(unsigned)(rows < 148 * 8 ? rows : 148 * 8), 256, 0, ST(stream)>>>(src, lds, idx, count, dst, ldd, rows, cols, accumulate);
  COUNT(1);
  EVLM_CUDA_RETURN();
}


// ------------------------------------------------------------------------------------------------ multi-tensor cast
// evlm_cast_table: the bf16 shadows of ALL weights an optimizer step touched, refreshed by ONE launch (the per-weight casts were ~100
// launches per step at 0.27 of HBM bandwidth: most weights are a few MB, a launch each cannot fill the machine).  Every block walks
// every entry with a grid stride, like the multi-pair MSE kernel.
namespace evlm {
__global__ void __launch_bounds__(256) cast_table_kernel(const evlm_cast_entry* __restrict__ tab, int n) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  for (int e = 0; e < n; ++e) {
    const evlm_cast_entry en = tab[e];
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(en.dst);
    const bool vec = (en.cols & 7) == 0 && (en.ldd & 7) == 0 && ((reinterpret_cast<uintptr_t>(en.src) | reinterpret_cast<uintptr_t>(en.dst)) & 15) == 0;
    if (vec) {
      const int64_t c8 = en.cols >> 3, n8 = en.rows * c8;
      for (int64_t i = tid; i < n8; i += stride) {
        const int64_t r = i / c8, c = (i - r * c8) << 3;
        const float4 a = *reinterpret_cast<const float4*>(en.src + r * en.cols + c), b = *reinterpret_cast<const float4*>(en.src + r * en.cols + c + 4);
        *reinterpret_cast<uint4*>(dst + r * en.ldd + c) =
            make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
      }
    } else {
      const int64_t nel = en.rows * en.cols;
      for (int64_t i = tid; i < nel; i += stride) {
        const int64_t r = i / en.cols, c = i - r * en.cols;
        dst[r * en.ldd + c] = __float2bfloat16(en.src[i]);
      }
    }
  }
}
}  // namespace evlm
extern "C" int evlm_cast_table(const evlm_cast_entry* table_dev, int32_t n, void* stream) {
  using namespace evlm;
  if (!table_dev || n <= 0) return EVLM_EINVAL;
  cast_table_kernel<<<148 * 8, 256, 0, ST(stream)>>>(table_dev, n);
  COUNT(1);
  EVLM_CUDA_RETURN();
}

// ---- greedy token selection of the decode loop (eff_bert.py:1510-1538 with do_sample = False, repetition_penalty = 1): per sequence
//   next = argmax(logits) (first maximal index, like torch.argmax); score = log_softmax(logits)[next];
//   tokens_to_add = next * unfinished + pad * (1 - unfinished); unfinished_out = unfinished * prod_e (tokens_to_add != eos_e)
// One block per row, one pass over the vocabulary (online log-sum-exp): replaces 17 framework launches per decoded token.
namespace evlm {
__global__ void __launch_bounds__(256) greedy_select_kernel(const float* __restrict__ logits, int64_t ld, int32_t vocab,
                                                            const int64_t* __restrict__ unfinished, int64_t pad, int32_t n_eos, int64_t eos0,
                                                            int64_t eos1, int64_t eos2, int64_t eos3, int64_t* __restrict__ next_token,
                                                            float* __restrict__ score, int64_t* __restrict__ tokens_to_add,
                                                            int64_t* __restrict__ unfinished_out) {
  __shared__ float s_v[8], s_m[8], s_l[8];
  __shared__ int s_i[8];
  const float* x = logits + (int64_t)blockIdx.x * ld;
  float best = -INFINITY, m = -INFINITY, l = 0.f;
  int bi = 0x7fffffff;
  for (int j = threadIdx.x; j < vocab; j += 256) {
    const float v = x[j];
    if (v > best) { best = v; bi = j; }      // ascending j per thread: the first maximal index survives
    const float mn = fmaxf(m, v);
    l = l * (m == -INFINITY ? 0.f : __expf(m - mn)) + __expf(v - mn);
    m = mn;
  }
  auto merge = [](float& v1, int& i1, float& m1, float& l1, float v2, int i2, float m2, float l2) {
    if (v2 > v1 || (v2 == v1 && i2 < i1)) { v1 = v2; i1 = i2; }
    const float mn = fmaxf(m1, m2);
    l1 = l1 * (m1 == -INFINITY ? 0.f : __expf(m1 - mn)) + l2 * (m2 == -INFINITY ? 0.f : __expf(m2 - mn));
    m1 = mn;
  };
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    merge(best, bi, m, l, __shfl_xor_sync(0xffffffffu, best, o), __shfl_xor_sync(0xffffffffu, bi, o), __shfl_xor_sync(0xffffffffu, m, o),
          __shfl_xor_sync(0xffffffffu, l, o));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_v[warp] = best; s_i[warp] = bi; s_m[warp] = m; s_l[warp] = l; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) merge(best, bi, m, l, s_v[w], s_i[w], s_m[w], s_l[w]);
    const int64_t r = blockIdx.x, u = unfinished[r];
    const int64_t tok = (int64_t)bi * u + pad * (1 - u);
    int64_t un = u;
    const int64_t eos[4] = {eos0, eos1, eos2, eos3};
    for (int e = 0; e < n_eos; ++e) un *= (tok != eos[e]) ? 1 : 0;
    next_token[r] = bi;
    score[r] = best - (m + __logf(l));
    tokens_to_add[r] = tok;
    unfinished_out[r] = un;
  }
}
}  // namespace evlm
extern "C" int evlm_greedy_select(const float* logits, int64_t ld, int32_t rows, int32_t vocab, const int64_t* unfinished, int64_t pad,
                                  const int64_t* eos_host, int32_t n_eos, int64_t* next_token, float* score, int64_t* tokens_to_add,
                                  int64_t* unfinished_out, void* stream) {
  using namespace evlm;
  if (!logits || !unfinished || !next_token || !score || !tokens_to_add || !unfinished_out || rows <= 0 || vocab <= 0 || n_eos < 0 || n_eos > 4 ||
      (n_eos > 0 && !eos_host))
    return EVLM_EINVAL;
  int64_t e[4] = {0, 0, 0, 0};
  for (int i = 0; i < n_eos; ++i) e[i] = eos_host[i];
  greedy_select_kernel<<<(unsigned)rows, 256, 0, ST(stream)>>>(logits, ld, vocab, unfinished, pad, n_eos, e[0], e[1], e[2], e[3], next_token, score,
                                                               tokens_to_add, unfinished_out);
  COUNT(1);
  EVLM_CUDA_RETURN();
}
