// Standalone bring-up / regression harness for evlm_gemm_bf16 (links libevlm_b200.so).
//   gemm_test <case>      one case per process so that a trap in one does not take the others down.
// Compares against a naive CUDA-core GEMM on the same bf16 inputs (fp32 accumulate) and, for the
// epilogue cases, against a host re-statement of the epilogue.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../../include/evlm.h"

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

__global__ void naive_gemm(const __nv_bfloat16* A, int64_t lda, int a_mn, const __nv_bfloat16* B, int64_t ldb, int b_mn, float* D,
                           int M, int N, int K) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int m = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= M || n >= N) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    float a = __bfloat162float(a_mn ? A[(int64_t)k * lda + m] : A[(int64_t)m * lda + k]);
    float b = __bfloat162float(b_mn ? B[(int64_t)k * ldb + n] : B[(int64_t)n * ldb + k]);
    acc += a * b;
  }
  D[(int64_t)m * N + n] = acc;
}

__global__ void fill_bf16(__nv_bfloat16* p, int64_t n, uint32_t seed, float scale) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t x = (uint32_t)i * 2654435761u ^ seed;
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  p[i] = __float2bfloat16(((x & 0xffff) / 65536.f - 0.5f) * scale);
}
__global__ void fill_f32(float* p, int64_t n, uint32_t seed, float scale, float offset) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t x = (uint32_t)i * 2654435761u ^ seed;
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  p[i] = ((x & 0xffff) / 65536.f - 0.5f) * scale + offset;
}

static float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }
static float qgelu(float x) { return x / (1.f + expf(-1.702f * x)); }
static float qgelu_g(float x) { float s = 1.f / (1.f + expf(-1.702f * x)); return s + 1.702f * x * s * (1 - s); }
static float gelu(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678f)); }
static float gelu_g(float x) { return 0.5f * (1.f + erff(x * 0.70710678f)) + x * 0.39894228f * expf(-0.5f * x * x); }

struct Case {
  const char* name;
  int M, N, K, a_mn, b_mn, d_dtype, splits;
  int epi;  // 0 plain, 1 fwd full epilogue (bias, alpha, quick_gelu, pre gate, residual f32, aux_out), 2 fwd gelu post gate bf16 residual,
            // 3 act-backward pre gate quick_gelu, 4 act-backward post gate gelu
  int time_it;
};

static Case cases[] = {
    {"fwd_small", 256, 256, 128, 0, 0, EVLM_F32, 1, 0, 0},
    {"fwd_small_bf16", 256, 256, 128, 0, 0, EVLM_BF16, 1, 0, 0},
    {"fwd_k64", 128, 128, 64, 0, 0, EVLM_F32, 1, 0, 0},
    {"fwd_ragged", 300, 200, 136, 0, 0, EVLM_F32, 1, 0, 0},
    {"dgrad_small", 256, 256, 128, 0, 1, EVLM_F32, 1, 0, 0},
    {"amn_small", 256, 256, 128, 1, 0, EVLM_F32, 1, 0, 0},
    {"wgrad_small", 256, 256, 128, 1, 1, EVLM_F32, 1, 0, 0},
    {"wgrad_split", 768, 768, 2048, 1, 1, EVLM_F32, 4, 0, 0},
    {"wgrad_ragged", 200, 328, 1000, 1, 1, EVLM_F32, 3, 0, 0},
    {"epi_fwd_vit", 384, 512, 256, 0, 0, EVLM_BF16, 1, 1, 0},
    {"epi_fwd_bert", 384, 500, 256, 0, 0, EVLM_F32, 1, 2, 0},
    {"epi_bwd_pre", 384, 512, 256, 0, 1, EVLM_BF16, 1, 3, 0},
    {"epi_bwd_post", 384, 504, 256, 0, 1, EVLM_BF16, 1, 4, 0},
    {"fwd_qkv", 25216, 2304, 768, 0, 0, EVLM_BF16, 1, 0, 1},
    {"fwd_fc1", 25216, 3072, 768, 0, 0, EVLM_BF16, 1, 0, 1},
    {"fwd_fc2", 25216, 768, 3072, 0, 0, EVLM_F32, 1, 0, 1},
    {"dgrad_fc1", 25216, 768, 3072, 0, 1, EVLM_BF16, 1, 0, 1},
    {"dgrad_fc2", 25216, 3072, 768, 0, 1, EVLM_BF16, 1, 0, 1},
    {"wgrad_fc1", 3072, 768, 25216, 1, 1, EVLM_F32, 2, 0, 1},
    {"wgrad_qkv", 2304, 768, 25216, 1, 1, EVLM_F32, 3, 0, 1},
    {"wgrad_proj", 768, 768, 25216, 1, 1, EVLM_F32, 4, 0, 1},
    {"fwd_vocab", 1024, 30522, 768, 0, 0, EVLM_F32, 1, 0, 1},
    {"bert_proj", 5120, 768, 768, 0, 0, EVLM_F32, 1, 0, 1},
    {"bert_qkv", 5120, 2304, 768, 0, 0, EVLM_BF16, 1, 0, 1},
    {"bert_fc1", 5120, 3072, 768, 0, 0, EVLM_BF16, 1, 0, 1},
    {"itm_proj", 15360, 768, 768, 0, 0, EVLM_F32, 1, 0, 1},
    {"bert_wgrad", 768, 768, 5120, 1, 1, EVLM_F32, 4, 0, 1},
    {"tiny", 128, 256, 768, 0, 0, EVLM_F32, 1, 0, 1},
    // the hot path's fused epilogues at full size: ViT fc1 forward (bias, quick-GELU, gate, saved pre-activation),
    // its backward (act-backward on dgrad), out-proj with fp32 residual, BERT output dense with dropout + residual
    {"act_fwd_fc1", 25216, 3072, 768, 0, 0, EVLM_BF16, 1, 5, 1},
    {"act_bwd_fc1", 25216, 3072, 768, 0, 1, EVLM_BF16, 1, 3, 1},
    {"res_proj", 25216, 768, 768, 0, 0, EVLM_F32, 1, 6, 1},
    {"res_fc2", 25216, 768, 3072, 0, 0, EVLM_F32, 1, 6, 1},
    {"bert_out_drop", 5120, 768, 3072, 0, 0, EVLM_F32, 1, 7, 1},
    {"bert_act_fc1", 5120, 3072, 768, 0, 0, EVLM_BF16, 1, 8, 1},
};

int main(int argc, char** argv) {
  if (argc < 2) {
    for (auto& c : cases) printf("%s\n", c.name);
    return 0;
  }
  const Case* c = nullptr;
  for (auto& cc : cases)
    if (!strcmp(cc.name, argv[1])) c = &cc;
  if (!c) { printf("unknown case %s\n", argv[1]); return 1; }
  const int M = c->M, N = c->N, K = c->K;
  // leading dimensions padded to a multiple of 8 elements (TMA needs 16-byte pitches)
  auto pad8 = [](int x) { return (int64_t)((x + 7) / 8 * 8); };
  const int64_t lda = c->a_mn ? pad8(M) : pad8(K);
  const int64_t ldb = c->b_mn ? pad8(N) : pad8(K);
  const int64_t a_rows = c->a_mn ? K : M, b_rows = c->b_mn ? K : N;
  const int64_t ldd = N;  // deliberately unpadded: exercises the unaligned store paths for odd N
  __nv_bfloat16 *A, *B;
  CK(cudaMalloc(&A, a_rows * lda * 2));
  CK(cudaMalloc(&B, b_rows * ldb * 2));
  fill_bf16<<<(a_rows * lda + 255) / 256, 256>>>(A, a_rows * lda, 1234u, 2.0f);
  fill_bf16<<<(b_rows * ldb + 255) / 256, 256>>>(B, b_rows * ldb, 777u, 2.0f);
  float* Dref;
  CK(cudaMalloc(&Dref, (int64_t)M * N * 4));
  void* D;
  const size_t dsz = (size_t)M * ldd * (c->d_dtype == EVLM_F32 ? 4 : 2);
  CK(cudaMalloc(&D, dsz));
  CK(cudaMemset(D, 0, dsz));
  float *bias = nullptr, *gate = nullptr, *res32 = nullptr;
  __nv_bfloat16 *aux_out = nullptr, *aux_in = nullptr, *res16 = nullptr;
  CK(cudaMalloc(&bias, N * 4)); CK(cudaMalloc(&gate, N * 4));
  CK(cudaMalloc(&res32, (int64_t)M * N * 4)); CK(cudaMalloc(&res16, (int64_t)M * N * 2));
  CK(cudaMalloc(&aux_out, (int64_t)M * N * 2)); CK(cudaMalloc(&aux_in, (int64_t)M * N * 2));
  fill_f32<<<(N + 255) / 256, 256>>>(bias, N, 5u, 1.0f, 0.f);
  fill_f32<<<(N + 255) / 256, 256>>>(gate, N, 6u, 1.0f, 0.5f);
  fill_f32<<<((int64_t)M * N + 255) / 256, 256>>>(res32, (int64_t)M * N, 7u, 2.0f, 0.f);
  fill_bf16<<<((int64_t)M * N + 255) / 256, 256>>>(res16, (int64_t)M * N, 8u, 2.0f);
  fill_bf16<<<((int64_t)M * N + 255) / 256, 256>>>(aux_in, (int64_t)M * N, 9u, 4.0f);
  CK(cudaMemset(aux_out, 0, (int64_t)M * N * 2));
  CK(cudaDeviceSynchronize());

  evlm_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.M = M; g.N = N; g.K = K;
  g.A = A; g.lda = lda; g.a_mn = c->a_mn;
  g.B = B; g.ldb = ldb; g.b_mn = c->b_mn;
  g.D = D; g.ldd = ldd; g.d_dtype = c->d_dtype;
  g.splits = c->splits;
  g.alpha = 1.f;
  const float scale = 1.0f / sqrtf((float)K);  // keep pre-activations O(1)
  if (c->epi == 1) {
    g.bias = bias; g.alpha = 0.125f; g.alpha_cols = 256; g.act = EVLM_ACT_QUICK_GELU; g.gate = gate; g.gate_mode = EVLM_GATE_PRE_ACT;
    g.aux_out = aux_out; g.ld_aux_out = N; g.residual = res32; g.ldr = N; g.res_dtype = EVLM_F32;
  } else if (c->epi == 2) {
    g.bias = bias; g.act = EVLM_ACT_GELU_ERF; g.gate = gate; g.gate_mode = EVLM_GATE_POST_ACT;
    g.residual = res16; g.ldr = N; g.res_dtype = EVLM_BF16;
  } else if (c->epi == 3) {
    g.epi_mode = EVLM_EPI_ACT_BACKWARD; g.act = EVLM_ACT_QUICK_GELU; g.gate = gate; g.gate_mode = EVLM_GATE_PRE_ACT;
    g.aux_in = aux_in; g.ld_aux_in = N; g.aux_out = aux_out; g.ld_aux_out = N;
  } else if (c->epi == 4) {
    g.epi_mode = EVLM_EPI_ACT_BACKWARD; g.act = EVLM_ACT_GELU_ERF; g.gate = gate; g.gate_mode = EVLM_GATE_POST_ACT;
    g.aux_in = aux_in; g.ld_aux_in = N; g.aux_out = aux_out; g.ld_aux_out = N;
  } else if (c->epi == 5) {
    g.bias = bias; g.act = EVLM_ACT_QUICK_GELU; g.gate = gate; g.gate_mode = EVLM_GATE_PRE_ACT; g.aux_out = aux_out; g.ld_aux_out = N;
  } else if (c->epi == 6) {
    g.bias = bias; g.residual = res32; g.ldr = N; g.res_dtype = EVLM_F32;
  } else if (c->epi == 7) {
    g.bias = bias; g.residual = res32; g.ldr = N; g.res_dtype = EVLM_F32; g.dropout_p = 0.1f; g.dropout_seed = 99; g.dropout_stream = 3;
  } else if (c->epi == 8) {
    g.bias = bias; g.act = EVLM_ACT_GELU_ERF; g.gate = gate; g.gate_mode = EVLM_GATE_POST_ACT; g.aux_out = aux_out; g.ld_aux_out = N;
  }
  (void)scale;

  int rc = evlm_gemm_bf16(&g, nullptr);
  if (rc) { printf("[%s] evlm_gemm_bf16 rc=%d\n", c->name, rc); return 3; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[%s] kernel failed: %s\n", c->name, cudaGetErrorString(e)); return 4; }

  dim3 blk(32, 8), grd((N + 31) / 32, (M + 7) / 8);
  naive_gemm<<<grd, blk>>>(A, lda, c->a_mn, B, ldb, c->b_mn, Dref, M, N, K);
  CK(cudaDeviceSynchronize());

  // compare on host (sampled for the big cases)
  std::vector<float> href((size_t)M * N);
  CK(cudaMemcpy(href.data(), Dref, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
  std::vector<float> hD32; std::vector<__nv_bfloat16> hD16;
  if (c->d_dtype == EVLM_F32) { hD32.resize((size_t)M * ldd); CK(cudaMemcpy(hD32.data(), D, dsz, cudaMemcpyDeviceToHost)); }
  else { hD16.resize((size_t)M * ldd); CK(cudaMemcpy(hD16.data(), D, dsz, cudaMemcpyDeviceToHost)); }
  std::vector<float> hbias(N), hgate(N), hres32; std::vector<__nv_bfloat16> hres16, hauxin, hauxout;
  CK(cudaMemcpy(hbias.data(), bias, N * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hgate.data(), gate, N * 4, cudaMemcpyDeviceToHost));
  if (c->epi) {
    hres32.resize((size_t)M * N); hres16.resize((size_t)M * N); hauxin.resize((size_t)M * N); hauxout.resize((size_t)M * N);
    CK(cudaMemcpy(hres32.data(), res32, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hres16.data(), res16, (size_t)M * N * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hauxin.data(), aux_in, (size_t)M * N * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hauxout.data(), aux_out, (size_t)M * N * 2, cudaMemcpyDeviceToHost));
  }
  double max_err = 0, max_ref = 0, max_aux_err = 0;
  int64_t bad = 0, first_bad = -1;
  const int64_t total = (int64_t)M * N;
  const int64_t step = total > (1 << 24) ? 7 : 1;
  for (int64_t i = 0; i < total; i += step) {
    const int m = (int)(i / N), n = (int)(i % N);
    float acc = href[i], want = acc, want_aux = 0.f;
    if (c->epi == 1) {
      float v = acc + hbias[n];
      if (n < 256) v *= 0.125f;
      want_aux = v;
      v = qgelu(v * hgate[n]);
      want = v + hres32[i];
    } else if (c->epi == 2) {
      float v = gelu(acc + hbias[n]) * hgate[n];
      want = v + bf2f(hres16[i]);
    } else if (c->epi == 3) {
      float u = bf2f(hauxin[i]), z = hgate[n], d = qgelu_g(z * u);
      want = acc * d * z; want_aux = acc * d * u;
    } else if (c->epi == 4) {
      float u = bf2f(hauxin[i]), z = hgate[n];
      want = acc * z * gelu_g(u); want_aux = acc * gelu(u);
    } else if (c->epi == 5) {
      want_aux = acc + hbias[n];
      want = qgelu(want_aux * hgate[n]);
    } else if (c->epi == 6) {
      want = acc + hbias[n] + hres32[i];
    } else if (c->epi == 7) {   // dropout: kept elements are (acc + bias) / 0.9 + res, dropped ones res alone
      const float got7 = hD32[(size_t)m * ldd + n];
      const float keep = (acc + hbias[n]) / 0.9f + hres32[i], drop = hres32[i];
      want = fabs(got7 - keep) < fabs(got7 - drop) ? keep : drop;
    } else if (c->epi == 8) {
      want_aux = acc + hbias[n];
      want = gelu(want_aux) * hgate[n];
    }
    float got = c->d_dtype == EVLM_F32 ? hD32[(size_t)m * ldd + n] : bf2f(hD16[(size_t)m * ldd + n]);
    double err = fabs((double)got - want);
    double tol = (c->d_dtype == EVLM_F32 ? 2e-3 : 1.2e-2) * (fabs(want) + 1.0) + (c->splits > 1 ? 1e-3 : 0);
    if (err > tol) { if (first_bad < 0) first_bad = i; ++bad; }
    if (err > max_err) max_err = err;
    if (fabs(want) > max_ref) max_ref = fabs(want);
    if (c->epi == 1 || c->epi == 3 || c->epi == 4 || c->epi == 5 || c->epi == 8) {
      double ea = fabs((double)bf2f(hauxout[i]) - want_aux);
      if (ea > 1.2e-2 * (fabs(want_aux) + 1.0)) { if (first_bad < 0) first_bad = i; ++bad; }
      if (ea > max_aux_err) max_aux_err = ea;
    }
  }
  printf("[%s] M=%d N=%d K=%d a_mn=%d b_mn=%d splits=%d epi=%d : max_err=%.4g (max |ref|=%.4g) aux_err=%.4g bad=%lld %s\n", c->name, M,
         N, K, c->a_mn, c->b_mn, c->splits, c->epi, max_err, max_ref, max_aux_err, (long long)bad, bad ? "FAIL" : "PASS");
  if (bad) {
    const int m = (int)(first_bad / N), n = (int)(first_bad % N);
    float got = c->d_dtype == EVLM_F32 ? hD32[(size_t)m * ldd + n] : bf2f(hD16[(size_t)m * ldd + n]);
    printf("   first bad at (%d,%d): got %g acc_ref %g\n", m, n, got, href[first_bad]);
    // print a coarse map of bad 32x32 blocks for the first 256x256 corner to help diagnose layout bugs
    for (int bm = 0; bm < 8 && bm * 32 < M; ++bm) {
      printf("   ");
      for (int bn = 0; bn < 8 && bn * 32 < N; ++bn) {
        int cnt = 0;
        for (int i2 = 0; i2 < 32; ++i2)
          for (int j2 = 0; j2 < 32; ++j2) {
            int mm = bm * 32 + i2, nn = bn * 32 + j2;
            if (mm >= M || nn >= N) continue;
            float gg = c->d_dtype == EVLM_F32 ? hD32[(size_t)mm * ldd + nn] : bf2f(hD16[(size_t)mm * ldd + nn]);
            if (c->epi == 0 && fabs(gg - href[(size_t)mm * N + nn]) > 1.2e-2 * (fabs(href[(size_t)mm * N + nn]) + 1.0)) ++cnt;
          }
        printf("%5d", cnt);
      }
      printf("\n");
    }
  }

  if (c->time_it && !bad) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    if (c->splits > 1) g.accumulate = 1;
    for (int i = 0; i < 3; ++i) evlm_gemm_bf16(&g, nullptr);
    const int iters = 20;
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) evlm_gemm_bf16(&g, nullptr);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
    printf("[%s] %.3f ms  %.1f TFLOP/s\n", c->name, ms, 2.0 * M * N * K / ms * 1e-9);
  }
  return bad ? 5 : 0;
}
