// C-ABI glue: error reporting and the thin exported wrappers around the building-block kernels.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void leod_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

unsigned long long g_leod_launches = 0;
int g_leod_pdl = [] {
  const char *e = getenv("LEOD_PDL");     // 0: plain stream-ordered launches
  return e ? atoi(e) : 1;
}();

// ------------------------------------------------------------------ event profiler
#include <vector>
namespace {
struct ProfRec { int kind; cudaEvent_t e0, e1; double flops, bytes; int d[3]; };
FILE *g_prof_csv = nullptr;
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
}  // namespace

bool leod_profiling_on() { return g_prof_on; }

ProfScope::ProfScope(int kind, double flops, double bytes, cudaStream_t st_, int d0, int d1, int d2) : slot(-1), st(st_) {
  if (!g_prof_on) return;
  ProfRec r;
  r.kind = kind; r.flops = flops; r.bytes = bytes; r.d[0] = d0; r.d[1] = d1; r.d[2] = d2;
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
  cudaEventRecord(r.e0, st);
  g_prof.push_back(r);
  slot = (int)g_prof.size() - 1;
}
ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(g_prof[slot].e1, st);
}

extern "C" int leod_profile_enable(int on) {
  g_prof_on = on != 0;
  return 0;
}
// out[kind*4 + {0,1,2,3}] = {launches, total ms, algorithmic flops, algorithmic bytes}; clears the log.
extern "C" int leod_profile_collect(double *out, int n_kinds) {
  LEOD_REQUIRE(out && n_kinds >= PK_COUNT, "leod_profile_collect: need room for %d kinds", (int)PK_COUNT);
  for (int i = 0; i < n_kinds * 4; ++i) out[i] = 0.0;
  for (auto &r : g_prof) {
    float ms = 0.f;
    LEOD_CUDA(cudaEventSynchronize(r.e1));
    LEOD_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
    out[r.kind * 4 + 0] += 1;
    out[r.kind * 4 + 1] += ms;
    out[r.kind * 4 + 2] += r.flops;
    out[r.kind * 4 + 3] += r.bytes;
    if (g_prof_csv) fprintf(g_prof_csv, "%d,%.3f,%.0f,%.0f,%d,%d,%d\n", r.kind, ms * 1e3, r.flops, r.bytes, r.d[0], r.d[1], r.d[2]);
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_prof.clear();
  return 0;
}
// Also write one CSV line per launch (kind,us,flops,bytes,d0,d1,d2) at the next collect; NULL stops.
extern "C" int leod_profile_csv(const char *path) {
  if (g_prof_csv) fclose(g_prof_csv);
  g_prof_csv = path ? fopen(path, "w") : nullptr;
  return (path && !g_prof_csv) ? -1 : 0;
}
extern "C" int leod_debug_force_simt_attention(int on) {
  g_attention_force_simt = on;
  return 0;
}
extern "C" unsigned long long leod_launch_count(void) { return g_leod_launches; }

extern "C" const char *leod_last_error(void) { return g_err; }
extern "C" int leod_abi_version(void) { return LEOD_ABI_VERSION; }

extern "C" int leod_gemm_nt(int impl, int dtype, const void *A, int lda, const void *A2, int lda2, int K1, const void *B, int ldb,
                            void *C, int ldc, int M, int N, int K, const float *bias, int epi, const void *R, int ldr, void *aux,
                            int ldaux, void *stream) {
  LEOD_REQUIRE(A && B && C, "leod_gemm_nt: null operand");
  LEOD_REQUIRE(dtype == LEOD_F32 || dtype == LEOD_BF16, "leod_gemm_nt: dtype %d", dtype);
  LEOD_REQUIRE(epi >= 0 && epi <= 3, "leod_gemm_nt: epilogue %d", epi);
  GemmNT g;
  g.A = A; g.lda = lda; g.A2 = A2; g.lda2 = lda2; g.K1 = A2 ? K1 : K;
  g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K;
  g.bias = bias; g.epi = epi; g.R = R; g.ldr = ldr; g.aux = aux; g.ldaux = ldaux;
  if (impl == 1) {
    LEOD_REQUIRE(dtype == LEOD_BF16, "leod_gemm_nt: the tensor-core kernel takes bf16 operands");
    return gemm_nt_tc(g, (cudaStream_t)stream);
  }
  return gemm_nt_simt(dtype, g, (cudaStream_t)stream);
}

extern "C" int leod_gemm_tn(int impl, int dtype, const void *dY, int ldy, const void *X, int ldx, float *dW, int ldw, float *dbias,
                            int M, int N, int K, void *stream) {
  LEOD_REQUIRE(dY && X && dW, "leod_gemm_tn: null operand");
  if (impl == 1) {
    LEOD_REQUIRE(dtype == LEOD_BF16, "leod_gemm_tn: the tensor-core kernel takes bf16 operands");
    return gemm_tn_tc(dY, ldy, X, ldx, dW, ldw, dbias, M, N, K, (cudaStream_t)stream);
  }
  return gemm_tn_simt(dtype, dY, ldy, X, ldx, dW, ldw, dbias, M, N, K, (cudaStream_t)stream);
}

extern "C" int leod_stem_conv_fwd(const void *x, int nimg, int Cin, int xh, int xw, int Ho, int Wo, int C, const void *W_f16, int ldw,
                                  void *y, void *stream) {
  LEOD_REQUIRE(x && W_f16 && y, "leod_stem_conv_fwd: null operand");
  LEOD_REQUIRE(stem_implicit_supported(Cin, xh, xw, Ho, Wo, C, x), "leod_stem_conv_fwd: unsupported geometry (Cin %d, %dx%d -> %dx%d, C %d)", Cin,
               xh, xw, Ho, Wo, C);
  return stem_fwd_tc((const uint8_t *)x, nimg, Cin, xh, xw, Ho, Wo, C, W_f16, ldw, y, (cudaStream_t)stream);
}
extern "C" int leod_stem_conv_wgrad(const void *x, int nimg, int Cin, int xh, int xw, int Ho, int Wo, int C, const void *dY, float *dW,
                                    int ldw, void *stream) {
  LEOD_REQUIRE(x && dY && dW, "leod_stem_conv_wgrad: null operand");
  LEOD_REQUIRE(stem_implicit_supported(Cin, xh, xw, Ho, Wo, C, x), "leod_stem_conv_wgrad: unsupported geometry (Cin %d, %dx%d -> %dx%d, C %d)",
               Cin, xh, xw, Ho, Wo, C);
  return stem_wgrad_tc((const uint8_t *)x, nimg, Cin, xh, xw, Ho, Wo, C, dY, dW, ldw, (cudaStream_t)stream);
}

extern "C" int leod_attention_fwd(int dtype, const void *qkv, void *out, int B, int H, int W, int C, int dim_head, int ph, int pw,
                                  int window, void *stream) {
  LEOD_REQUIRE(qkv && out, "leod_attention_fwd: null operand");
  return attention_fwd(dtype, qkv, out, B, H, W, C, dim_head, ph, pw, window, (cudaStream_t)stream);
}

extern "C" int leod_attention_bwd(int dtype, const void *qkv, const void *dout, void *dqkv, int B, int H, int W, int C,
                                  int dim_head, int ph, int pw, int window, void *stream) {
  LEOD_REQUIRE(qkv && dout && dqkv, "leod_attention_bwd: null operand");
  return attention_bwd(dtype, qkv, dout, dqkv, B, H, W, C, dim_head, ph, pw, window, (cudaStream_t)stream);
}

extern "C" int leod_layernorm_fwd(int dtype, const void *x, const float *w, const float *b, void *y, int M, int C, float eps,
                                  void *stream) {
  LEOD_REQUIRE(x && w && b && y, "leod_layernorm_fwd: null operand");
  return layernorm_fwd(dtype, x, w, b, y, M, C, eps, (cudaStream_t)stream);
}
extern "C" int leod_layernorm_bwd(int dtype, const void *x, const float *w, const void *dy, const void *dres, void *dx, float *dw,
                                  float *db, int M, int C, float eps, void *stream) {
  LEOD_REQUIRE(x && w && dy && dx && dw && db, "leod_layernorm_bwd: null operand");
  return layernorm_bwd(dtype, x, w, dy, dres, dx, dw, db, M, C, eps, (cudaStream_t)stream);
}
extern "C" int leod_lstm_gates_fwd(int dtype, void *gates, const void *c_prev, void *h_out, void *c_out, int M, int C, void *stream) {
  LEOD_REQUIRE(gates && h_out && c_out, "leod_lstm_gates_fwd: null operand");
  return lstm_pointwise_fwd(dtype, gates, c_prev, h_out, c_out, M, C, (cudaStream_t)stream);
}
extern "C" int leod_lstm_gates_bwd(int dtype, const void *gates, const void *c_prev, const void *c_out, const void *dh,
                                   const void *dh2, const void *dc, void *dgates, void *dc_prev, int M, int C, void *stream) {
  LEOD_REQUIRE(gates && c_out && dgates && dc_prev, "leod_lstm_gates_bwd: null operand");
  return lstm_pointwise_bwd(dtype, gates, c_prev, c_out, dh, dh2, dc, dgates, dc_prev, M, C, (cudaStream_t)stream);
}
