// Shared device/host helpers for the leod_b200 kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <utility>

#include "../../include/leod_b200.h"

typedef __nv_bfloat16 bf16;

void leod_set_error(const char *fmt, ...);

#define LEOD_CUDA(expr)                                                                         \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      leod_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__));   \
      return -1;                                                                                \
    }                                                                                           \
  } while (0)

#define LEOD_REQUIRE(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      leod_set_error(__VA_ARGS__);     \
      return -2;                       \
    }                                  \
  } while (0)

extern unsigned long long g_leod_launches;  // kernels launched by this library (gpu_launches in bench.py)
#define LEOD_LAUNCH_CHECK()        \
  do {                             \
    ++g_leod_launches;             \
    LEOD_CUDA(cudaGetLastError()); \
  } while (0)

// ------------------------------------------------------------------ programmatic dependent launch
// Every kernel of this library is launched with cudaLaunchAttributeProgrammaticStreamSerialization (LEOD_PDL=0 disables): its CTAs
// may be scheduled while the previous kernel of the stream is still draining (the tail of a persistent GEMM, the last wave of an
// elementwise kernel), run their prologue (barrier init, TMEM allocation, tensor-map fetch) and then block in pdl_wait() until the
// predecessor has completed and its memory is visible.  Every kernel calls pdl_prologue() / pdl_wait() before its first access
// to global memory, reads AND writes, so the overlap never changes results.  The steps of this path are chains of several hundred
// short dependent kernels: the launch gap between them is the cost this removes.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  pdl_launch_dependents();
  pdl_wait();
}
extern int g_leod_pdl;
template <typename... KArgs, typename... Args>
inline cudaError_t leod_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_leod_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}
#define LEOD_LAUNCH(kernel, grid, block, smem, st, ...) (void)leod_launch(kernel, dim3(grid), dim3(block), (size_t)(smem), st, ##__VA_ARGS__)

// Optional per-kernel-class timing with CUDA events on the launching stream (bench.py roofline pass).
enum ProfKind { PK_GEMM_NT = 0, PK_GEMM_TN, PK_ATTN_FWD, PK_ATTN_BWD, PK_LAYERNORM, PK_LSTM, PK_PATCH, PK_OTHER, PK_CONV, PK_COUNT };
struct ProfScope {
  int slot;
  cudaStream_t st;
  ProfScope(int kind, double flops, double bytes, cudaStream_t st, int d0 = 0, int d1 = 0, int d2 = 0);
  ~ProfScope();
};

bool leod_profiling_on();

#define LEOD_TRY(expr)        \
  do {                        \
    int r__ = (expr);         \
    if (r__ != 0) return r__; \
  } while (0)

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<uint8_t>(uint8_t v) { return (float)v; }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
static inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
static inline size_t dtype_size(int dt) { return dt == LEOD_BF16 ? 2 : (dt == LEOD_U8 ? 1 : 4); }

// epilogue modes of the NT GEMM (see leod_b200.h)
enum { EPI_NONE = 0, EPI_GELU = 1, EPI_RESID = 2, EPI_GELU_BWD = 3 };

// Implicit-GEMM taps over a zero-bordered ("padded flat") channels-last matrix: a k x k stride-1 convolution is the sum
// over taps of A[row + off[tap], 0..cin) * B[n, tap*cinp .. tap*cinp + cin)^T (kernels_conv.cu explains the layout).
struct ConvTaps {
  int n = 0;      // taps (0 / 1 = plain GEMM)
  int cin = 0;    // A columns per tap
  int cinp = 0;   // B columns per tap (cin rounded up to a multiple of 64, zero filled)
  int off[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
};

struct GemmNT {
  const void *A; int lda;
  const void *A2; int lda2; int K1;
  const void *B; int ldb;
  void *C; int ldc;
  int M, N, K;
  const float *bias;
  int epi;
  const void *R; int ldr;
  void *aux; int ldaux;
  ConvTaps taps;       // taps.n > 1: K = taps.n * taps.cinp, A has taps.cin columns (A2 unused)
  int out_f32 = 0;     // C is fp32 whatever the operand type (EPI_NONE only)
};

// kernels_gemm_simt.cu
int gemm_nt_simt(int dtype, const GemmNT &g, cudaStream_t st);
int gemm_tn_simt(int dtype, const void *dY, int ldy, const void *X, int ldx, float *dW, int ldw, float *dbias, int M, int N,
                 int K, cudaStream_t st, const ConvTaps *taps = nullptr);
// kernels_gemm_tc.cu
int gemm_nt_tc(const GemmNT &g, cudaStream_t st);
int gemm_tn_tc(const void *dY, int ldy, const void *X, int ldx, float *dW, int ldw, float *dbias, int M, int N, int K,
               cudaStream_t st, const ConvTaps *taps = nullptr);
int lstm_seq_fwd_tc(void *gates, const void *Wh, int ldw, const void *h0, const void *c0, void *h_all, void *c_all, unsigned *flags,
                    int M, int C, int L, cudaStream_t st);
int lstm_seq_bwd_tc(const void *gates, const void *c_all, const void *c0, const void *dh, const void *dc_last, void *dgates, void *dc_ws,
                    void *dc0, void *dh0, const void *WhT, int ldw, unsigned *flags, int M, int C, int L, cudaStream_t st);
// kernels_attention.cu
int attention_fwd(int dtype, const void *qkv, void *out, int B, int H, int W, int C, int dh, int ph, int pw, int window,
                  cudaStream_t st);
int attention_bwd(int dtype, const void *qkv, const void *dout, void *dqkv, int B, int H, int W, int C, int dh, int ph,
                  int pw, int window, cudaStream_t st);
// kernels_attention_tc.cu (bf16, mma.sync): return 1 when the shape has no instantiation
int attention_fwd_tc(const void *qkv, void *out, int B, int H, int W, int C, int dh, int ph, int pw, int window, cudaStream_t st);
int attention_bwd_tc(const void *qkv, const void *dout, void *dqkv, int B, int H, int W, int C, int dh, int ph, int pw, int window,
                     cudaStream_t st);
extern int g_attention_force_simt;
// kernels_elem.cu
int layernorm_fwd(int dtype, const void *x, const float *w, const float *b, void *y, int M, int C, float eps, cudaStream_t st);
// dx = (dres ? dres : 0) + LN'(dy); dw += ..., db += ...
int layernorm_bwd(int dtype, const void *x, const float *w, const void *dy, const void *dres, void *dx, float *dw, float *db,
                  int M, int C, float eps, cudaStream_t st);
int lstm_pointwise_fwd(int dtype, void *gates /*in: pre-act, out: activated*/, const void *c_prev, void *h_out, void *c_out,
                       int M, int C, cudaStream_t st);
int lstm_pointwise_bwd(int dtype, const void *gates, const void *c_prev, const void *c_out, const void *dh, const void *dh2,
                       const void *dc, void *dgates, void *dc_prev, int M, int C, cudaStream_t st, const void *dh3 = nullptr);
int device_copy(void *dst, const void *src, size_t bytes, cudaStream_t st);   // SM-side (not copy-engine) copy
int device_zero_u32(unsigned *p, int64_t n, cudaStream_t st);
int add_tensors(int dtype, const void *a, const void *b, void *out, int64_t n, cudaStream_t st);
int im2col_nchw(int x_dtype, int dtype, const void *x, void *col, int B, int Cin, int xh, int xw, int Hp, int Wp, int ksz,
                int stride, int pad, int ldcol, cudaStream_t st);
int im2col_nhwc(int dtype, const void *x, void *col, int B, int Hi, int Wi, int Cin, int ksz, int stride, int pad, int ldcol,
                cudaStream_t st);
int col2im_nhwc(int dtype, const void *dcol, int ldcol, const void *dres, void *dx, int B, int Hi, int Wi, int Cin, int ksz,
                int stride, int pad, cudaStream_t st);
// weight preparation: dst[n, k] = scale[n] * src(n, k)  (+ transposed copy dstT[k, n])
//   perm: 0 = src is [N, K] row-major;  1 = src is conv weight [N, Cin, kh, kw], k index = (ky*kw + kx)*Cin + cin
// batched weight preparation / LayerScale gradient finalisation (kernels_elem.cu): items are passed in the kernel parameters
constexpr int PREP_BATCH_MAX = 30, LS_BATCH_MAX = 32;
struct PrepItem {
  const float *src, *scale;
  void *dst, *dstT;
  float *bias_out;     // != NULL: a scaled bias (out = src * scale, N elements) instead of a weight matrix
  int ldd, lddT, N, K, perm, Cin, ksz, first_block;
};
struct PrepBatch { PrepItem it[PREP_BATCH_MAX]; int n = 0; };
struct LsItem {
  float *G, *s;
  const float *W, *b, *gamma;
  float *dW, *db, *dgamma;
  int N, K, first_block, pad_;
};
struct LsBatch { LsItem it[LS_BATCH_MAX]; int n = 0; };
int prep_batch_add(PrepBatch &pb, int &blocks, const float *src, const float *scale, void *dst, int ldd, void *dstT, int lddT, int N, int K,
                   int perm, int Cin, int ksz, float *bias_out);
int prep_batch_launch(int dtype, const PrepBatch &pb, int blocks, cudaStream_t st);
int ls_batch_add(LsBatch &lb, int &blocks, float *G, float *s, const float *W, const float *b, const float *gamma, float *dW, float *db,
                 float *dgamma, int N, int K);
int ls_batch_launch(const LsBatch &lb, int blocks, cudaStream_t st);
int prep_weight(int dtype, const float *src, const float *scale, void *dst, int ldd, void *dstT, int lddT, int N, int K, int perm,
                int Cin, int ksz, cudaStream_t st);
int prep_scaled_bias(const float *b, const float *scale, float *out, int N, cudaStream_t st);
int layerscale_grad_finalize(const float *G, const float *s, const float *W, const float *b, const float *gamma, float *dW,
                             float *db, float *dgamma, int N, int K, cudaStream_t st);

// ------------------------------------------------------------------ neck / head building blocks (kernels_conv.cu)
// "Padded flat" geometry of one pyramid level: (h+2)*(w+2) rows per image, zero border.
struct PadGeom {
  int B, h, w, w2, P;
  int64_t R;   // B * P rows
};
PadGeom make_pad_geom(int B, int h, int w);
// One BatchNorm parameter segment (the fused twin convolutions of a CSP layer / head tower have two).
struct BnSeg {
  const float *gamma = nullptr, *beta = nullptr;
  float *rmean = nullptr, *rvar = nullptr;
  long long *nbt = nullptr;
  float *dgamma = nullptr, *dbeta = nullptr;
};
int pad_gather(int dtype, const void *src, void *dst, int ldd, const PadGeom &g, int C, cudaStream_t st);
int pad_scatter(int dtype, const void *src, int lds, void *dst, const PadGeom &g, int C, cudaStream_t st);
int upsample2x(int dtype, const void *src, int lds, const PadGeom &gs, void *dst, int ldd, const PadGeom &gd, int C, cudaStream_t st);
int upsample2x_bwd(int dtype, const void *ddst, int ldd, const PadGeom &gd, void *dsrc, int lds, const PadGeom &gs, int C, int accumulate,
                   cudaStream_t st);
int im2col_pad_s2(int dtype, const void *src, int lds, const PadGeom &gs, void *col, int ldcol, const PadGeom &go, int cin, int cinp,
                  cudaStream_t st);
int col2im_pad_s2(int dtype, const void *dcol, int ldcol, const PadGeom &go, void *dsrc, int lds, const PadGeom &gs, int cin, int cinp,
                  int accumulate, cudaStream_t st);
// One BatchNorm layer's buffers.  stats (double): [0,C) sum, [C,2C) sum of squares, [2C] count, [2C+1] ticket.  fin (float): mean, rstd.
// dloc / dglob (double): [0,C) sum dyhat, [C,2C) sum dyhat*xhat, [2C] ticket (local / over all ranks).  dfin (float): the two means.
struct BnLayer {
  int C = 0, cseg = 0;
  BnSeg s0, s1;
  double *stats = nullptr, *dloc = nullptr, *dglob = nullptr;
  float *fin = nullptr, *dfin = nullptr;
  float eps = 1e-5f, momentum = 0.1f;
};
int bn_stats(int dtype, const void *Y, int ldy, const PadGeom &g, const BnLayer &l, int finalize, cudaStream_t st);
int bn_finalize(const BnLayer &l, cudaStream_t st);
int bn_apply_silu(int dtype, const void *Y, int ldy, const PadGeom &g, const BnLayer &l, void *Z, int ldz, int training, cudaStream_t st);
int bn_bwd_reduce(int dtype, const void *dZ, int lddz, const void *Y, int ldy, const PadGeom &g, const BnLayer &l, int finalize, cudaStream_t st);
int bn_bwd_finalize(const BnLayer &l, cudaStream_t st);
int bn_bwd_apply(int dtype, const void *dZ, int lddz, const void *Y, int ldy, const PadGeom &g, const BnLayer &l, void *dY, int lddy, cudaStream_t st);
int device_zero_bytes(void *p, size_t bytes, cudaStream_t st);

// ------------------------------------------------------------------ head decode + SimOTA loss (kernels_simota.cu)
struct HeadGeom {   // three pyramid levels, anchors level-major / row-major (yolo_head.py:297-299)
  int B, A, C;
  int h[3], w[3], a0[3], stride[3], P[3];
};
struct HeadPtrs { const float *raw[3]; };
struct HeadGradPtrs { void *draw[3]; };
struct SimotaCfg {
  float ignore_label;
  int n_thresh;
  float thresh[8];
  float reg_w, obj_w, cls_w;
};
int head_decode(const HeadPtrs &rp, const HeadGeom &g, const SimotaCfg &cfg, float *preds, float *tout, const float *labels, int nmax,
                uint8_t *flags, int *match_cnt, int *match_gt, cudaStream_t st);
int simota_loss_fwd(const HeadGeom &g, const SimotaCfg &cfg, const float *tout, const float *labels, int nmax, const uint8_t *flags,
                    int *match_cnt, int *match_gt, int *assign, float *miou, double *sums, float *losses, cudaStream_t st);
int simota_loss_bwd(int dtype, const HeadGeom &g, const SimotaCfg &cfg, const float *tout, const float *labels, int nmax, const uint8_t *flags,
                    const int *assign, const float *miou, const double *sums, const float *gscale, const HeadGradPtrs &dp, cudaStream_t st);
int raw_grad_copy(int dtype, const HeadGeom &g, float *flat, const HeadGradPtrs &dp, int set, cudaStream_t st);

// ------------------------------------------------------------------ implicit-GEMM stem (kernels_stem.cu)
bool stem_implicit_supported(int Cin, int xh, int xw, int Ho, int Wo, int C, const void *x);
int stem_weight_to_f16(const void *W_bf16, void *W_f16, int64_t n, cudaStream_t st);
int stem_fwd_tc(const uint8_t *x, int nimg, int Cin, int xh, int xw, int Ho, int Wo, int C, const void *W, int ldw, void *y, cudaStream_t st);
int stem_wgrad_tc(const uint8_t *x, int nimg, int Cin, int xh, int xw, int Ho, int Wo, int C, const void *dY, float *dW, int ldw, cudaStream_t st);
