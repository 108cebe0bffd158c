// SIMT (CUDA-core, fp32 accumulate) GEMMs.  This is the exact-arithmetic path: it serves the
// LEOD_F32 mode (parity <= 1e-3 against the fp32 oracle) and cross-checks the tcgen05 kernels on
// the device.  The LEOD_BF16 product path uses kernels_gemm_tc.cu.
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256, PADW = 4;

template <typename T>
__device__ __forceinline__ void epilogue_store(const GemmNT &g, int m, int n, float v) {
  if (g.bias) v += g.bias[n];
  T *C = (T *)g.C;
  switch (g.epi) {
    case EPI_GELU:   // aux receives gelu'(v): the backward epilogue is then a plain multiply
      ((T *)g.aux)[(size_t)m * g.ldaux + n] = from_f<T>(gelu_grad_f(v));
      v = gelu_f(v);
      break;
    case EPI_RESID:
      v += to_f<T>(((const T *)g.R)[(size_t)m * g.ldr + n]);
      break;
    case EPI_GELU_BWD:
      v *= to_f<T>(((const T *)g.aux)[(size_t)m * g.ldaux + n]);
      break;
    default:
      break;
  }
  if (g.out_f32)
    ((float *)g.C)[(size_t)m * g.ldc + n] = v;
  else
    C[(size_t)m * g.ldc + n] = from_f<T>(v);
}

template <typename T>
__global__ void __launch_bounds__(NT) gemm_nt_simt_kernel(const GemmNT g) {
  pdl_prologue();
  __shared__ float As[BK][BM + PADW];
  __shared__ float Bs[BK][BN + PADW];
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const T *A = (const T *)g.A, *A2 = (const T *)g.A2, *B = (const T *)g.B;
  float acc[4][4] = {};
  const int lr = tid / 4, lk = (tid % 4) * 4;
  for (int k0 = 0; k0 < g.K; k0 += BK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + lk + j;
      const int m = m0 + lr, n = n0 + lr;
      float av = 0.f, bv = 0.f;
      if (k < g.K) {
        if (g.taps.n > 1) {   // implicit-GEMM taps: A row shifted per tap, rows outside [0, M) read as zero
          const int tap = k / g.taps.cinp, kc = k - tap * g.taps.cinp;
          const int r = m + g.taps.off[tap];
          if (m < g.M && kc < g.taps.cin && r >= 0 && r < g.M) av = to_f<T>(A[(size_t)r * g.lda + kc]);
        } else if (m < g.M) {
          av = (k < g.K1) ? to_f<T>(A[(size_t)m * g.lda + k]) : to_f<T>(A2[(size_t)m * g.lda2 + (k - g.K1)]);
        }
        if (n < g.N) bv = to_f<T>(B[(size_t)n * g.ldb + k]);
      }
      As[lk + j][lr] = av;
      Bs[lk + j][lr] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < g.N) epilogue_store<T>(g, m, n, acc[i][j]);
    }
  }
}

// dW[n,k] += sum_m dY[m,n] X[m,k];  grid = (tiles_n, tiles_k, splits over m)
template <typename T>
__global__ void __launch_bounds__(NT) gemm_tn_simt_kernel(const T *__restrict__ dY, int ldy, const T *__restrict__ X, int ldx,
                                                         float *__restrict__ dW, int ldw, float *__restrict__ dbias, int M,
                                                         int N, int K, int rows_per_split, const ConvTaps taps, int tiles_k) {
  pdl_prologue();
  __shared__ float Ys[BK][BN + PADW];
  __shared__ float Xs[BK][BN + PADW];
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const int tap = blockIdx.y / tiles_k;
  const int n0 = blockIdx.x * BN, k0 = (blockIdx.y % tiles_k) * BN;
  const int xoff = taps.n > 1 ? taps.off[tap] : 0;     // X row shift of this tap
  dW += taps.n > 1 ? (size_t)tap * taps.cinp : 0;
  const int m_begin = blockIdx.z * rows_per_split;
  const int m_end = min(M, m_begin + rows_per_split);
  float acc[4][4] = {};
  float colsum = 0.f;
  const bool do_bias = (dbias != nullptr) && blockIdx.y == 0;
  const int lm = tid / 16, lc = (tid % 16) * 4;
  for (int mb = m_begin; mb < m_end; mb += BK) {
    const int m = mb + lm;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float yv = 0.f, xv = 0.f;
      if (m < m_end) {
        if (n0 + lc + j < N) yv = to_f<T>(dY[(size_t)m * ldy + n0 + lc + j]);
        const int xr = m + xoff;
        if (k0 + lc + j < K && xr >= 0 && xr < M) xv = to_f<T>(X[(size_t)xr * ldx + k0 + lc + j]);
      }
      Ys[lm][lc + j] = yv;
      Xs[lm][lc + j] = xv;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < BK; ++r) {
      const float4 a4 = *reinterpret_cast<const float4 *>(&Ys[r][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4 *>(&Xs[r][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (do_bias && tid < BN) {
#pragma unroll
      for (int r = 0; r < BK; ++r) colsum += Ys[r][tid];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < K) atomicAdd(&dW[(size_t)n * ldw + k], acc[i][j]);
    }
  }
  if (do_bias && tid < BN && n0 + tid < N) atomicAdd(&dbias[n0 + tid], colsum);
}

}  // namespace

int gemm_nt_simt(int dtype, const GemmNT &g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return 0;
  dim3 grid(ceil_div(g.N, BN), ceil_div(g.M, BM));
  if (dtype == LEOD_F32)
    LEOD_LAUNCH((gemm_nt_simt_kernel<float>), grid, NT, 0, st, g);
  else
    LEOD_LAUNCH((gemm_nt_simt_kernel<bf16>), grid, NT, 0, st, g);
  LEOD_LAUNCH_CHECK();
  return 0;
}

int gemm_tn_simt(int dtype, const void *dY, int ldy, const void *X, int ldx, float *dW, int ldw, float *dbias, int M, int N,
                 int K, cudaStream_t st, const ConvTaps *taps) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  ConvTaps tp;
  if (taps) tp = *taps;
  const int ntap = tp.n > 1 ? tp.n : 1;
  const int tn = ceil_div(N, BN), tk = ceil_div(K, BN);
  int splits = ceil_div(148 * 4, tn * tk);
  const int max_splits = ceil_div(M, 4 * BK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int rows = (int)round_up(ceil_div(M, splits), BK);
  splits = ceil_div(M, rows);
  dim3 grid(tn, tk * ntap, splits);
  if (dtype == LEOD_F32)
    LEOD_LAUNCH((gemm_tn_simt_kernel<float>), grid, NT, 0, st, (const float *)dY, ldy, (const float *)X, ldx, dW, ldw, dbias, M, N, K, rows, tp, tk);
  else
    LEOD_LAUNCH((gemm_tn_simt_kernel<bf16>), grid, NT, 0, st, (const bf16 *)dY, ldy, (const bf16 *)X, ldx, dW, ldw, dbias, M, N, K, rows, tp, tk);
  LEOD_LAUNCH_CHECK();
  return 0;
}
