// Window / grid multi-head self-attention over channels-last token matrices, forward and backward.
// The partition (maxvit.py:273-304) is an index map applied while loading q/k/v and while storing the
// result, so the four permute+contiguous copies of the reference never exist.
// One CTA per (group, head); fp32 math in shared memory (SIMT version: exact path for LEOD_F32 and
// the bring-up path for LEOD_BF16 — the tensor-core version lives in kernels_attention_tc.cu).
#include "common.cuh"

namespace {

__device__ __forceinline__ int token_row(int g, int t, int H, int W, int ph, int pw, int window) {
  const int nh = H / ph, nw = W / pw;
  const int per_img = nh * nw;
  const int b = g / per_img, r = g % per_img;
  const int gy = r / nw, gx = r % nw;
  const int i = t / pw, j = t % pw;
  const int y = window ? gy * ph + i : i * nh + gy;
  const int x = window ? gx * pw + j : j * nw + gx;
  return (b * H + y) * W + x;
}

constexpr int ATT_THREADS = 128;

template <typename T>
__global__ void __launch_bounds__(ATT_THREADS) attn_fwd_kernel(const T *__restrict__ qkv, T *__restrict__ out, int H, int W, int C,
                                                               int dh, int ph, int pw, int window, float scale) {
  pdl_prologue();
  extern __shared__ float sm[];
  const int Tn = ph * pw, ldh = dh + 1, lds = Tn + 1;
  float *q = sm, *k = q + Tn * ldh, *v = k + Tn * ldh, *S = v + Tn * ldh;
  int *rows = (int *)(S + Tn * lds);
  const int g = blockIdx.x, head = blockIdx.y, tid = threadIdx.x;
  for (int t = tid; t < Tn; t += ATT_THREADS) rows[t] = token_row(g, t, H, W, ph, pw, window);
  __syncthreads();
  for (int idx = tid; idx < Tn * dh; idx += ATT_THREADS) {
    const int t = idx / dh, d = idx % dh;
    const T *src = qkv + (size_t)rows[t] * 3 * C + head * 3 * dh + d;
    q[t * ldh + d] = to_f<T>(src[0]) * scale;
    k[t * ldh + d] = to_f<T>(src[dh]);
    v[t * ldh + d] = to_f<T>(src[2 * dh]);
  }
  __syncthreads();
  for (int idx = tid; idx < Tn * Tn; idx += ATT_THREADS) {
    const int i = idx / Tn, j = idx % Tn;
    float a = 0.f;
    for (int d = 0; d < dh; ++d) a = fmaf(q[i * ldh + d], k[j * ldh + d], a);
    S[i * lds + j] = a;
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  for (int i = warp; i < Tn; i += ATT_THREADS / 32) {
    float mx = -INFINITY;
    for (int j = lane; j < Tn; j += 32) mx = fmaxf(mx, S[i * lds + j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < Tn; j += 32) {
      const float e = __expf(S[i * lds + j] - mx);
      S[i * lds + j] = e;
      sum += e;
    }
    const float inv = 1.f / warp_sum(sum);
    for (int j = lane; j < Tn; j += 32) S[i * lds + j] *= inv;
  }
  __syncthreads();
  for (int idx = tid; idx < Tn * dh; idx += ATT_THREADS) {
    const int i = idx / dh, d = idx % dh;
    float a = 0.f;
    for (int j = 0; j < Tn; ++j) a = fmaf(S[i * lds + j], v[j * ldh + d], a);
    out[(size_t)rows[i] * C + head * dh + d] = from_f<T>(a);
  }
}

template <typename T>
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_kernel(const T *__restrict__ qkv, const T *__restrict__ dout,
                                                               T *__restrict__ dqkv, int H, int W, int C, int dh, int ph, int pw,
                                                               int window, float scale) {
  pdl_prologue();
  extern __shared__ float sm[];
  const int Tn = ph * pw, ldh = dh + 1, lds = Tn + 1;
  float *q = sm, *k = q + Tn * ldh, *v = k + Tn * ldh, *dO = v + Tn * ldh;
  float *P = dO + Tn * ldh, *dS = P + Tn * lds;
  int *rows = (int *)(dS + Tn * lds);
  const int g = blockIdx.x, head = blockIdx.y, tid = threadIdx.x;
  for (int t = tid; t < Tn; t += ATT_THREADS) rows[t] = token_row(g, t, H, W, ph, pw, window);
  __syncthreads();
  for (int idx = tid; idx < Tn * dh; idx += ATT_THREADS) {
    const int t = idx / dh, d = idx % dh;
    const T *src = qkv + (size_t)rows[t] * 3 * C + head * 3 * dh + d;
    q[t * ldh + d] = to_f<T>(src[0]);
    k[t * ldh + d] = to_f<T>(src[dh]);
    v[t * ldh + d] = to_f<T>(src[2 * dh]);
    dO[t * ldh + d] = to_f<T>(dout[(size_t)rows[t] * C + head * dh + d]);
  }
  __syncthreads();
  for (int idx = tid; idx < Tn * Tn; idx += ATT_THREADS) {
    const int i = idx / Tn, j = idx % Tn;
    float a = 0.f, b = 0.f;
    for (int d = 0; d < dh; ++d) {
      a = fmaf(q[i * ldh + d], k[j * ldh + d], a);
      b = fmaf(dO[i * ldh + d], v[j * ldh + d], b);
    }
    P[i * lds + j] = a * scale;
    dS[i * lds + j] = b;  // dP
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  for (int i = warp; i < Tn; i += ATT_THREADS / 32) {
    float mx = -INFINITY;
    for (int j = lane; j < Tn; j += 32) mx = fmaxf(mx, P[i * lds + j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < Tn; j += 32) {
      const float e = __expf(P[i * lds + j] - mx);
      P[i * lds + j] = e;
      sum += e;
    }
    const float inv = 1.f / warp_sum(sum);
    float dsum = 0.f;
    for (int j = lane; j < Tn; j += 32) {
      const float p = P[i * lds + j] * inv;
      P[i * lds + j] = p;
      dsum += p * dS[i * lds + j];
    }
    dsum = warp_sum(dsum);
    for (int j = lane; j < Tn; j += 32) dS[i * lds + j] = P[i * lds + j] * (dS[i * lds + j] - dsum);
  }
  __syncthreads();
  for (int idx = tid; idx < Tn * dh; idx += ATT_THREADS) {
    const int t = idx / dh, d = idx % dh;
    float dq = 0.f, dk = 0.f, dv = 0.f;
    for (int j = 0; j < Tn; ++j) {
      dq = fmaf(dS[t * lds + j], k[j * ldh + d], dq);
      dk = fmaf(dS[j * lds + t], q[j * ldh + d], dk);
      dv = fmaf(P[j * lds + t], dO[j * ldh + d], dv);
    }
    T *dst = dqkv + (size_t)rows[t] * 3 * C + head * 3 * dh + d;
    dst[0] = from_f<T>(dq * scale);
    dst[dh] = from_f<T>(dk * scale);
    dst[2 * dh] = from_f<T>(dv);
  }
}

}  // namespace

int g_attention_force_simt = 0;

int attention_fwd(int dtype, const void *qkv, void *out, int B, int H, int W, int C, int dh, int ph, int pw, int window,
                  cudaStream_t st) {
  LEOD_REQUIRE(H % ph == 0 && W % pw == 0, "attention: %dx%d not divisible by partition %dx%d", H, W, ph, pw);
  LEOD_REQUIRE(C % dh == 0, "attention: C=%d not divisible by dim_head=%d", C, dh);
  if (dtype == LEOD_BF16 && !g_attention_force_simt) {
    const int rc = attention_fwd_tc(qkv, out, B, H, W, C, dh, ph, pw, window, st);
    if (rc <= 0) return rc;  // 1 = no tensor-core instantiation for this (tokens, dim_head)
  }
  const int Tn = ph * pw, groups = B * (H / ph) * (W / pw), heads = C / dh;
  const size_t smem = sizeof(float) * (3 * Tn * (dh + 1) + Tn * (Tn + 1)) + sizeof(int) * Tn;
  LEOD_REQUIRE(smem <= 200 * 1024, "attention: partition of %d tokens needs %zu B of shared memory", Tn, smem);
  const float scale = 1.0f / sqrtf((float)dh);
  dim3 grid(groups, heads);
  ProfScope ps(PK_ATTN_FWD, 4.0 * groups * heads * Tn * Tn * dh, 4.0 * B * H * W * C * dtype_size(dtype), st);
  if (dtype == LEOD_F32) {
    LEOD_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LEOD_LAUNCH((attn_fwd_kernel<float>), grid, ATT_THREADS, smem, st, (const float *)qkv, (float *)out, H, W, C, dh, ph, pw, window, scale);
  } else {
    LEOD_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LEOD_LAUNCH((attn_fwd_kernel<bf16>), grid, ATT_THREADS, smem, st, (const bf16 *)qkv, (bf16 *)out, H, W, C, dh, ph, pw, window, scale);
  }
  LEOD_LAUNCH_CHECK();
  return 0;
}

int attention_bwd(int dtype, const void *qkv, const void *dout, void *dqkv, int B, int H, int W, int C, int dh, int ph, int pw,
                  int window, cudaStream_t st) {
  LEOD_REQUIRE(H % ph == 0 && W % pw == 0, "attention: %dx%d not divisible by partition %dx%d", H, W, ph, pw);
  LEOD_REQUIRE(C % dh == 0, "attention: C=%d not divisible by dim_head=%d", C, dh);
  if (dtype == LEOD_BF16 && !g_attention_force_simt) {
    const int rc = attention_bwd_tc(qkv, dout, dqkv, B, H, W, C, dh, ph, pw, window, st);
    if (rc <= 0) return rc;
  }
  const int Tn = ph * pw, groups = B * (H / ph) * (W / pw), heads = C / dh;
  const size_t smem = sizeof(float) * (4 * Tn * (dh + 1) + 2 * Tn * (Tn + 1)) + sizeof(int) * Tn;
  LEOD_REQUIRE(smem <= 200 * 1024, "attention: partition of %d tokens needs %zu B of shared memory", Tn, smem);
  const float scale = 1.0f / sqrtf((float)dh);
  dim3 grid(groups, heads);
  ProfScope ps(PK_ATTN_BWD, 10.0 * groups * heads * Tn * Tn * dh, 7.0 * B * H * W * C * dtype_size(dtype), st);
  if (dtype == LEOD_F32) {
    LEOD_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LEOD_LAUNCH((attn_bwd_kernel<float>), grid, ATT_THREADS, smem, st, (const float *)qkv, (const float *)dout, (float *)dqkv, H, W, C, dh,
                                                           ph, pw, window, scale);
  } else {
    LEOD_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LEOD_LAUNCH((attn_bwd_kernel<bf16>), grid, ATT_THREADS, smem, st, (const bf16 *)qkv, (const bf16 *)dout, (bf16 *)dqkv, H, W, C, dh, ph, pw,
                                                          window, scale);
  }
  LEOD_LAUNCH_CHECK();
  return 0;
}
