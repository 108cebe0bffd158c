// Tensor-core window / grid attention for bf16 activations: one CTA per (group, head), one warp per
// 16 query rows, QK^T and PV (and the five backward products) on mma.sync.m16n8k16 with fp32
// accumulation, softmax in registers.  The per-(window, head) problems are 80x80x24 / 60x60x32 —
// far below a tcgen05 128-row tile — so the warp-level MMA is the right granularity here; the
// partition index map is applied while staging q/k/v, so no permute copies exist.
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

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&v);
}
// A fragment of the 16x16 block at (row0, k0) of a row-major [m][k] bf16 matrix with pitch ld
__device__ __forceinline__ void load_a(uint32_t (&a)[4], const bf16 *A, int ld, int row0, int k0, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const bf16 *p = A + (size_t)(row0 + g) * ld + k0 + 2 * t;
  a[0] = *reinterpret_cast<const uint32_t *>(p);
  a[1] = *reinterpret_cast<const uint32_t *>(p + 8 * ld);
  a[2] = *reinterpret_cast<const uint32_t *>(p + 8);
  a[3] = *reinterpret_cast<const uint32_t *>(p + 8 * ld + 8);
}
// B fragment (k16 x n8) from a matrix stored [n][k] (k contiguous) with pitch ld
__device__ __forceinline__ void load_b(uint32_t &b0, uint32_t &b1, const bf16 *Bm, int ld, int n0, int k0, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const bf16 *p = Bm + (size_t)(n0 + g) * ld + k0 + 2 * t;
  b0 = *reinterpret_cast<const uint32_t *>(p);
  b1 = *reinterpret_cast<const uint32_t *>(p + 8);
}

// NT_S: 8-column tiles of the score matrix (even); NT_O: dh / 8; KS: ceil(dh / 16)
template <int NT_S, int NT_O, int KS>
__global__ void __launch_bounds__(32 * ((NT_S * 8 + 15) / 16))
    attn_fwd_tc_kernel(const bf16 *__restrict__ qkv, bf16 *__restrict__ out, int H, int W, int C, int ph, int pw, int window,
                       float scale) {
  constexpr int TP = ((NT_S * 8 + 15) / 16) * 16;  // padded tokens
  constexpr int DH = NT_O * 8, DHP = KS * 16;
  constexpr int LDQ = DHP + 8, LDT = TP + 8;
  constexpr int NTHREADS = 32 * (TP / 16);
  __shared__ __align__(16) bf16 sQ[TP * LDQ];
  __shared__ __align__(16) bf16 sK[TP * LDQ];
  __shared__ __align__(16) bf16 sVt[DH * LDT];
  __shared__ int rows[TP];
  const int T = ph * pw;
  const int g = blockIdx.x, head = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < TP * LDQ / 2; i += NTHREADS) {
    reinterpret_cast<uint32_t *>(sQ)[i] = 0u;
    reinterpret_cast<uint32_t *>(sK)[i] = 0u;
  }
  for (int i = tid; i < DH * LDT / 2; i += NTHREADS) reinterpret_cast<uint32_t *>(sVt)[i] = 0u;
  for (int t = tid; t < TP; t += NTHREADS) rows[t] = t < T ? token_row(g, t, H, W, ph, pw, window) : 0;
  __syncthreads();
  constexpr int VPR = DH / 8;  // 16-byte vectors per q/k/v row
  for (int idx = tid; idx < T * 3 * VPR; idx += NTHREADS) {
    const int t = idx / (3 * VPR), rem = idx % (3 * VPR), which = rem / VPR, v8 = rem % VPR;
    const uint4 val = *reinterpret_cast<const uint4 *>(qkv + (size_t)rows[t] * 3 * C + head * 3 * DH + which * DH + v8 * 8);
    if (which == 0) {
      *reinterpret_cast<uint4 *>(&sQ[t * LDQ + v8 * 8]) = val;
    } else if (which == 1) {
      *reinterpret_cast<uint4 *>(&sK[t * LDQ + v8 * 8]) = val;
    } else {
      const bf16 *e = reinterpret_cast<const bf16 *>(&val);
#pragma unroll
      for (int j = 0; j < 8; ++j) sVt[(v8 * 8 + j) * LDT + t] = e[j];
    }
  }
  __syncthreads();
  const int row0 = warp * 16;
  float s[NT_S][4];
#pragma unroll
  for (int n = 0; n < NT_S; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    uint32_t a[4];
    load_a(a, sQ, LDQ, row0, ks * 16, lane);
#pragma unroll
    for (int n = 0; n < NT_S; ++n) {
      uint32_t b0, b1;
      load_b(b0, b1, sK, LDQ, n * 8, ks * 16, lane);
      mma16816(s[n], a, b0, b1);
    }
  }
  // softmax over the row (rows g and g+8 of this warp's slab); columns >= T are padding
  const int tq = lane & 3;
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int n = 0; n < NT_S; ++n) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const bool ok = n * 8 + 2 * tq + e < T;
      s[n][e] = ok ? s[n][e] * scale : -INFINITY;
      s[n][2 + e] = ok ? s[n][2 + e] * scale : -INFINITY;
      mx0 = fmaxf(mx0, s[n][e]);
      mx1 = fmaxf(mx1, s[n][2 + e]);
    }
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int n = 0; n < NT_S; ++n) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      s[n][e] = __expf(s[n][e] - mx0);
      s[n][2 + e] = __expf(s[n][2 + e] - mx1);
      sum0 += s[n][e];
      sum1 += s[n][2 + e];
    }
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
  float o[NT_O][4];
#pragma unroll
  for (int n = 0; n < NT_O; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < TP / 16; ++kk) {
    uint32_t a[4];
    if (2 * kk + 1 < NT_S) {
      a[0] = pack2(s[2 * kk][0], s[2 * kk][1]);
      a[1] = pack2(s[2 * kk][2], s[2 * kk][3]);
      a[2] = pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      a[3] = pack2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
    } else {
      a[0] = pack2(s[2 * kk][0], s[2 * kk][1]);
      a[1] = pack2(s[2 * kk][2], s[2 * kk][3]);
      a[2] = a[3] = 0u;
    }
#pragma unroll
    for (int n = 0; n < NT_O; ++n) {
      uint32_t b0, b1;
      load_b(b0, b1, sVt, LDT, n * 8, kk * 16, lane);
      mma16816(o[n], a, b0, b1);
    }
  }
  const int gq = lane >> 2;
  const int r0 = row0 + gq, r1 = row0 + gq + 8;
#pragma unroll
  for (int n = 0; n < NT_O; ++n) {
    const int col = head * DH + n * 8 + 2 * tq;
    if (r0 < T) *reinterpret_cast<uint32_t *>(out + (size_t)rows[r0] * C + col) = pack2(o[n][0] * inv0, o[n][1] * inv0);
    if (r1 < T) *reinterpret_cast<uint32_t *>(out + (size_t)rows[r1] * C + col) = pack2(o[n][2] * inv1, o[n][3] * inv1);
  }
}

// Backward: recompute P, dP = dO V^T, dS = P o (dP - rowsum(P o dP)); then dQ = dS K, dK = dS^T Q,
// dV = P^T dO with the transposed operands staged in shared memory.
template <int NT_S, int NT_O, int KS>
__global__ void __launch_bounds__(32 * ((NT_S * 8 + 15) / 16))
    attn_bwd_tc_kernel(const bf16 *__restrict__ qkv, const bf16 *__restrict__ dout, bf16 *__restrict__ dqkv, int H, int W, int C,
                       int ph, int pw, int window, float scale) {
  constexpr int TP = ((NT_S * 8 + 15) / 16) * 16;
  constexpr int DH = NT_O * 8, DHP = KS * 16;
  constexpr int LDQ = DHP + 8, LDT = TP + 8;
  constexpr int NTHREADS = 32 * (TP / 16);
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  bf16 *sQ = reinterpret_cast<bf16 *>(smem_dyn);     // [TP][LDQ]
  bf16 *sK = sQ + TP * LDQ;                          // [TP][LDQ]
  bf16 *sV = sK + TP * LDQ;                          // [TP][LDQ]
  bf16 *sdO = sV + TP * LDQ;                         // [TP][LDQ]
  bf16 *sQt = sdO + TP * LDQ;                        // [DH][LDT]
  bf16 *sKt = sQt + DH * LDT;                        // [DH][LDT]
  bf16 *sdOt = sKt + DH * LDT;                       // [DH][LDT]
  bf16 *sdS = sdOt + DH * LDT;                       // [TP][LDT]
  bf16 *sdSt = sdS + TP * LDT;                       // [TP][LDT]
  bf16 *sPt = sdSt + TP * LDT;                       // [TP][LDT]
  int *rows = reinterpret_cast<int *>(sPt + TP * LDT);
  constexpr int TOTAL_BF16 = 4 * TP * LDQ + 3 * DH * LDT + 3 * TP * LDT;
  const int T = ph * pw;
  const int g = blockIdx.x, head = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < TOTAL_BF16 / 2; i += NTHREADS) reinterpret_cast<uint32_t *>(smem_dyn)[i] = 0u;
  for (int t = tid; t < TP; t += NTHREADS) rows[t] = t < T ? token_row(g, t, H, W, ph, pw, window) : 0;
  __syncthreads();
  constexpr int VPR = DH / 8;
  for (int idx = tid; idx < T * 4 * VPR; idx += NTHREADS) {
    const int t = idx / (4 * VPR), rem = idx % (4 * VPR), which = rem / VPR, v8 = rem % VPR;
    uint4 val;
    if (which < 3)
      val = *reinterpret_cast<const uint4 *>(qkv + (size_t)rows[t] * 3 * C + head * 3 * DH + which * DH + v8 * 8);
    else
      val = *reinterpret_cast<const uint4 *>(dout + (size_t)rows[t] * C + head * DH + v8 * 8);
    bf16 *dst = which == 0 ? sQ : (which == 1 ? sK : (which == 2 ? sV : sdO));
    *reinterpret_cast<uint4 *>(&dst[t * LDQ + v8 * 8]) = val;
    bf16 *dstT = which == 0 ? sQt : (which == 1 ? sKt : (which == 3 ? sdOt : nullptr));
    if (dstT) {
      const bf16 *e = reinterpret_cast<const bf16 *>(&val);
#pragma unroll
      for (int j = 0; j < 8; ++j) dstT[(v8 * 8 + j) * LDT + t] = e[j];
    }
  }
  __syncthreads();
  const int row0 = warp * 16;
  const int tq = lane & 3, gq = lane >> 2;
  {
    float s[NT_S][4], dp[NT_S][4];
#pragma unroll
    for (int n = 0; n < NT_S; ++n) {
      s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
      dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
    }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t a[4], ad[4];
      load_a(a, sQ, LDQ, row0, ks * 16, lane);
      load_a(ad, sdO, LDQ, row0, ks * 16, lane);
#pragma unroll
      for (int n = 0; n < NT_S; ++n) {
        uint32_t b0, b1;
        load_b(b0, b1, sK, LDQ, n * 8, ks * 16, lane);
        mma16816(s[n], a, b0, b1);
        load_b(b0, b1, sV, LDQ, n * 8, ks * 16, lane);
        mma16816(dp[n], ad, b0, b1);
      }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < NT_S; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool ok = n * 8 + 2 * tq + e < T;
        s[n][e] = ok ? s[n][e] * scale : -INFINITY;
        s[n][2 + e] = ok ? s[n][2 + e] * scale : -INFINITY;
        mx0 = fmaxf(mx0, s[n][e]);
        mx1 = fmaxf(mx1, s[n][2 + e]);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int n = 0; n < NT_S; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[n][e] = __expf(s[n][e] - mx0);
        s[n][2 + e] = __expf(s[n][2 + e] - mx1);
        sum0 += s[n][e];
        sum1 += s[n][2 + e];
      }
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int n = 0; n < NT_S; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[n][e] *= inv0;
        s[n][2 + e] *= inv1;
        d0 += s[n][e] * dp[n][e];
        d1 += s[n][2 + e] * dp[n][2 + e];
      }
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
    d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    const int r0 = row0 + gq, r1 = row0 + gq + 8;
#pragma unroll
    for (int n = 0; n < NT_S; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = n * 8 + 2 * tq + e;
        if (col < TP) {
          const float p0 = s[n][e], p1 = s[n][2 + e];
          const bf16 ds0 = __float2bfloat16_rn(p0 * (dp[n][e] - d0) * scale);
          const bf16 ds1 = __float2bfloat16_rn(p1 * (dp[n][2 + e] - d1) * scale);
          sdS[r0 * LDT + col] = ds0;
          sdS[r1 * LDT + col] = ds1;
          sdSt[col * LDT + r0] = ds0;
          sdSt[col * LDT + r1] = ds1;
          sPt[col * LDT + r0] = __float2bfloat16_rn(p0);
          sPt[col * LDT + r1] = __float2bfloat16_rn(p1);
        }
      }
    }
  }
  __syncthreads();
  // second phase: this warp's 16 rows of dQ (rows = queries) and of dK, dV (rows = keys)
  float dq[NT_O][4], dk[NT_O][4], dv[NT_O][4];
#pragma unroll
  for (int n = 0; n < NT_O; ++n) {
    dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f;
    dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
    dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
  }
#pragma unroll
  for (int kk = 0; kk < TP / 16; ++kk) {
    uint32_t a1[4], a2[4], a3[4];
    load_a(a1, sdS, LDT, row0, kk * 16, lane);
    load_a(a2, sdSt, LDT, row0, kk * 16, lane);
    load_a(a3, sPt, LDT, row0, kk * 16, lane);
#pragma unroll
    for (int n = 0; n < NT_O; ++n) {
      uint32_t b0, b1;
      load_b(b0, b1, sKt, LDT, n * 8, kk * 16, lane);
      mma16816(dq[n], a1, b0, b1);
      load_b(b0, b1, sQt, LDT, n * 8, kk * 16, lane);
      mma16816(dk[n], a2, b0, b1);
      load_b(b0, b1, sdOt, LDT, n * 8, kk * 16, lane);
      mma16816(dv[n], a3, b0, b1);
    }
  }
  const int r0 = row0 + gq, r1 = row0 + gq + 8;
#pragma unroll
  for (int n = 0; n < NT_O; ++n) {
    const int col = head * 3 * DH + n * 8 + 2 * tq;
    if (r0 < T) {
      bf16 *dst = dqkv + (size_t)rows[r0] * 3 * C + col;
      *reinterpret_cast<uint32_t *>(dst) = pack2(dq[n][0], dq[n][1]);
      *reinterpret_cast<uint32_t *>(dst + DH) = pack2(dk[n][0], dk[n][1]);
      *reinterpret_cast<uint32_t *>(dst + 2 * DH) = pack2(dv[n][0], dv[n][1]);
    }
    if (r1 < T) {
      bf16 *dst = dqkv + (size_t)rows[r1] * 3 * C + col;
      *reinterpret_cast<uint32_t *>(dst) = pack2(dq[n][2], dq[n][3]);
      *reinterpret_cast<uint32_t *>(dst + DH) = pack2(dk[n][2], dk[n][3]);
      *reinterpret_cast<uint32_t *>(dst + 2 * DH) = pack2(dv[n][2], dv[n][3]);
    }
  }
}

template <int NT_S, int NT_O, int KS>
int launch_fwd(const bf16 *qkv, bf16 *out, int groups, int heads, int H, int W, int C, int ph, int pw, int window, float scale,
               cudaStream_t st) {
  constexpr int TP = ((NT_S * 8 + 15) / 16) * 16;
  attn_fwd_tc_kernel<NT_S, NT_O, KS><<<dim3(groups, heads), 32 * (TP / 16), 0, st>>>(qkv, out, H, W, C, ph, pw, window, scale);
  return 0;
}
template <int NT_S, int NT_O, int KS>
int launch_bwd(const bf16 *qkv, const bf16 *dout, bf16 *dqkv, int groups, int heads, int H, int W, int C, int ph, int pw, int window,
               float scale, cudaStream_t st) {
  constexpr int TP = ((NT_S * 8 + 15) / 16) * 16;
  constexpr int DH = NT_O * 8, DHP = KS * 16, LDQ = DHP + 8, LDT = TP + 8;
  constexpr size_t smem = 2 * (4 * TP * LDQ + 3 * DH * LDT + 3 * TP * LDT) + 4 * TP;
  static bool set = false;
  if (!set) {
    LEOD_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<NT_S, NT_O, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    set = true;
  }
  attn_bwd_tc_kernel<NT_S, NT_O, KS><<<dim3(groups, heads), 32 * (TP / 16), smem, st>>>(qkv, dout, dqkv, H, W, C, ph, pw, window, scale);
  return 0;
}

}  // namespace

// Returns 1 when the (tokens, dim_head) combination has no tensor-core instantiation (caller uses the
// SIMT kernel), 0 on success, < 0 on error.
int attention_fwd_tc(const void *qkv, void *out, int B, int H, int W, int C, int dh, int ph, int pw, int window, cudaStream_t st) {
  const int T = ph * pw, groups = B * (H / ph) * (W / pw), heads = C / dh;
  const float scale = 1.0f / sqrtf((float)dh);
  const bf16 *q = (const bf16 *)qkv;
  bf16 *o = (bf16 *)out;
  if (C % 8 != 0) return 1;
  ProfScope ps(PK_ATTN_FWD, 4.0 * groups * heads * T * T * dh, 4.0 * B * H * W * C * 2, st);
  if (T == 80 && dh == 24) launch_fwd<10, 3, 2>(q, o, groups, heads, H, W, C, ph, pw, window, scale, st);
  else if (T == 80 && dh == 32) launch_fwd<10, 4, 2>(q, o, groups, heads, H, W, C, ph, pw, window, scale, st);
  else if (T == 60 && dh == 32) launch_fwd<8, 4, 2>(q, o, groups, heads, H, W, C, ph, pw, window, scale, st);
  else if (T == 60 && dh == 24) launch_fwd<8, 3, 2>(q, o, groups, heads, H, W, C, ph, pw, window, scale, st);
  else return 1;
  LEOD_LAUNCH_CHECK();
  return 0;
}

int attention_bwd_tc(const void *qkv, const void *dout, void *dqkv, int B, int H, int W, int C, int dh, int ph, int pw, int window,
                     cudaStream_t st) {
  const int T = ph * pw, groups = B * (H / ph) * (W / pw), heads = C / dh;
  const float scale = 1.0f / sqrtf((float)dh);
  const bf16 *q = (const bf16 *)qkv, *d = (const bf16 *)dout;
  bf16 *o = (bf16 *)dqkv;
  if (C % 8 != 0) return 1;
  ProfScope ps(PK_ATTN_BWD, 10.0 * groups * heads * T * T * dh, 7.0 * B * H * W * C * 2, st);
  int rc;
  if (T == 80 && dh == 24) rc = launch_bwd<10, 3, 2>(q, d, o, groups, heads, H, W, C, ph, pw, window, scale, st);
  else if (T == 80 && dh == 32) rc = launch_bwd<10, 4, 2>(q, d, o, groups, heads, H, W, C, ph, pw, window, scale, st);
  else if (T == 60 && dh == 32) rc = launch_bwd<8, 4, 2>(q, d, o, groups, heads, H, W, C, ph, pw, window, scale, st);
  else if (T == 60 && dh == 24) rc = launch_bwd<8, 3, 2>(q, d, o, groups, heads, H, W, C, ph, pw, window, scale, st);
  else return 1;
  if (rc != 0) return rc;
  LEOD_LAUNCH_CHECK();
  return 0;
}
