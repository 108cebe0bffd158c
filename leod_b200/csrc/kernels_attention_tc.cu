// Tensor-core window / grid attention for bf16 activations: one CTA per (group, head), one warp per
// 16 query rows, QK^T and PV (and the five backward products) on mma.sync.m16n8k16 with fp32
// accumulation, softmax in registers.  The per-(window, head) problems are 80x80x24 / 60x60x32 —
// far below a tcgen05 128-row tile — so the warp-level MMA is the right granularity here; the
// partition index map is applied while staging q/k/v, so no permute copies exist.  Operands are staged
// row-major once; every transposed view (V for PV, K for dQ, dS^T / P^T / Q / dO for dK and dV) is a
// transposing ldmatrix, results leave through a shared-memory tile as 16-byte row pieces.
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
// ---- ldmatrix fragment loaders (m16n8k16, bf16).  All matrices live in shared memory row-major with a pitch that is a
// multiple of 16 bytes and conflict-free for 8 consecutive rows (80 B and 176 B pitches below).
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t &r0, uint32_t &r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t &r0, uint32_t &r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
// A fragment (16 rows x 16 k) of a matrix stored [row][k]
__device__ __forceinline__ void frag_a(uint32_t (&a)[4], const bf16 *M, int ld, int row0, int k0, int lane) {
  ldsm_x4(a, smem_addr(M + (size_t)(row0 + (lane & 7) + ((lane >> 3) & 1) * 8) * ld + k0 + (lane >> 4) * 8));
}
// A fragment of the TRANSPOSE of a matrix stored [k][row]: A[row0+i][k0+j] = M[k0+j][row0+i]
__device__ __forceinline__ void frag_a_t(uint32_t (&a)[4], const bf16 *M, int ld, int row0, int k0, int lane) {
  const int i = lane >> 3;
  ldsm_x4_t(a, smem_addr(M + (size_t)(k0 + (lane & 7) + (i >> 1) * 8) * ld + row0 + (i & 1) * 8));
}
// B fragment (16 k x 8 n) of a matrix stored [n][k] (k contiguous)
__device__ __forceinline__ void frag_b(uint32_t &b0, uint32_t &b1, const bf16 *M, int ld, int n0, int k0, int lane) {
  ldsm_x2(b0, b1, smem_addr(M + (size_t)(n0 + (lane & 7)) * ld + k0 + ((lane >> 3) & 1) * 8));
}
// B fragment of a matrix stored [k][n] (n contiguous)
__device__ __forceinline__ void frag_b_t(uint32_t &b0, uint32_t &b1, const bf16 *M, int ld, int n0, int k0, int lane) {
  ldsm_x2_t(b0, b1, smem_addr(M + (size_t)(k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * ld + n0));
}

template <int NT_S>
__device__ __forceinline__ void softmax_rows(float (&s)[NT_S][4], int T, float scale, int tq, float &inv0, float &inv1) {
  // rows g and g+8 of the warp's 16-row slab; columns >= T are padding.  exp via ex2 with log2(e) folded into the scale.
  const float sl = scale * 1.4426950408889634f;
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int n = 0; n < NT_S; ++n) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const bool ok = n * 8 + 2 * tq + e < T;
      s[n][e] = ok ? s[n][e] * sl : -INFINITY;
      s[n][2 + e] = ok ? s[n][2 + e] * sl : -INFINITY;
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
      s[n][e] = exp2f(s[n][e] - mx0);
      s[n][2 + e] = exp2f(s[n][2 + e] - mx1);
      sum0 += s[n][e];
      sum1 += s[n][2 + e];
    }
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  inv0 = 1.f / sum0;
  inv1 = 1.f / sum1;
}

// Stage the [q|k|v] rows (and optionally dO) of one (group, head) into shared memory with 16-byte copies; the k-padding
// columns DH..DHP-1 and the token-padding rows T..TP-1 are zero.
template <int TP, int DH, int DHP, int LDQ, int NTHREADS, bool WITH_DO>
__device__ __forceinline__ void stage_qkv(const bf16 *__restrict__ qkv, const bf16 *__restrict__ dout, bf16 *sQ, bf16 *sK, bf16 *sV,
                                          bf16 *sdO, const int *rows, int T, int C, int head, int tid) {
  constexpr int VPR = DH / 8, VPP = DHP / 8, NM = WITH_DO ? 4 : 3;
  // asynchronous 16-byte copies (zero-filled where padded): every load of the CTA is in flight at once instead of one
  // dependent load -> store pair per loop iteration
  for (int idx = tid; idx < TP * NM * VPP; idx += NTHREADS) {
    const int t = idx / (NM * VPP), rem = idx % (NM * VPP), which = rem / VPP, v8 = rem % VPP;
    const bool ok = t < T && v8 < VPR;
    const bf16 *src = qkv;
    if (ok) src = which < 3 ? qkv + (size_t)rows[t] * 3 * C + head * 3 * DH + which * DH + v8 * 8
                            : dout + (size_t)rows[t] * C + head * DH + v8 * 8;
    bf16 *dst = which == 0 ? sQ : (which == 1 ? sK : (which == 2 ? sV : sdO));
    const int sz = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr(&dst[t * LDQ + v8 * 8])), "l"(src), "r"(sz) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// NT_S: 8-column tiles of the score matrix; NT_O: dh / 8; KS: ceil(dh / 16)
template <int NT_S, int NT_O, int KS>
__global__ void __launch_bounds__(32 * ((NT_S * 8 + 15) / 16))
    attn_fwd_tc_kernel(const bf16 *__restrict__ qkv, bf16 *__restrict__ out, int H, int W, int C, int ph, int pw, int window,
                       float scale) {
  pdl_prologue();
  constexpr int TP = ((NT_S * 8 + 15) / 16) * 16;  // padded tokens
  constexpr int DH = NT_O * 8, DHP = KS * 16;
  constexpr int LDQ = DHP + 8;
  constexpr int NTHREADS = 32 * (TP / 16);
  __shared__ __align__(16) bf16 sQ[TP * LDQ];   // reused as the output staging tile
  __shared__ __align__(16) bf16 sK[TP * LDQ];
  __shared__ __align__(16) bf16 sV[TP * LDQ];
  __shared__ int rows[TP];
  const int T = ph * pw;
  const int g = blockIdx.x, head = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int t = tid; t < TP; t += NTHREADS) rows[t] = t < T ? token_row(g, t, H, W, ph, pw, window) : 0;
  __syncthreads();
  stage_qkv<TP, DH, DHP, LDQ, NTHREADS, false>(qkv, nullptr, sQ, sK, sV, nullptr, rows, T, C, head, tid);
  __syncthreads();
  const int row0 = warp * 16;
  float s[NT_S][4];
#pragma unroll
  for (int n = 0; n < NT_S; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    uint32_t a[4];
    frag_a(a, sQ, LDQ, row0, ks * 16, lane);
#pragma unroll
    for (int n = 0; n < NT_S; ++n) {
      uint32_t b0, b1;
      frag_b(b0, b1, sK, LDQ, n * 8, ks * 16, lane);
      mma16816(s[n], a, b0, b1);
    }
  }
  const int tq = lane & 3, gq = lane >> 2;
  float inv0, inv1;
  softmax_rows<NT_S>(s, T, scale, tq, inv0, inv1);
  float o[NT_O][4];
#pragma unroll
  for (int n = 0; n < NT_O; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < TP / 16; ++kk) {
    uint32_t a[4];
    a[0] = pack2(s[2 * kk][0], s[2 * kk][1]);
    a[1] = pack2(s[2 * kk][2], s[2 * kk][3]);
    if (2 * kk + 1 < NT_S) {
      a[2] = pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      a[3] = pack2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
    } else {
      a[2] = a[3] = 0u;
    }
#pragma unroll
    for (int n = 0; n < NT_O; ++n) {
      uint32_t b0, b1;
      frag_b_t(b0, b1, sV, LDQ, n * 8, kk * 16, lane);   // V is [token][dh]: transposed load gives the [k=token][n=dh] operand
      mma16816(o[n], a, b0, b1);
    }
  }
  // stage the warp's 16 output rows in its own (now dead) slab of sQ, then write 16-byte pieces
  __syncwarp();
#pragma unroll
  for (int n = 0; n < NT_O; ++n) {
    *reinterpret_cast<uint32_t *>(&sQ[(row0 + gq) * LDQ + n * 8 + 2 * tq]) = pack2(o[n][0] * inv0, o[n][1] * inv0);
    *reinterpret_cast<uint32_t *>(&sQ[(row0 + gq + 8) * LDQ + n * 8 + 2 * tq]) = pack2(o[n][2] * inv1, o[n][3] * inv1);
  }
  __syncwarp();
  constexpr int VPR = DH / 8;
  for (int idx = lane; idx < 16 * VPR; idx += 32) {
    const int r = row0 + idx / VPR, v8 = idx % VPR;
    if (r < T) *reinterpret_cast<uint4 *>(out + (size_t)rows[r] * C + head * DH + v8 * 8) = *reinterpret_cast<const uint4 *>(&sQ[r * LDQ + v8 * 8]);
  }
}

// Backward: recompute P, dP = dO V^T, dS = P o (dP - rowsum(P o dP)) * scale; dQ = dS K from registers; P and dS go
// to shared memory once and dK = dS^T Q, dV = P^T dO read them (and Q, dO) through transposing ldmatrix loads.
template <int NT_S, int NT_O, int KS>
__global__ void __launch_bounds__(32 * ((NT_S * 8 + 15) / 16), 4)
    attn_bwd_tc_kernel(const bf16 *__restrict__ qkv, const bf16 *__restrict__ dout, bf16 *__restrict__ dqkv, int H, int W, int C,
                       int ph, int pw, int window, float scale) {
  pdl_prologue();
  constexpr int TP = ((NT_S * 8 + 15) / 16) * 16;
  constexpr int DH = NT_O * 8, DHP = KS * 16;
  constexpr int LDQ = DHP + 8, LDT = TP + 8, LDO = 3 * DH + 8;
  constexpr int NTHREADS = 32 * (TP / 16);
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  bf16 *sQ = reinterpret_cast<bf16 *>(smem_dyn);     // [TP][LDQ]
  bf16 *sK = sQ + TP * LDQ;
  bf16 *sV = sK + TP * LDQ;
  bf16 *sdO = sV + TP * LDQ;
  bf16 *sdS = sdO + TP * LDQ;                        // [TP][LDT]  (rows = queries)
  bf16 *sP = sdS + TP * LDT;                         // [TP][LDT]
  int *rows = reinterpret_cast<int *>(sP + TP * LDT);
  bf16 *sOut = sdS;                                  // [TP][LDO] output staging, aliases sdS/sP after the last read
  static_assert(TP * LDO <= 2 * TP * LDT, "output staging must fit");
  const int T = ph * pw;
  const int g = blockIdx.x, head = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int t = tid; t < TP; t += NTHREADS) rows[t] = t < T ? token_row(g, t, H, W, ph, pw, window) : 0;
  __syncthreads();
  stage_qkv<TP, DH, DHP, LDQ, NTHREADS, true>(qkv, dout, sQ, sK, sV, sdO, rows, T, C, head, tid);
  __syncthreads();
  const int row0 = warp * 16;
  const int tq = lane & 3, gq = lane >> 2;
  float dq[NT_O][4];
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
      frag_a(a, sQ, LDQ, row0, ks * 16, lane);
      frag_a(ad, sdO, LDQ, row0, ks * 16, lane);
#pragma unroll
      for (int n = 0; n < NT_S; ++n) {
        uint32_t b0, b1;
        frag_b(b0, b1, sK, LDQ, n * 8, ks * 16, lane);
        mma16816(s[n], a, b0, b1);
        frag_b(b0, b1, sV, LDQ, n * 8, ks * 16, lane);
        mma16816(dp[n], ad, b0, b1);
      }
    }
    float inv0, inv1;
    softmax_rows<NT_S>(s, T, scale, tq, inv0, inv1);
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
    // P and dS: packed to bf16 once; to shared memory (for dK / dV) and kept as A fragments (for dQ)
    uint32_t ds_lo[NT_S], ds_hi[NT_S];
#pragma unroll
    for (int n = 0; n < NT_S; ++n) {
      ds_lo[n] = pack2(s[n][0] * (dp[n][0] - d0) * scale, s[n][1] * (dp[n][1] - d0) * scale);
      ds_hi[n] = pack2(s[n][2] * (dp[n][2] - d1) * scale, s[n][3] * (dp[n][3] - d1) * scale);
      const int col = n * 8 + 2 * tq;
      *reinterpret_cast<uint32_t *>(&sdS[(row0 + gq) * LDT + col]) = ds_lo[n];
      *reinterpret_cast<uint32_t *>(&sdS[(row0 + gq + 8) * LDT + col]) = ds_hi[n];
      *reinterpret_cast<uint32_t *>(&sP[(row0 + gq) * LDT + col]) = pack2(s[n][0], s[n][1]);
      *reinterpret_cast<uint32_t *>(&sP[(row0 + gq + 8) * LDT + col]) = pack2(s[n][2], s[n][3]);
    }
    if (NT_S * 8 < TP) {   // zero the token-padding columns (they are read as the k dimension below)
#pragma unroll
      for (int c = NT_S * 8 + 2 * tq; c < TP; c += 8) {
        *reinterpret_cast<uint32_t *>(&sdS[(row0 + gq) * LDT + c]) = 0u;
        *reinterpret_cast<uint32_t *>(&sdS[(row0 + gq + 8) * LDT + c]) = 0u;
        *reinterpret_cast<uint32_t *>(&sP[(row0 + gq) * LDT + c]) = 0u;
        *reinterpret_cast<uint32_t *>(&sP[(row0 + gq + 8) * LDT + c]) = 0u;
      }
    }
    // dQ = dS K  (k = key tokens, n = dh): K is [token][dh] -> transposed B loads
#pragma unroll
    for (int n = 0; n < NT_O; ++n) dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < TP / 16; ++kk) {
      uint32_t a[4];
      a[0] = ds_lo[2 * kk];
      a[1] = ds_hi[2 * kk];
      if (2 * kk + 1 < NT_S) {
        a[2] = ds_lo[2 * kk + 1];
        a[3] = ds_hi[2 * kk + 1];
      } else {
        a[2] = a[3] = 0u;
      }
#pragma unroll
      for (int n = 0; n < NT_O; ++n) {
        uint32_t b0, b1;
        frag_b_t(b0, b1, sK, LDQ, n * 8, kk * 16, lane);
        mma16816(dq[n], a, b0, b1);
      }
    }
  }
  __syncthreads();
  // this warp's 16 KEY rows of dK = dS^T Q and dV = P^T dO  (k = query tokens)
  float dk[NT_O][4], dv[NT_O][4];
#pragma unroll
  for (int n = 0; n < NT_O; ++n) {
    dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
    dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
  }
#pragma unroll
  for (int kk = 0; kk < TP / 16; ++kk) {
    uint32_t a2[4], a3[4];
    frag_a_t(a2, sdS, LDT, row0, kk * 16, lane);
    frag_a_t(a3, sP, LDT, row0, kk * 16, lane);
#pragma unroll
    for (int n = 0; n < NT_O; ++n) {
      uint32_t b0, b1;
      frag_b_t(b0, b1, sQ, LDQ, n * 8, kk * 16, lane);
      mma16816(dk[n], a2, b0, b1);
      frag_b_t(b0, b1, sdO, LDQ, n * 8, kk * 16, lane);
      mma16816(dv[n], a3, b0, b1);
    }
  }
  __syncthreads();   // every warp is done reading sdS / sP: reuse them as the [token][dq|dk|dv] staging tile
#pragma unroll
  for (int n = 0; n < NT_O; ++n) {
    const int col = n * 8 + 2 * tq;
    bf16 *r0p = sOut + (row0 + gq) * LDO + col, *r1p = sOut + (row0 + gq + 8) * LDO + col;
    *reinterpret_cast<uint32_t *>(r0p) = pack2(dq[n][0], dq[n][1]);
    *reinterpret_cast<uint32_t *>(r1p) = pack2(dq[n][2], dq[n][3]);
    *reinterpret_cast<uint32_t *>(r0p + DH) = pack2(dk[n][0], dk[n][1]);
    *reinterpret_cast<uint32_t *>(r1p + DH) = pack2(dk[n][2], dk[n][3]);
    *reinterpret_cast<uint32_t *>(r0p + 2 * DH) = pack2(dv[n][0], dv[n][1]);
    *reinterpret_cast<uint32_t *>(r1p + 2 * DH) = pack2(dv[n][2], dv[n][3]);
  }
  __syncwarp();
  constexpr int VPO = 3 * DH / 8;   // 16-byte pieces per token: the head's [dq|dk|dv] block is contiguous in dqkv
  for (int idx = lane; idx < 16 * VPO; idx += 32) {
    const int r = row0 + idx / VPO, v8 = idx % VPO;
    if (r < T)
      *reinterpret_cast<uint4 *>(dqkv + (size_t)rows[r] * 3 * C + head * 3 * DH + v8 * 8) = *reinterpret_cast<const uint4 *>(&sOut[r * LDO + v8 * 8]);
  }
}

template <int NT_S, int NT_O, int KS>
int launch_fwd(const bf16 *qkv, bf16 *out, int groups, int heads, int H, int W, int C, int ph, int pw, int window, float scale,
               cudaStream_t st) {
  constexpr int TP = ((NT_S * 8 + 15) / 16) * 16;
  LEOD_LAUNCH((attn_fwd_tc_kernel<NT_S, NT_O, KS>), dim3(groups, heads), 32 * (TP / 16), 0, st, qkv, out, H, W, C, ph, pw, window, scale);
  return 0;
}
template <int NT_S, int NT_O, int KS>
int launch_bwd(const bf16 *qkv, const bf16 *dout, bf16 *dqkv, int groups, int heads, int H, int W, int C, int ph, int pw, int window,
               float scale, cudaStream_t st) {
  constexpr int TP = ((NT_S * 8 + 15) / 16) * 16;
  constexpr int DHP = KS * 16, LDQ = DHP + 8, LDT = TP + 8;
  constexpr size_t smem = 2 * (4 * TP * LDQ + 2 * TP * LDT) + 4 * TP;
  static bool set = false;
  if (!set) {
    LEOD_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<NT_S, NT_O, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    set = true;
  }
  LEOD_LAUNCH((attn_bwd_tc_kernel<NT_S, NT_O, KS>), dim3(groups, heads), 32 * (TP / 16), smem, st, qkv, dout, dqkv, H, W, C, ph, pw, window, scale);
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
