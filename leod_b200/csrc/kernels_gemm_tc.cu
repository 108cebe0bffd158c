// Tensor-core GEMMs for sm_100a: TMA (cp.async.bulk.tensor) -> 128B-swizzled shared memory ->
// tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) -> tcgen05.ld epilogue with fused bias / GELU / residual.
//
//   NT:  C[M,N] = epi(A[M,K] * B[N,K]^T + bias)      forward layers and input gradients
//   TN:  dW[N,K] += dY[M,N]^T * X[M,K]               weight gradients (both operands MN-major)
//
// One 128-row output tile per CTA, warp-specialised: warp 0 = TMA producer, warp 1 = TMEM owner +
// MMA issuer (one elected lane), warps 2..5 = epilogue (one TMEM lane quarter each).
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int TILE_M = 128;
constexpr int TILE_K = 64;  // 64 bf16 = 128 B = one swizzle-128B row
constexpr int A_STAGE_BYTES = TILE_M * TILE_K * 2;
constexpr int NUM_THREADS = 192;

struct NtArgs {
  GemmNT g;
  int BN;         // N tile (multiple of 16, <= 256)
  int stages;     // smem ring depth
  int nkb1, nkb;  // k-blocks of source 1 / total
  int tiles_m, tiles_n;
  int acc_cols;   // TMEM columns per accumulator buffer (BN rounded up to 32)
  int tmem_cols;  // power of two >= 2 * acc_cols
  int staged;     // epilogue writes through shared memory (N multiple of 16: every chunk is full)
  int kpb;        // implicit-GEMM taps (g.taps.n > 1): k-blocks per tap
  int nbuf;       // accumulator buffers in TMEM (2..4): narrow tiles let the MMA warp run further ahead of the epilogue
  int tma_store;  // full 64-column groups of the staged epilogue leave through cp.async.bulk.tensor stores (mapC / mapAux)
  int bias_smem;  // the bias vector (N floats) is staged in shared memory once per CTA (staged epilogue, N <= 2048)
};

// epilogue warps: a multiple of 4 (one TMEM lane quarter each); the column range of a tile is split between the warps that
// share a quarter.  The GELU epilogues are instruction-bound (erf + two outputs), so they get more warps.
#ifndef NT_EPI_WARPS
#define NT_EPI_WARPS 8
#endif
#ifndef NT_EPI_WARPS_GELU
#define NT_EPI_WARPS_GELU 12
#endif
template <int EPI> struct NtEpiWarps { static constexpr int value = EPI == EPI_GELU ? NT_EPI_WARPS_GELU : NT_EPI_WARPS; };
template <int EPI> struct NtStageBytes { static constexpr int value = EPI == EPI_GELU ? 8192 : 4096; };   // per epilogue warp


// One 16-column chunk of one accumulator row: bias, fused epilogue, 2 x 16-byte stores.
__device__ __forceinline__ void nt_epilogue_chunk(const GemmNT &g, const uint32_t (&r)[16], int m, int n) {
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
  const int nvalid = min(16, g.N - n);
  if (g.bias) {
    if (nvalid == 16) {
      const float4 *b4 = reinterpret_cast<const float4 *>(g.bias + n);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 b = __ldg(b4 + q);
        v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < nvalid) v[j] += g.bias[n + j];
    }
  }
  if (g.out_f32) {
    float *cx = (float *)g.C + (size_t)m * g.ldc + n;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < nvalid) cx[j] = v[j];
    return;
  }
  if (nvalid == 16) {
    if (g.epi == EPI_GELU) {
      bf16 *ax = (bf16 *)g.aux + (size_t)m * g.ldaux + n;
      __align__(16) bf16 t[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        t[j] = __float2bfloat16_rn(gelu_grad_f(v[j]));
        v[j] = gelu_f(v[j]);
      }
      ((uint4 *)ax)[0] = ((uint4 *)t)[0];
      ((uint4 *)ax)[1] = ((uint4 *)t)[1];
    } else if (g.epi == EPI_RESID) {
      const bf16 *rx = (const bf16 *)g.R + (size_t)m * g.ldr + n;
      __align__(16) bf16 t[16];
      ((uint4 *)t)[0] = ((const uint4 *)rx)[0];
      ((uint4 *)t)[1] = ((const uint4 *)rx)[1];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] += __bfloat162float(t[j]);
    } else if (g.epi == EPI_GELU_BWD) {
      const bf16 *ax = (const bf16 *)g.aux + (size_t)m * g.ldaux + n;
      __align__(16) bf16 t[16];
      ((uint4 *)t)[0] = ((const uint4 *)ax)[0];
      ((uint4 *)t)[1] = ((const uint4 *)ax)[1];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] *= __bfloat162float(t[j]);
    }
    __align__(16) bf16 o[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = __float2bfloat16_rn(v[j]);
    bf16 *cx = (bf16 *)g.C + (size_t)m * g.ldc + n;
    ((uint4 *)cx)[0] = ((uint4 *)o)[0];
    ((uint4 *)cx)[1] = ((uint4 *)o)[1];
  } else {
    for (int j = 0; j < nvalid; ++j) {
      float x = v[j];
      if (g.epi == EPI_GELU) {
        ((bf16 *)g.aux)[(size_t)m * g.ldaux + n + j] = __float2bfloat16_rn(gelu_grad_f(x));
        x = gelu_f(x);
      } else if (g.epi == EPI_RESID) {
        x += __bfloat162float(((const bf16 *)g.R)[(size_t)m * g.ldr + n + j]);
      } else if (g.epi == EPI_GELU_BWD) {
        x *= __bfloat162float(((const bf16 *)g.aux)[(size_t)m * g.ldaux + n + j]);
      }
      ((bf16 *)g.C)[(size_t)m * g.ldc + n + j] = __float2bfloat16_rn(x);
    }
  }
}

// Packed fp32 arithmetic (sm_100 FFMA2 / FMUL2 / FADD2: two lanes per instruction).  The GELU epilogue is bound by instruction
// issue, not by HBM (profiles/r02_ncu_full_gemm_nt_gelu_stage1_details.txt: 30 instructions per element, IPC 2.1 of 4), so the
// polynomial part is evaluated on pairs of columns.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{.reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n mov.b64 rc, {%6,%7};\n fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0,%1}, rd;}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("{.reg .b64 ra, rb, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n mul.rn.f32x2 rd, ra, rb;\n mov.b64 {%0,%1}, rd;}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{.reg .b64 ra, rb, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n add.rn.f32x2 rd, ra, rb;\n mov.b64 {%0,%1}, rd;}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }
// GELU(x) = x Phi(x) and GELU'(x) = Phi(x) + x phi(x) of two columns.  erf by Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7, far below
// bf16 resolution) on the MUFU rcp / ex2 units, rearranged so that every multiply-add is packed:  E = phi(x) = 2^(-x^2 log2(e)/2 + log2(1/sqrt(2 pi))),
// h = 0.5 erfc(|x|/sqrt2) = (poly(t) t) E with the 0.5 sqrt(2 pi) folded into the coefficients,  Phi = 0.5 + sign(x) (0.5 - h).
__device__ __forceinline__ void gelu_pair(float2 x, float2 &H, float2 &G) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 u = ffma2(ax, splat2(0.3275911f * 0.70710678118654752440f), splat2(1.0f));
  float2 t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(u.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(u.y));
  constexpr float k = 0.5f / 0.39894228040143267794f;
  float2 p = ffma2(t, splat2(1.061405429f * k), splat2(-1.453152027f * k));
  p = ffma2(p, t, splat2(1.421413741f * k));
  p = ffma2(p, t, splat2(-0.284496736f * k));
  p = ffma2(p, t, splat2(0.254829592f * k));
  p = fmul2(p, t);
  const float2 ea = ffma2(fmul2(x, x), splat2(-0.72134752044448170368f), splat2(-1.32574806473615f));
  float2 E;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(E.x) : "f"(ea.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(E.y) : "f"(ea.y));
  const float2 h = fmul2(p, E);
  float2 d = ffma2(h, splat2(-1.0f), splat2(0.5f));      // 0.5 - h >= 0 up to rounding
  d.x = __uint_as_float((__float_as_uint(d.x) & 0x7fffffffu) | (__float_as_uint(x.x) & 0x80000000u));
  d.y = __uint_as_float((__float_as_uint(d.y) & 0x7fffffffu) | (__float_as_uint(x.y) & 0x80000000u));
  const float2 cdf = fadd2(d, splat2(0.5f));
  H = fmul2(x, cdf);
  G = ffma2(x, E, cdf);
}
// Same arithmetic for a full 16-column chunk, but the results stay in registers (packed bf16) so the caller can
// stage them in shared memory and write whole 128-byte lines: o = output chunk, t = pre-GELU chunk (EPI_GELU only).
template <int EPI>
__device__ __forceinline__ void nt_epilogue_compute(const GemmNT &g, const uint32_t (&r)[16], int m, int n, bool row_ok,
                                                    uint4 (&o)[2], uint4 (&t)[2], const float *s_bias, const uint4 (&pre)[2]) {
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
  if (s_bias) {   // staged once per CTA: a broadcast shared-memory read instead of a global load per row and chunk
    const float4 *b4 = reinterpret_cast<const float4 *>(s_bias + n);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b = b4[q];
      v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
    }
  } else if (g.bias) {
    const float4 *b4 = reinterpret_cast<const float4 *>(g.bias + n);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b = __ldg(b4 + q);
      v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
    }
  }
  if (EPI == EPI_GELU) {
    uint32_t *tw = reinterpret_cast<uint32_t *>(t);
#pragma unroll
    for (int j = 0; j < 8; ++j) {   // gelu and gelu' share erf and exp: aux gets the derivative, the backward is a multiply
      float2 H, G;
      gelu_pair(make_float2(v[2 * j], v[2 * j + 1]), H, G);
      const __nv_bfloat162 pr = __floats2bfloat162_rn(G.x, G.y);
      tw[j] = *reinterpret_cast<const uint32_t *>(&pr);
      v[2 * j] = H.x;
      v[2 * j + 1] = H.y;
    }
  } else if (EPI == EPI_RESID) {
    if (row_ok) {   // the residual row piece was prefetched by the caller (global latency off the per-chunk chain)
      const uint32_t w[8] = {pre[0].x, pre[0].y, pre[0].z, pre[0].w, pre[1].x, pre[1].y, pre[1].z, pre[1].w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {   // bf16 -> fp32 is a 16-bit shift
        v[2 * j] += __uint_as_float(w[j] << 16);
        v[2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
      }
    }
  } else if (EPI == EPI_GELU_BWD) {
    if (row_ok) {
      const uint32_t w[8] = {pre[0].x, pre[0].y, pre[0].z, pre[0].w, pre[1].x, pre[1].y, pre[1].z, pre[1].w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[2 * j] *= __uint_as_float(w[j] << 16);
        v[2 * j + 1] *= __uint_as_float(w[j] & 0xffff0000u);
      }
    }
  }
  uint32_t *ow = reinterpret_cast<uint32_t *>(o);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const __nv_bfloat162 pr = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    ow[j] = *reinterpret_cast<const uint32_t *>(&pr);
  }
}

// Write a staged [32 rows x 16*GC columns] bf16 block of one warp to global memory, 128 contiguous bytes per 8 lanes.
// Staging layout: row pitch 128 B, 16-byte chunk c of row r at physical chunk c ^ (r & 7).
template <int GC>
__device__ __forceinline__ void nt_flush_stage_n(const uint8_t *stage, bf16 *dst, int ld, int m_base, int rows_ok, int n_base, int lane) {
  constexpr int CPR = 2 * GC;          // 16-byte chunks per row
#pragma unroll
  for (int it = 0; it < CPR; ++it) {
    const int q = lane + 32 * it;
    const int row = q / CPR, c = q % CPR;
    if (row < rows_ok) {
      const uint4 v = *reinterpret_cast<const uint4 *>(stage + row * 128 + ((c ^ (row & 7)) << 4));
      *reinterpret_cast<uint4 *>(dst + (size_t)(m_base + row) * ld + n_base + c * 8) = v;
    }
  }
}
__device__ __forceinline__ void nt_flush_stage(const uint8_t *stage, bf16 *dst, int ld, int m_base, int M, int n_base, int gc,
                                               int lane) {
  const int rows_ok = min(32, M - m_base);
  switch (gc) {
    case 4: nt_flush_stage_n<4>(stage, dst, ld, m_base, rows_ok, n_base, lane); break;
    case 3: nt_flush_stage_n<3>(stage, dst, ld, m_base, rows_ok, n_base, lane); break;
    case 2: nt_flush_stage_n<2>(stage, dst, ld, m_base, rows_ok, n_base, lane); break;
    default: nt_flush_stage_n<1>(stage, dst, ld, m_base, rows_ok, n_base, lane); break;
  }
}

// TMA store of a staged [32 rows x 64 columns] bf16 block (SWIZZLE_128B layout = the epilogue's staging layout) and its bookkeeping
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Persistent: every CTA walks the tile list with stride gridDim.x (n-tile fastest, so CTAs running side by side share
// an A tile through L2).  The TMA ring runs ahead across tile boundaries, the accumulator is double-buffered in TMEM
// so the MMAs of tile j+1 overlap the epilogue of tile j, and eight epilogue warps (two per TMEM lane quarter, each
// taking half of the columns) drain it.
template <int EPI>   // EPI_* : staged epilogue specialised for that mode;  -1 : generic per-thread stores (ragged N)
__global__ void __launch_bounds__(64 + 32 * NtEpiWarps<EPI>::value, 1) gemm_nt_tc_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                    const __grid_constant__ CUtensorMap mapA2,
                                                                    const __grid_constant__ CUtensorMap mapB,
                                                                    const __grid_constant__ CUtensorMap mapB2,
                                                                    const __grid_constant__ CUtensorMap mapC,
                                                                    const __grid_constant__ CUtensorMap mapAux, const NtArgs a) {
  pdl_launch_dependents();   // the next kernel of the stream may start its own prologue
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int b_stage_bytes = a.BN * TILE_K * 2;
  const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;
  uint64_t *bars = (uint64_t *)(smem + (size_t)a.stages * stage_bytes);
  uint64_t *full_bar = bars, *empty_bar = bars + a.stages, *tmem_full = bars + 2 * a.stages, *tmem_empty = tmem_full + 4;
  uint32_t *tmem_slot = (uint32_t *)(tmem_empty + 4);
  // per epilogue warp: 4 KB staging blocks (32 rows x 64 columns bf16, 128-byte rows, 16-byte chunk c of row r at chunk c ^ (r & 7)):
  // 1024-byte aligned, so a block is exactly a SWIZZLE_128B TMA box {64, 32}
  uint8_t *stage_base = (uint8_t *)(((uintptr_t)(tmem_slot + 4) + 1023) & ~(uintptr_t)1023);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = a.tiles_m * a.tiles_n;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 4; ++b) {
      mbar_init(smem_u32(&tmem_full[b]), 1);
      mbar_init(smem_u32(&tmem_empty[b]), NtEpiWarps<EPI>::value);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // barriers, TMEM and tensor maps are set up: only now depend on the previous kernel's results

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / a.tiles_n) * TILE_M, n0 = (tile % a.tiles_n) * a.BN;
        for (int kb = 0; kb < a.nkb; ++kb, ++it) {
          const int s = it % a.stages;
          const uint32_t ph = (it / a.stages) & 1;
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
          const uint32_t fb = smem_u32(&full_bar[s]);
          mbar_expect_tx(fb, stage_bytes);
          const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes), sb = sa + A_STAGE_BYTES;
          if (a.kpb > 0) {
            // implicit-GEMM tap: the A tile is the same 128 rows shifted by the tap's row offset (rows outside the
            // matrix are zero-filled by TMA), the B tile is that tap's column block of the weight
            const int tap = kb / a.kpb, kc = kb - tap * a.kpb;
            tma_load_2d(sa, &mapA, fb, kc * TILE_K, m0 + a.g.taps.off[tap]);
            tma_load_2d(sb, &mapB, fb, tap * a.g.taps.cinp + kc * TILE_K, n0);
          } else if (kb < a.nkb1) {
            tma_load_2d(sa, &mapA, fb, kb * TILE_K, m0);
            tma_load_2d(sb, &mapB, fb, kb * TILE_K, n0);
          } else {
            tma_load_2d(sa, &mapA2, fb, (kb - a.nkb1) * TILE_K, m0);
            tma_load_2d(sb, &mapB2, fb, (kb - a.nkb1) * TILE_K, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(TILE_M, a.BN, 0, 0);
      int it = 0, j = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++j) {
        const int buf = j % a.nbuf;
        mbar_wait(smem_u32(&tmem_empty[buf]), ((j / a.nbuf) & 1) ^ 1);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(buf * a.acc_cols);
        for (int kb = 0; kb < a.nkb; ++kb, ++it) {
          const int s = it % a.stages;
          const uint32_t ph = (it / a.stages) & 1;
          mbar_wait(smem_u32(&full_bar[s]), ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes), sb = sa + A_STAGE_BYTES;
          const uint64_t adesc = make_desc(sa, 16, 1024), bdesc = make_desc(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < TILE_K / 16; ++k) {
            // +32 bytes per 16-element K step inside the 128B swizzle row (>>4 encoded: +2)
            tc_mma_bf16(acc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          tc_commit(smem_u32(&empty_bar[s]));
        }
        tc_commit(smem_u32(&tmem_full[buf]));
      }
    }
  } else {
    // epilogue: TMEM lane quarter = warp % 4 (hardware rule), column part = (warp - 2) / 4, one accumulator row per thread
    constexpr int CS = NtEpiWarps<EPI>::value / 4;
    const int quarter = warp & 3, part = (warp - 2) >> 2;
    const int nchunks = a.BN >> 4;
    const int c_begin = (nchunks * part) / CS, c_end = (nchunks * (part + 1)) / CS;
    const GemmNT &g = a.g;
    const float *s_bias = nullptr;
    if (EPI >= 0 && a.bias_smem) {
      float *sb = reinterpret_cast<float *>(stage_base + NtEpiWarps<EPI>::value * NtStageBytes<EPI>::value);
      for (int i = threadIdx.x - 64; i < g.N; i += 32 * NtEpiWarps<EPI>::value) sb[i] = __ldg(g.bias + i);
      asm volatile("bar.sync 1, %0;" ::"n"(32 * NtEpiWarps<EPI>::value) : "memory");   // epilogue warps only
      s_bias = sb;
    }
    int j = 0;
    bool store_pending = false;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++j) {
      const int m0 = (tile / a.tiles_n) * TILE_M, n0 = (tile % a.tiles_n) * a.BN;
      const int buf = j % a.nbuf;
      const int m = m0 + quarter * 32 + lane;
      const bool row_ok = m < g.M;
      // residual / GELU' operand of the epilogue: loaded one chunk ahead (the first one before the accumulator is even ready),
      // so that the HBM latency of these per-row reads is not part of every chunk's dependent chain
      constexpr bool PRE = (EPI == EPI_RESID || EPI == EPI_GELU_BWD);
      uint4 pre[2][2];
      const bf16 *psrc = EPI == EPI_RESID ? (const bf16 *)g.R : (const bf16 *)g.aux;
      const int pld = EPI == EPI_RESID ? g.ldr : g.ldaux;
      auto prefetch = [&](int cc, uint4(&dst)[2]) {
        if (PRE && row_ok) {
          const uint4 *q = reinterpret_cast<const uint4 *>(psrc + (size_t)m * pld + n0 + cc * 16);
          dst[0] = q[0];
          dst[1] = q[1];
        }
      };
      if (c_begin < c_end) prefetch(c_begin, pre[0]);
      mbar_wait(smem_u32(&tmem_full[buf]), (j / a.nbuf) & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * a.acc_cols);
      uint32_t rbuf[2][16];
      if (c_begin < c_end) tmem_ld16_async(trow + (uint32_t)(c_begin * 16), rbuf[0]);
      if (EPI >= 0) {
        uint8_t *stC = stage_base + (warp - 2) * NtStageBytes<EPI>::value, *stT = stC + 4096;
        const int m_base = m0 + quarter * 32;
        int gs = c_begin;   // first chunk of the current group of <= 4 chunks
#pragma unroll 1
        for (int c = c_begin; c < c_end; c += 2) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int cc = c + hh;
            if (cc >= c_end) break;
            // 14 warps are allocated as 16 (granularity of four): 128 registers per thread.  The GELU epilogue does not fit a
            // second in-flight TMEM chunk in that budget (it spilled into the hot loop), so it loads chunk by chunk; the other
            // eleven warps hide the tcgen05.ld latency
            constexpr int RB = EPI == EPI_GELU ? 0 : 1;
            tmem_ld_wait(rbuf[hh & RB]);
            if (RB && cc + 1 < c_end) tmem_ld16_async(trow + (uint32_t)((cc + 1) * 16), rbuf[(hh ^ 1) & RB]);
            if (cc + 1 < c_end) prefetch(cc + 1, pre[hh ^ 1]);
            uint4 o[2], t[2];
            nt_epilogue_compute<EPI>(g, rbuf[hh & RB], m, n0 + cc * 16, row_ok, o, t, s_bias, pre[hh]);
            if (!RB && cc + 1 < c_end) tmem_ld16_async(trow + (uint32_t)((cc + 1) * 16), rbuf[0]);
            const int slot = (cc - gs) * 2, sw = lane & 7;
            if (store_pending && cc == gs) {   // the previous group's tensor store must have read the staging block
              if (lane == 0) tma_store_wait_read();
              __syncwarp();
              store_pending = false;
            }
            *reinterpret_cast<uint4 *>(stC + lane * 128 + (((slot) ^ sw) << 4)) = o[0];
            *reinterpret_cast<uint4 *>(stC + lane * 128 + (((slot + 1) ^ sw) << 4)) = o[1];
            if (EPI == EPI_GELU) {
              *reinterpret_cast<uint4 *>(stT + lane * 128 + (((slot) ^ sw) << 4)) = t[0];
              *reinterpret_cast<uint4 *>(stT + lane * 128 + (((slot + 1) ^ sw) << 4)) = t[1];
            }
            if (cc - gs == 3 || cc == c_end - 1) {
              const int gc = cc - gs + 1;
              if (a.tma_store && gc == 4) {
                // a full 64-column group: the staged block leaves as ONE asynchronous tensor store per output (rows beyond M are
                // clipped by the tensor map) instead of 8 x (ld.shared + st.global) per lane on the warps that pace the kernel
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                  tma_store_2d(&mapC, smem_u32(stC), n0 + gs * 16, m_base);
                  if (EPI == EPI_GELU) tma_store_2d(&mapAux, smem_u32(stT), n0 + gs * 16, m_base);
                  tma_store_commit();
                }
                store_pending = true;
              } else {
                __syncwarp();
                nt_flush_stage(stC, (bf16 *)g.C, g.ldc, m_base, g.M, n0 + gs * 16, gc, lane);
                if (EPI == EPI_GELU) nt_flush_stage(stT, (bf16 *)g.aux, g.ldaux, m_base, g.M, n0 + gs * 16, gc, lane);
                __syncwarp();
              }
              gs = cc + 1;
            }
          }
        }
      } else {
#pragma unroll 1
        for (int c = c_begin; c < c_end; c += 2) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int cc = c + hh;
            if (cc >= c_end) break;
            tmem_ld_wait(rbuf[hh]);
            if (cc + 1 < c_end) tmem_ld16_async(trow + (uint32_t)((cc + 1) * 16), rbuf[hh ^ 1]);
            const int n = n0 + cc * 16;
            if (row_ok && n < g.N) nt_epilogue_chunk(g, rbuf[hh], m, n);
          }
        }
      }
      // all tcgen05.ld of this warp have completed (wait::ld above): hand the accumulator back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[buf]));
    }
    if (a.tma_store && lane == 0) tma_store_wait_all();   // outstanding tensor stores complete before the CTA exits
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------ tensor maps
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
  });
  return fn;
}

struct MapKey {
  const void *ptr;
  uint64_t inner, outer, ld;
  uint32_t box_inner, box_outer;
  bool operator<(const MapKey &o) const {
    return std::tie(ptr, inner, outer, ld, box_inner, box_outer) <
           std::tie(o.ptr, o.inner, o.outer, o.ld, o.box_inner, o.box_outer);
  }
};

// 2D bf16 tensor map: `inner` contiguous elements per row, `outer` rows, row pitch `ld` elements.
int make_map(CUtensorMap *out, const void *ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
             uint32_t box_outer) {
  static std::map<MapKey, CUtensorMap> cache;
  static std::mutex mu;
  MapKey key{ptr, inner, outer, ld, box_inner, box_outer};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  auto fn = get_encode_fn();
  LEOD_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
  LEOD_REQUIRE(((uintptr_t)ptr & 15) == 0, "TMA operand %p is not 16-byte aligned", ptr);
  LEOD_REQUIRE((ld * 2) % 16 == 0, "TMA operand row pitch %llu elements is not a multiple of 8", (unsigned long long)ld);
  LEOD_REQUIRE(box_inner * 2 <= 128 && box_outer <= 256, "TMA box %ux%u out of range", box_inner, box_outer);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LEOD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (ptr %p dims %llux%llu ld %llu box %ux%u)", (int)r, ptr,
               (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, box_inner, box_outer);
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 4096) cache.clear();
  cache[key] = *out;
  return 0;
}

// N tile: as wide as possible (A is then read once and B stays L2-resident), but small-M problems need enough tiles to
// cover the 148 SMs, so narrow the tile until there are >= ~120 of them.
int pick_bn(int M, int N, int K) {
  const int cap = 256;
  (void)K;
  const int tiles_m = ceil_div(M, TILE_M);
  if (tiles_m < 2 * 148) {
    // Small problems (the neck / head layers, the late backbone stages): a handful of waves at most, so what matters is how
    // many SMs take part and how long the longest CTA runs.  Cost model: waves x (tile width + a fixed per-tile share for the
    // A tile, pipeline fill and epilogue start-up, about 64 columns' worth).
    int best = 0;
    long best_cost = 0;
    for (int bn = cap; bn >= 16; bn -= 16) {
      if (N % bn != 0 && !(N <= bn && bn == (int)round_up(N, 16))) continue;
      const long waves = ceil_div((long)tiles_m * ceil_div(N, bn), 148);
      const long cost = waves * (bn + 64);
      if (best == 0 || cost < best_cost) { best = bn; best_cost = cost; }
    }
    if (best) return best;
    return N <= cap ? (int)round_up(N, 16) : 256;
  }
  int best = 0;
  for (int bn = cap; bn >= 32; bn -= 16) {
    if (N % bn != 0 && !(N <= bn && bn == (int)round_up(N, 16))) continue;
    if (best == 0) best = bn;
    if (tiles_m * ceil_div(N, bn) >= 120) return bn;
    best = bn;
  }
  if (best) return best;
  return N <= cap ? (int)round_up(N, 16) : (cap == 128 ? 128 : 256);
}


// ------------------------------------------------------------------ TN (weight gradient) kernel
// dW[n,k] += sum_m dY[m,n] X[m,k].  Both operands are MN-major for the MMA (the reduction index m is
// the row index in memory), staged as 64-token x 64-element TMA boxes with 128B swizzle:
//   canonical MN-major SW128 layout: 64 contiguous MN elements per token row (128 B), 8-row groups
//   1024 B apart (SBO), successive 64-element MN blocks one box (8192 B) apart (LBO).
struct TnArgs {
  float *dW; int ldw;
  float *dbias;       // != NULL: dbias[n] += sum_m dY[m,n], computed by the k-tile-0 CTAs as one more MMA against a tile of ones
  int M, N, K;
  int BKt;            // output columns per CTA (multiple of 64, <= 256)
  int rows_per_split; // multiple of 64
  int stages;
  int tmem_cols;
  ConvTaps taps;      // taps.n > 1: blockIdx.y also enumerates taps; X rows shifted by the tap offset, dW column block per tap
  int tiles_k;
};

constexpr int TN_BOX_BYTES = 64 * 64 * 2;  // 64 tokens x 64 elements

__global__ void __launch_bounds__(NUM_THREADS) gemm_tn_tc_kernel(const __grid_constant__ CUtensorMap mapY,
                                                                 const __grid_constant__ CUtensorMap mapX, const TnArgs a) {
  pdl_launch_dependents();   // the next kernel of the stream may start its own prologue
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem_al = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *ones = smem_al;                                 // one operand box of bf16 1.0 (bias gradient), then the ring
  uint8_t *smem = smem_al + TN_BOX_BYTES;
  const bool do_bias = a.dbias != nullptr && blockIdx.y == 0;
  const int nbx = a.BKt / 64;                              // X boxes per stage
  const int stage_bytes = (2 + nbx) * TN_BOX_BYTES;
  uint64_t *bars = (uint64_t *)(smem + (size_t)a.stages * stage_bytes);
  uint64_t *full_bar = bars, *empty_bar = bars + a.stages, *tmem_full = bars + 2 * a.stages;
  uint32_t *tmem_slot = (uint32_t *)(bars + 2 * a.stages + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tap = blockIdx.y / a.tiles_k;
  const int n0 = blockIdx.x * TILE_M, k0 = (blockIdx.y - tap * a.tiles_k) * a.BKt;
  const int xoff = a.taps.n > 1 ? a.taps.off[tap] : 0;
  const int m_begin = blockIdx.z * a.rows_per_split;
  const int m_end = min(a.M, m_begin + a.rows_per_split);
  const int nkb = (m_end - m_begin + 63) / 64;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(tmem_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (do_bias) {   // every element is 1.0, so the swizzled operand layout needs no care
    for (int i = threadIdx.x; i < TN_BOX_BYTES / 16; i += blockDim.x)
      reinterpret_cast<uint4 *>(ones)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // barriers, TMEM and tensor maps are set up: only now depend on the previous kernel's results

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % a.stages;
        const uint32_t ph = (kb / a.stages) & 1;
        mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
        const uint32_t fb = smem_u32(&full_bar[s]);
        mbar_expect_tx(fb, stage_bytes);
        const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
        const int m = m_begin + kb * 64;
        // rows beyond m_end belong to the next split: they are loaded (or zero-filled past M) but the
        // split boundaries are multiples of 64, so only the global tail is ever partial
        tma_load_2d(sa, &mapY, fb, n0, m);
        tma_load_2d(sa + TN_BOX_BYTES, &mapY, fb, n0 + 64, m);
        for (int j = 0; j < nbx; ++j) tma_load_2d(sa + (2 + j) * TN_BOX_BYTES, &mapX, fb, k0 + 64 * j, m + xoff);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(TILE_M, a.BKt, 1, 1), idesc_ones = make_idesc(TILE_M, 64, 1, 1);
      const uint64_t odesc = make_desc(smem_u32(ones), TN_BOX_BYTES, 1024);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % a.stages;
        const uint32_t ph = (kb / a.stages) & 1;
        mbar_wait(smem_u32(&full_bar[s]), ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes), sb = sa + 2 * TN_BOX_BYTES;
        const uint64_t adesc = make_desc(sa, TN_BOX_BYTES, 1024), bdesc = make_desc(sb, TN_BOX_BYTES, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // 16 tokens per MMA = 16 rows of 128 B = 2048 B (>>4: +128)
          tc_mma_bf16(tmem_base, adesc + 128 * k, bdesc + 128 * k, idesc, (kb | k) != 0);
        }
        if (do_bias) {   // column sums of the dY tile: the same A operand against ones, accumulated right of the dW tile
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_base + (uint32_t)a.BKt, adesc + 128 * k, odesc + 128 * k, idesc_ones, (kb | k) != 0);
        }
        tc_commit(smem_u32(&empty_bar[s]));
      }
      tc_commit(smem_u32(tmem_full));
    }
  } else {
    const int quarter = warp & 3;
    const int n = n0 + quarter * 32 + lane;
    mbar_wait(smem_u32(tmem_full), 0);
    tc_fence_after();
    if (do_bias) {
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)a.BKt, r);
      if (n < a.N) atomicAdd(a.dbias + n, __uint_as_float(r[0]));
    }
    for (int c = 0; c < a.BKt; c += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c, r);
      if (n >= a.N) continue;
      float *dst = a.dW + (size_t)n * a.ldw + (a.taps.n > 1 ? tap * a.taps.cinp : 0) + k0 + c;
      if (gridDim.z == 1 && k0 + c + 16 <= a.K && (((uintptr_t)dst) & 15) == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 v = reinterpret_cast<float4 *>(dst)[q];
          v.x += __uint_as_float(r[4 * q]); v.y += __uint_as_float(r[4 * q + 1]);
          v.z += __uint_as_float(r[4 * q + 2]); v.w += __uint_as_float(r[4 * q + 3]);
          reinterpret_cast<float4 *>(dst)[q] = v;
        }
      } else if (k0 + c + 16 <= a.K && (((uintptr_t)dst) & 15) == 0) {
        // split-M partial sums: vector reductions (4 floats per L2 operation)
#pragma unroll
        for (int q = 0; q < 4; ++q)
          asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * q), "f"(__uint_as_float(r[4 * q])),
                       "f"(__uint_as_float(r[4 * q + 1])), "f"(__uint_as_float(r[4 * q + 2])), "f"(__uint_as_float(r[4 * q + 3]))
                       : "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (k0 + c + j < a.K) atomicAdd(dst + j, __uint_as_float(r[j]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
  }
}

// ============================================================================================ fused ConvLSTM recurrence
// One launch runs ALL L timesteps of a stage's recurrence (models/layers/rnn.py:53-68 with the input half of the gates
// precomputed):   gates_t = Gx_t + h_{t-1} W_h^T ;  f,i,o = sigmoid, g = tanh ;  c_t = f c_{t-1} + i g ;  h_t = o tanh(c_t).
// The "conv" is 1x1, so token rows are independent: a CTA owns 128-token tiles and walks them through time.  h_t is
// written to global memory (it is an output anyway) and comes back as the next step's A operand through TMA; a
// per-(tile, t) arrival counter in global memory orders the two (generic-proxy stores -> fence.proxy.async -> release
// add; acquire load -> TMA).  Wide stages (4C > 256 accumulator columns) run in passes of CW channels x 4 gates; when
// tiles x passes fits on the GPU every (tile, pass) gets its own CTA and the counters also synchronise the passes.
// Warps: 0 = TMA producer, 1 = MMA issuer, 2..5 = gate math (one TMEM lane quarter each).  The gate-math warps move
// Gx / c / h / activated gates through swizzled shared-memory staging so every global access is a 128-byte row segment.
struct LstmSeqArgs {
  bf16 *gates;          // [L][M][4C]  in: x-half pre-activations (+bias); out: activated gates (f|i|o|g)
  const bf16 *c0;       // [M][C] or null
  bf16 *h_all, *c_all;  // [L][M][C]
  unsigned *flags;      // [tiles_m][L] arrival counters, zero on entry
  int M, C, L, CW, npass, split, has_h0, nkb, stages, tiles_m;
};
constexpr int LS_THREADS = 192;
constexpr int LS_WARP_STAGE = 4 * 4096 + 4096 + 4096;   // per gate-math warp: gates [4][32 rows][128 B], c, h

__device__ __forceinline__ int64_t round_up_dev(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, bool valid) {
  const int sz = valid ? 16 : 0;   // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

__global__ void __launch_bounds__(LS_THREADS, 1) lstm_seq_fwd_kernel(const __grid_constant__ CUtensorMap mapH,   // h_all [C, M, L]
                                                                      const __grid_constant__ CUtensorMap mapH0,  // h0 [C, M, 1]
                                                                      const __grid_constant__ CUtensorMap mapW,   // W_h [C, 4C]
                                                                      const LstmSeqArgs a) {
  pdl_launch_dependents();   // the next kernel of the stream may start its own prologue
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int b_stage_bytes = 4 * a.CW * TILE_K * 2;
  const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;
  uint8_t *stage_base = smem + (size_t)a.stages * stage_bytes;          // 4 x LS_WARP_STAGE, 1024-aligned
  uint64_t *bars = (uint64_t *)(stage_base + 4 * LS_WARP_STAGE);
  uint64_t *full_bar = bars, *empty_bar = bars + a.stages, *tmem_full = bars + 2 * a.stages, *tmem_empty = tmem_full + 2;
  uint32_t *tmem_slot = (uint32_t *)(tmem_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int acc_cols = 4 * a.CW;               // <= 256
  const int C = a.C, M = a.M;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tmem_full[b]), 1);
      mbar_init(smem_u32(&tmem_empty[b]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // barriers, TMEM and tensor maps are set up: only now depend on the previous kernel's results

  // work list of this CTA: split -> one (tile, pass); otherwise tiles blockIdx.x, +gridDim.x, ... with all passes
  const int tile_first = a.split ? (int)blockIdx.x / a.npass : (int)blockIdx.x;
  const int tile_step = a.split ? a.tiles_m : (int)gridDim.x;       // split: exactly one tile
  const int pass_first = a.split ? (int)blockIdx.x % a.npass : 0;
  const int pass_count = a.split ? 1 : a.npass;
  const unsigned flag_target = 4u * (unsigned)a.npass;              // 4 gate-math warps per (tile, pass)

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = 0; t < a.L; ++t) {
        if (t == 0 && !a.has_h0) continue;
        for (int tile = tile_first; tile < a.tiles_m; tile += tile_step) {
          if (t > 0) {   // h_{t-1} of this tile must be complete (all passes) and visible to the async proxy
            const unsigned *f = a.flags + (size_t)tile * a.L + (t - 1);
            unsigned v;
            const long long t0 = clock64();
            do {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
              if (v < flag_target && clock64() - t0 > 20000000000LL) __trap();
            } while (v < flag_target);
            asm volatile("fence.proxy.async;" ::: "memory");
          }
          for (int p = 0; p < pass_count; ++p) {
            const int pass = pass_first + p;
            for (int kb = 0; kb < a.nkb; ++kb, ++it) {
              const int s = it % a.stages;
              const uint32_t ph = (it / a.stages) & 1;
              mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
              const uint32_t fb = smem_u32(&full_bar[s]);
              mbar_expect_tx(fb, stage_bytes);
              const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes), sb = sa + A_STAGE_BYTES;
              if (t > 0) tma_load_3d(sa, &mapH, fb, kb * TILE_K, tile * TILE_M, t - 1);
              else tma_load_3d(sa, &mapH0, fb, kb * TILE_K, tile * TILE_M, 0);
              for (int g = 0; g < 4; ++g)
                tma_load_2d(sb + g * a.CW * TILE_K * 2, &mapW, fb, kb * TILE_K, g * C + pass * a.CW);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(TILE_M, acc_cols, 0, 0);
      int it = 0, j = 0;
      for (int t = 0; t < a.L; ++t) {
        if (t == 0 && !a.has_h0) continue;
        for (int tile = tile_first; tile < a.tiles_m; tile += tile_step) {
          for (int p = 0; p < pass_count; ++p, ++j) {
            const int buf = j & 1;
            mbar_wait(smem_u32(&tmem_empty[buf]), ((j >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + (uint32_t)(buf * 256);
            for (int kb = 0; kb < a.nkb; ++kb, ++it) {
              const int s = it % a.stages;
              const uint32_t ph = (it / a.stages) & 1;
              mbar_wait(smem_u32(&full_bar[s]), ph);
              tc_fence_after();
              const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes), sb = sa + A_STAGE_BYTES;
              const uint64_t adesc = make_desc(sa, 16, 1024), bdesc = make_desc(sb, 16, 1024);
#pragma unroll
              for (int k = 0; k < TILE_K / 16; ++k) tc_mma_bf16(acc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
              tc_commit(smem_u32(&empty_bar[s]));
            }
            tc_commit(smem_u32(&tmem_full[buf]));
          }
        }
      }
    }
  } else {
    const int quarter = warp & 3;
    uint8_t *stG = stage_base + (warp - 2) * LS_WARP_STAGE, *stC = stG + 4 * 4096, *stH = stC + 4096;
    const int ppr = a.CW >> 3;                    // 16-byte pieces per row segment (2..8); a lane moves ppr pieces per tensor
    const int nchunk = a.CW >> 4;                 // 16-channel chunks per pass
    const uint32_t stG_u32 = smem_u32(stG), stC_u32 = smem_u32(stC);
    int prow[8], pcol[8];
    uint32_t poff[8];                             // piece k of this lane: tile row, channel offset, swizzled staging offset
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int q = lane + 32 * k;
      prow[k] = q / ppr;
      const int c16 = q - prow[k] * ppr;
      pcol[k] = c16 * 8;
      poff[k] = prow[k] * 128 + ((c16 ^ (prow[k] & 7)) << 4);
    }
    int j = 0;
    for (int t = 0; t < a.L; ++t) {
      const bool mma = !(t == 0 && !a.has_h0);
      for (int tile = tile_first; tile < a.tiles_m; tile += tile_step) {
        const int m_base = tile * TILE_M + quarter * 32;
        for (int p = 0; p < pass_count; ++p) {
          const int pass = pass_first + p;
          const int ch0 = pass * a.CW;
          // ---- stage Gx (4 gates) and c_{t-1}: coalesced 16-byte async copies into the swizzled tiles
          {
            const bf16 *gbase = a.gates + (size_t)t * M * 4 * C + ch0;
            const bf16 *cbase = t > 0 ? a.c_all + (size_t)(t - 1) * M * C + ch0 : (a.c0 ? a.c0 + ch0 : nullptr);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              if (k < ppr) {
                const int m = m_base + prow[k];
                const bool ok = m < M;
                const size_t mo = ok ? (size_t)m : 0;
                const bf16 *gsrc = gbase + mo * 4 * C + pcol[k];
#pragma unroll
                for (int g = 0; g < 4; ++g) cp_async16(stG_u32 + g * 4096 + poff[k], gsrc + (size_t)g * C, ok);
                cp_async16(stC_u32 + poff[k], cbase ? (const void *)(cbase + mo * C + pcol[k]) : (const void *)a.gates, ok && cbase != nullptr);
              }
            }
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
          int buf = 0;
          if (mma) {
            buf = j & 1;
            mbar_wait(smem_u32(&tmem_full[buf]), (j >> 1) & 1);
            tc_fence_after();
          }
          asm volatile("cp.async.wait_group 0;" ::: "memory");
          __syncwarp();
          const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * 256);
          const uint32_t rowoff = lane * 128;
          const int sw = lane & 7;
          for (int cc = 0; cc < nchunk; ++cc) {
            float pre[4][16];
            if (mma) {
              uint32_t r[4][16];
#pragma unroll
              for (int g = 0; g < 4; ++g) tmem_ld16_async(trow + (uint32_t)(g * a.CW + cc * 16), r[g]);
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                tmem_ld_wait(r[g]);
#pragma unroll
                for (int e = 0; e < 16; ++e) pre[g][e] = __uint_as_float(r[g][e]);
              }
            } else {
#pragma unroll
              for (int g = 0; g < 4; ++g)
#pragma unroll
                for (int e = 0; e < 16; ++e) pre[g][e] = 0.f;
            }
            const uint32_t o0 = rowoff + (((2 * cc) ^ sw) << 4), o1 = rowoff + (((2 * cc + 1) ^ sw) << 4);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 x0 = *reinterpret_cast<const uint4 *>(stG + g * 4096 + o0), x1 = *reinterpret_cast<const uint4 *>(stG + g * 4096 + o1);
              const uint32_t w[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                pre[g][2 * e] += __uint_as_float(w[e] << 16);
                pre[g][2 * e + 1] += __uint_as_float(w[e] & 0xffff0000u);
              }
            }
            float cp[16];
            {
              const uint4 x0 = *reinterpret_cast<const uint4 *>(stC + o0), x1 = *reinterpret_cast<const uint4 *>(stC + o1);
              const uint32_t w[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                cp[2 * e] = __uint_as_float(w[e] << 16);
                cp[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
              }
            }
            uint32_t og[4][8], oc[8], oh[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float f[2], ig[2], o[2], gg[2], cn[2], hn[2];
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                f[u] = sigmoid_fast(pre[0][2 * e + u]);
                ig[u] = sigmoid_fast(pre[1][2 * e + u]);
                o[u] = sigmoid_fast(pre[2][2 * e + u]);
                gg[u] = tanh_fast(pre[3][2 * e + u]);
                cn[u] = fmaf(f[u], cp[2 * e + u], ig[u] * gg[u]);
                hn[u] = o[u] * tanh_fast(cn[u]);
              }
              __nv_bfloat162 v;
              v = __floats2bfloat162_rn(f[0], f[1]); og[0][e] = *reinterpret_cast<uint32_t *>(&v);
              v = __floats2bfloat162_rn(ig[0], ig[1]); og[1][e] = *reinterpret_cast<uint32_t *>(&v);
              v = __floats2bfloat162_rn(o[0], o[1]); og[2][e] = *reinterpret_cast<uint32_t *>(&v);
              v = __floats2bfloat162_rn(gg[0], gg[1]); og[3][e] = *reinterpret_cast<uint32_t *>(&v);
              v = __floats2bfloat162_rn(cn[0], cn[1]); oc[e] = *reinterpret_cast<uint32_t *>(&v);
              v = __floats2bfloat162_rn(hn[0], hn[1]); oh[e] = *reinterpret_cast<uint32_t *>(&v);
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              *reinterpret_cast<uint4 *>(stG + g * 4096 + o0) = make_uint4(og[g][0], og[g][1], og[g][2], og[g][3]);
              *reinterpret_cast<uint4 *>(stG + g * 4096 + o1) = make_uint4(og[g][4], og[g][5], og[g][6], og[g][7]);
            }
            *reinterpret_cast<uint4 *>(stC + o0) = make_uint4(oc[0], oc[1], oc[2], oc[3]);
            *reinterpret_cast<uint4 *>(stC + o1) = make_uint4(oc[4], oc[5], oc[6], oc[7]);
            *reinterpret_cast<uint4 *>(stH + o0) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
            *reinterpret_cast<uint4 *>(stH + o1) = make_uint4(oh[4], oh[5], oh[6], oh[7]);
          }
          if (mma) {   // accumulator drained: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[buf]));
            ++j;
          } else {
            __syncwarp();
          }
          // ---- write activated gates, c_t, h_t as 16-byte pieces of contiguous row segments
          {
            bf16 *gbase = a.gates + (size_t)t * M * 4 * C + ch0;
            bf16 *cdst = a.c_all + (size_t)t * M * C + ch0, *hdst = a.h_all + (size_t)t * M * C + ch0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              if (k < ppr) {
                const int m = m_base + prow[k];
                if (m < M) {
                  bf16 *gdst = gbase + (size_t)m * 4 * C + pcol[k];
#pragma unroll
                  for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4 *>(gdst + (size_t)g * C) = *reinterpret_cast<const uint4 *>(stG + g * 4096 + poff[k]);
                  *reinterpret_cast<uint4 *>(cdst + (size_t)m * C + pcol[k]) = *reinterpret_cast<const uint4 *>(stC + poff[k]);
                  *reinterpret_cast<uint4 *>(hdst + (size_t)m * C + pcol[k]) = *reinterpret_cast<const uint4 *>(stH + poff[k]);
                }
              }
            }
          }
          // publish: h_t (this pass's channels) is in global memory; make it visible to TMA reads and count the warp in
          // (generic-proxy stores -> proxy fence -> warp barrier -> one cumulative release per warp)
          asm volatile("fence.proxy.async;" ::: "memory");
          __syncwarp();
          if (lane == 0 && t + 1 < a.L) {
            unsigned *f = a.flags + (size_t)tile * a.L + t;
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(f) : "memory");
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- backward of the recurrence, same skeleton run backwards in time:
//   dh_t = dh_ext_t + dgates_{t+1} W_h ;  dc = dc_{t+1} + dh_t o (1 - tanh(c_t)^2) ;  dgates_t = (dc c_{t-1} f(1-f), dc g i(1-i),
//   dh_t tanh(c_t) o(1-o), dc i (1-g^2)) ;  dc_t-1 = dc f.        (models/layers/rnn.py:57-68 differentiated)
// dgates_t goes to global memory (the weight-gradient and input-gradient GEMMs need it anyway) and is the A operand
// (K = 4C) of the next step through TMA; one extra GEMM-only iteration produces the gradient w.r.t. h0.
struct LstmBwdArgs {
  const bf16 *gates, *c_all, *c0, *dh, *dc_last;
  bf16 *dgates, *dc_ws, *dc0, *dh0;
  unsigned *flags;
  int M, C, L, CWb, npass, split, nkb, stages, tiles_m;
};
constexpr int LB_WARP_STAGE = 8 * 4096;   // per gate-math warp: gates [4], c_prev, c, dh, dc tiles of [32 rows][128 B]

__global__ void __launch_bounds__(LS_THREADS, 1) lstm_seq_bwd_kernel(const __grid_constant__ CUtensorMap mapG,   // dgates [4C, M, L]
                                                                      const __grid_constant__ CUtensorMap mapW,   // W_h^T [4C, C]
                                                                      const LstmBwdArgs a) {
  pdl_launch_dependents();   // the next kernel of the stream may start its own prologue
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int b_stage_bytes = a.CWb * TILE_K * 2;
  const int stage_bytes = A_STAGE_BYTES + (int)round_up_dev(b_stage_bytes, 1024);
  uint8_t *stage_base = smem + (size_t)a.stages * stage_bytes;
  uint64_t *bars = (uint64_t *)(stage_base + 4 * LB_WARP_STAGE);
  uint64_t *full_bar = bars, *empty_bar = bars + a.stages, *tmem_full = bars + 2 * a.stages, *tmem_empty = tmem_full + 2;
  uint32_t *tmem_slot = (uint32_t *)(tmem_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = a.C, M = a.M, L = a.L;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tmem_full[b]), 1);
      mbar_init(smem_u32(&tmem_empty[b]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // barriers, TMEM and tensor maps are set up: only now depend on the previous kernel's results

  const int tile_first = a.split ? (int)blockIdx.x / a.npass : (int)blockIdx.x;
  const int tile_step = a.split ? a.tiles_m : (int)gridDim.x;
  const int pass_first = a.split ? (int)blockIdx.x % a.npass : 0;
  const int pass_count = a.split ? 1 : a.npass;
  const unsigned flag_target = 4u * (unsigned)a.npass;
  const int t_last = a.dh0 ? -1 : 0;    // t = -1: GEMM-only iteration for the gradient w.r.t. h0

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = L - 2; t >= t_last; --t) {
        for (int tile = tile_first; tile < a.tiles_m; tile += tile_step) {
          {   // dgates_{t+1} of this tile: all channel passes written and visible to the async proxy
            const unsigned *f = a.flags + (size_t)tile * L + (t + 1);
            unsigned v;
            const long long t0 = clock64();
            do {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
              if (v < flag_target && clock64() - t0 > 20000000000LL) __trap();
            } while (v < flag_target);
            asm volatile("fence.proxy.async;" ::: "memory");
          }
          for (int p = 0; p < pass_count; ++p) {
            const int pass = pass_first + p;
            for (int kb = 0; kb < a.nkb; ++kb, ++it) {
              const int s = it % a.stages;
              const uint32_t ph = (it / a.stages) & 1;
              mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
              const uint32_t fb = smem_u32(&full_bar[s]);
              mbar_expect_tx(fb, A_STAGE_BYTES + b_stage_bytes);
              const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes), sb = sa + A_STAGE_BYTES;
              tma_load_3d(sa, &mapG, fb, kb * TILE_K, tile * TILE_M, t + 1);
              tma_load_2d(sb, &mapW, fb, kb * TILE_K, pass * a.CWb);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(TILE_M, a.CWb, 0, 0);
      int it = 0, j = 0;
      for (int t = L - 2; t >= t_last; --t) {
        for (int tile = tile_first; tile < a.tiles_m; tile += tile_step) {
          for (int p = 0; p < pass_count; ++p, ++j) {
            const int buf = j & 1;
            mbar_wait(smem_u32(&tmem_empty[buf]), ((j >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + (uint32_t)(buf * 256);
            for (int kb = 0; kb < a.nkb; ++kb, ++it) {
              const int s = it % a.stages;
              const uint32_t ph = (it / a.stages) & 1;
              mbar_wait(smem_u32(&full_bar[s]), ph);
              tc_fence_after();
              const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes), sb = sa + A_STAGE_BYTES;
              const uint64_t adesc = make_desc(sa, 16, 1024), bdesc = make_desc(sb, 16, 1024);
#pragma unroll
              for (int k = 0; k < TILE_K / 16; ++k) tc_mma_bf16(acc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
              tc_commit(smem_u32(&empty_bar[s]));
            }
            tc_commit(smem_u32(&tmem_full[buf]));
          }
        }
      }
    }
  } else {
    const int quarter = warp & 3;
    uint8_t *st = stage_base + (warp - 2) * LB_WARP_STAGE;
    uint8_t *stG = st, *stCP = st + 4 * 4096, *stCC = stCP + 4096, *stDH = stCC + 4096, *stDC = stDH + 4096;
    const uint32_t st_u32 = smem_u32(st);
    int j = 0;
    for (int t = L - 1; t >= t_last; --t) {
      const bool mma = t < L - 1;
      for (int tile = tile_first; tile < a.tiles_m; tile += tile_step) {
        const int m_base = tile * TILE_M + quarter * 32;
        for (int p = 0; p < pass_count; ++p) {
          const int pass = pass_first + p;
          int buf = 0;
          bool waited = false;
          for (int sg0 = 0; sg0 < a.CWb; sg0 += 64) {
            const int sgw = min(64, a.CWb - sg0);
            const int ch0 = pass * a.CWb + sg0;
            const int ppr = sgw >> 3, nchunk = sgw >> 4;
            if (t >= 0) {
              // ---- stage the eight operand tiles of this channel group with coalesced async copies
              const bf16 *gbase = a.gates + (size_t)t * M * 4 * C + ch0;
              const bf16 *ccb = a.c_all + (size_t)t * M * C + ch0;
              const bf16 *cpb = t > 0 ? a.c_all + (size_t)(t - 1) * M * C + ch0 : (a.c0 ? a.c0 + ch0 : nullptr);
              const bf16 *dhb = a.dh ? a.dh + (size_t)t * M * C + ch0 : nullptr;
              const bf16 *dcb = t == L - 1 ? (a.dc_last ? a.dc_last + ch0 : nullptr) : a.dc_ws + ch0;
              for (int q = lane; q < 32 * ppr; q += 32) {
                const int row = q / ppr, c16 = q - row * ppr;
                const int m = m_base + row;
                const bool ok = m < M;
                const size_t mo = ok ? (size_t)m : 0;
                const uint32_t off = row * 128 + ((c16 ^ (row & 7)) << 4);
                const bf16 *gsrc = gbase + mo * 4 * C + c16 * 8;
#pragma unroll
                for (int g = 0; g < 4; ++g) cp_async16(st_u32 + g * 4096 + off, gsrc + (size_t)g * C, ok);
                cp_async16(st_u32 + 4 * 4096 + off, cpb ? (const void *)(cpb + mo * C + c16 * 8) : (const void *)a.gates, ok && cpb);
                cp_async16(st_u32 + 5 * 4096 + off, ccb + mo * C + c16 * 8, ok);
                cp_async16(st_u32 + 6 * 4096 + off, dhb ? (const void *)(dhb + mo * C + c16 * 8) : (const void *)a.gates, ok && dhb);
                cp_async16(st_u32 + 7 * 4096 + off, dcb ? (const void *)(dcb + mo * C + c16 * 8) : (const void *)a.gates, ok && dcb);
              }
              asm volatile("cp.async.commit_group;" ::: "memory");
            }
            if (mma && !waited) {
              buf = j & 1;
              mbar_wait(smem_u32(&tmem_full[buf]), (j >> 1) & 1);
              tc_fence_after();
              waited = true;
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * 256 + sg0);
            const uint32_t rowoff = lane * 128;
            const int sw = lane & 7;
            for (int cc = 0; cc < nchunk; ++cc) {
              float dhn[16];
              if (mma) {
                uint32_t r[16];
                tmem_ld16(trow + (uint32_t)(cc * 16), r);
#pragma unroll
                for (int e = 0; e < 16; ++e) dhn[e] = __uint_as_float(r[e]);
              } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) dhn[e] = 0.f;
              }
              const uint32_t o0 = rowoff + (((2 * cc) ^ sw) << 4), o1 = rowoff + (((2 * cc + 1) ^ sw) << 4);
              if (t < 0) {   // gradient w.r.t. h0: just the GEMM result
                uint32_t w[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  __nv_bfloat162 v = __floats2bfloat162_rn(dhn[2 * e], dhn[2 * e + 1]);
                  w[e] = *reinterpret_cast<uint32_t *>(&v);
                }
                *reinterpret_cast<uint4 *>(stDH + o0) = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4 *>(stDH + o1) = make_uint4(w[4], w[5], w[6], w[7]);
                continue;
              }
              float in[8][16];   // f, i, o, g, c_prev, c, dh_ext, dc_in
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const uint4 x0 = *reinterpret_cast<const uint4 *>(st + k * 4096 + o0), x1 = *reinterpret_cast<const uint4 *>(st + k * 4096 + o1);
                const uint32_t w[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  in[k][2 * e] = __uint_as_float(w[e] << 16);
                  in[k][2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
                }
              }
              uint32_t og[4][8], odc[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float r4[4][2], rdc[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  const int x = 2 * e + u;
                  const float f = in[0][x], ig = in[1][x], o = in[2][x], g = in[3][x], cp = in[4][x];
                  const float tc = tanh_fast(in[5][x]);
                  const float dhv = in[6][x] + dhn[x];
                  const float dcv = fmaf(dhv * o, 1.f - tc * tc, in[7][x]);
                  r4[0][u] = dcv * cp * f * (1.f - f);
                  r4[1][u] = dcv * g * ig * (1.f - ig);
                  r4[2][u] = dhv * tc * o * (1.f - o);
                  r4[3][u] = dcv * ig * (1.f - g * g);
                  rdc[u] = dcv * f;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  __nv_bfloat162 v = __floats2bfloat162_rn(r4[k][0], r4[k][1]);
                  og[k][e] = *reinterpret_cast<uint32_t *>(&v);
                }
                __nv_bfloat162 v = __floats2bfloat162_rn(rdc[0], rdc[1]);
                odc[e] = *reinterpret_cast<uint32_t *>(&v);
              }
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                *reinterpret_cast<uint4 *>(stG + k * 4096 + o0) = make_uint4(og[k][0], og[k][1], og[k][2], og[k][3]);
                *reinterpret_cast<uint4 *>(stG + k * 4096 + o1) = make_uint4(og[k][4], og[k][5], og[k][6], og[k][7]);
              }
              *reinterpret_cast<uint4 *>(stDC + o0) = make_uint4(odc[0], odc[1], odc[2], odc[3]);
              *reinterpret_cast<uint4 *>(stDC + o1) = make_uint4(odc[4], odc[5], odc[6], odc[7]);
            }
            __syncwarp();
            // ---- write dgates_t and dc (or dh0) as contiguous row segments
            if (t >= 0) {
              bf16 *gdb = a.dgates + (size_t)t * M * 4 * C + ch0;
              bf16 *dcd = ((t == 0 && a.dc0) ? a.dc0 : a.dc_ws) + ch0;
              for (int q = lane; q < 32 * ppr; q += 32) {
                const int row = q / ppr, c16 = q - row * ppr;
                const int m = m_base + row;
                if (m < M) {
                  const uint32_t off = row * 128 + ((c16 ^ (row & 7)) << 4);
                  bf16 *gdst = gdb + (size_t)m * 4 * C + c16 * 8;
#pragma unroll
                  for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4 *>(gdst + (size_t)g * C) = *reinterpret_cast<const uint4 *>(stG + g * 4096 + off);
                  *reinterpret_cast<uint4 *>(dcd + (size_t)m * C + c16 * 8) = *reinterpret_cast<const uint4 *>(stDC + off);
                }
              }
            } else {
              bf16 *hd = a.dh0 + ch0;
              for (int q = lane; q < 32 * ppr; q += 32) {
                const int row = q / ppr, c16 = q - row * ppr;
                const int m = m_base + row;
                if (m < M) *reinterpret_cast<uint4 *>(hd + (size_t)m * C + c16 * 8) = *reinterpret_cast<const uint4 *>(stDH + row * 128 + ((c16 ^ (row & 7)) << 4));
              }
            }
            __syncwarp();
          }
          if (mma) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[buf]));
            ++j;
          }
          if (t >= 0) {   // publish dgates_t of this (tile, pass)
            asm volatile("fence.proxy.async;" ::: "memory");
            __syncwarp();
            if (lane == 0 && t > t_last) {
              unsigned *f = a.flags + (size_t)tile * L + t;
              asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(f) : "memory");
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// 3D bf16 tensor map [d0 (contiguous), d1, d2] with element pitches ld1, ld2; box [b0, b1, 1], 128B swizzle
int make_map3(CUtensorMap *out, const void *ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld1, uint64_t ld2, uint32_t b0,
              uint32_t b1) {
  auto fn = get_encode_fn();
  LEOD_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
  LEOD_REQUIRE(((uintptr_t)ptr & 15) == 0 && (ld1 * 2) % 16 == 0 && (ld2 * 2) % 16 == 0, "TMA operand %p / pitches not 16-byte aligned", ptr);
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {ld1 * 2, ld2 * 2};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LEOD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed with %d", (int)r);
  return 0;
}
}  // namespace

// 2D bf16 tensor map with 128-byte swizzle for the other tcgen05 translation units (kernels_stem.cu)
int tc_make_map_2d(CUtensorMap *out, const void *ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer) {
  return make_map(out, ptr, inner, outer, ld, box_inner, box_outer);
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

int gemm_nt_tc(const GemmNT &g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return 0;
  LEOD_REQUIRE(g.K > 0, "gemm_nt_tc: K = %d", g.K);
  LEOD_REQUIRE(g.out_f32 || (g.ldc % 8 == 0 && (!g.R || g.ldr % 8 == 0) && (!g.aux || g.ldaux % 8 == 0)),
               "gemm_nt_tc: output/residual pitches must be multiples of 8 elements");
  LEOD_REQUIRE(g.out_f32 || (((uintptr_t)g.C) & 15) == 0, "gemm_nt_tc: C not 16-byte aligned");
  LEOD_REQUIRE(!g.out_f32 || g.epi == EPI_NONE, "gemm_nt_tc: fp32 output has no fused epilogue");
  const bool tapped = g.taps.n > 1;
  LEOD_REQUIRE(!tapped || (!g.A2 && g.taps.n <= 9 && g.taps.cinp % TILE_K == 0 && g.taps.cin <= g.taps.cinp && g.K == g.taps.n * g.taps.cinp),
               "gemm_nt_tc: bad tap description (n %d cin %d cinp %d K %d)", g.taps.n, g.taps.cin, g.taps.cinp, g.K);
  NtArgs a;
  a.g = g;
  a.BN = pick_bn(g.M, g.N, g.K);
  const int K1 = g.A2 ? g.K1 : g.K, K2 = g.K - K1;
  LEOD_REQUIRE(K1 > 0 && K2 >= 0 && (K2 == 0 || K1 % 8 == 0), "gemm_nt_tc: bad K split %d/%d", K1, g.K);
  a.nkb1 = ceil_div(K1, TILE_K);
  a.nkb = a.nkb1 + (K2 > 0 ? ceil_div(K2, TILE_K) : 0);
  a.kpb = tapped ? g.taps.cinp / TILE_K : 0;
  a.tiles_m = ceil_div(g.M, TILE_M);
  a.tiles_n = ceil_div(g.N, a.BN);
  a.acc_cols = (int)round_up(a.BN, 32);
  a.nbuf = std::max(2, std::min(4, 512 / a.acc_cols));
  a.tmem_cols = 32;
  while (a.tmem_cols < a.nbuf * a.acc_cols) a.tmem_cols *= 2;
  const int stage_bytes = A_STAGE_BYTES + a.BN * TILE_K * 2;
  const int total_kb = a.nkb * ceil_div(a.tiles_m * a.tiles_n, num_sms());
  a.staged = (g.N % 16 == 0 && (((uintptr_t)g.bias) & 15) == 0 && !g.out_f32) ? 1 : 0;
  const bool gelu_like = a.staged && g.epi == EPI_GELU;
  const int epi_warps = gelu_like ? NtEpiWarps<EPI_GELU>::value : NtEpiWarps<EPI_NONE>::value;
  a.bias_smem = (a.staged && g.bias && g.N <= 2048) ? 1 : 0;
  const int epi_bytes = epi_warps * ((a.staged && g.epi == EPI_GELU) ? 8192 : 4096) + 1024 /*staging alignment*/ + (a.bias_smem ? (int)round_up(g.N * 4, 16) : 0);
  a.stages = std::min(std::min(4, (int)((224 * 1024 - epi_bytes - 2048) / stage_bytes)), std::max(total_kb, 1));
  CUtensorMap mA, mA2, mB, mB2;
  LEOD_TRY(make_map(&mA, g.A, tapped ? g.taps.cin : K1, g.M, g.lda, TILE_K, TILE_M));
  LEOD_TRY(make_map(&mB, g.B, tapped ? g.K : K1, g.N, g.ldb, TILE_K, a.BN));
  if (K2 > 0) {
    LEOD_TRY(make_map(&mA2, g.A2, K2, g.M, g.lda2, TILE_K, TILE_M));
    LEOD_TRY(make_map(&mB2, (const bf16 *)g.B + K1, K2, g.N, g.ldb, TILE_K, a.BN));
  } else {
    mA2 = mA;
    mB2 = mB;
  }
  // tensor stores of the staged epilogue: C (and the GELU' side output) as [M, N] bf16 matrices, box = one staging block
  CUtensorMap mC = mA, mAux = mA;
  a.tma_store = 0;
  if (a.staged && g.N >= 64 && (((uintptr_t)g.C) & 15) == 0 && g.ldc % 8 == 0 &&
      (g.epi != EPI_GELU || (g.aux && (((uintptr_t)g.aux) & 15) == 0 && g.ldaux % 8 == 0))) {
    LEOD_TRY(make_map(&mC, g.C, g.N, g.M, g.ldc, 64, 32));
    if (g.epi == EPI_GELU) LEOD_TRY(make_map(&mAux, g.aux, g.N, g.M, g.ldaux, 64, 32));
    a.tma_store = 1;
  }
  const size_t smem = (size_t)a.stages * stage_bytes + 1024 /*align*/ + (2 * a.stages + 8) * 8 + 16 + epi_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    LEOD_CUDA(cudaFuncSetAttribute(gemm_nt_tc_kernel<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(226 * 1024)));
    LEOD_CUDA(cudaFuncSetAttribute(gemm_nt_tc_kernel<EPI_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(226 * 1024)));
    LEOD_CUDA(cudaFuncSetAttribute(gemm_nt_tc_kernel<EPI_GELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(226 * 1024)));
    LEOD_CUDA(cudaFuncSetAttribute(gemm_nt_tc_kernel<EPI_RESID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(226 * 1024)));
    LEOD_CUDA(cudaFuncSetAttribute(gemm_nt_tc_kernel<EPI_GELU_BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(226 * 1024)));
    attr_set = true;
  }
  // balanced persistent grid: the smallest CTA count that still finishes in the minimal number of waves
  const int tiles = a.tiles_m * a.tiles_n;
  const int waves = ceil_div(tiles, num_sms());
  const int grid = ceil_div(tiles, waves);
  const int threads = 64 + 32 * epi_warps;
  if (!a.staged) {
    LEOD_LAUNCH((gemm_nt_tc_kernel<-1>), grid, threads, smem, st, mA, mA2, mB, mB2, mC, mAux, a);
  } else if (g.epi == EPI_GELU) {
    LEOD_LAUNCH((gemm_nt_tc_kernel<EPI_GELU>), grid, threads, smem, st, mA, mA2, mB, mB2, mC, mAux, a);
  } else if (g.epi == EPI_RESID) {
    LEOD_LAUNCH((gemm_nt_tc_kernel<EPI_RESID>), grid, threads, smem, st, mA, mA2, mB, mB2, mC, mAux, a);
  } else if (g.epi == EPI_GELU_BWD) {
    LEOD_LAUNCH((gemm_nt_tc_kernel<EPI_GELU_BWD>), grid, threads, smem, st, mA, mA2, mB, mB2, mC, mAux, a);
  } else {
    LEOD_LAUNCH((gemm_nt_tc_kernel<EPI_NONE>), grid, threads, smem, st, mA, mA2, mB, mB2, mC, mAux, a);
  }
  LEOD_LAUNCH_CHECK();
  return 0;
}


// Whole-window ConvLSTM recurrence of one stage (see lstm_seq_fwd_kernel).  gates: [L][M][4C] holding the x-half
// pre-activations (+bias) on entry, the activated gates on exit.  Wh: prepared [4C][ldw] weight, columns 0..C-1 = hidden half.
int lstm_seq_fwd_tc(void *gates, const void *Wh, int ldw, const void *h0, const void *c0, void *h_all, void *c_all, unsigned *flags, int M,
                    int C, int L, cudaStream_t st) {
  LEOD_REQUIRE(C % 16 == 0 && M > 0 && L > 0, "lstm_seq_fwd_tc: bad shape M=%d C=%d L=%d", M, C, L);
  LstmSeqArgs a;
  a.gates = (bf16 *)gates; a.c0 = (const bf16 *)c0; a.h_all = (bf16 *)h_all; a.c_all = (bf16 *)c_all; a.flags = flags;
  a.M = M; a.C = C; a.L = L;
  // channel pass: <= 64 channels (4 gates x 64 = 256 accumulator columns); stages with few token tiles are split into
  // more, narrower passes so that ~120 CTAs share a timestep
  a.tiles_m = ceil_div(M, TILE_M);
  const int want_pass = a.tiles_m >= 40 ? 1 : ceil_div(120, a.tiles_m);
  a.CW = 16;
  for (int cw = 64; cw >= 16; cw -= 16)
    if (C % cw == 0 && C / cw >= want_pass) { a.CW = cw; break; }
  if (a.tiles_m * (C / a.CW) > num_sms())   // the narrow split only pays when every (tile, pass) gets its own CTA
    for (int cw = 64; cw >= 16; cw -= 16)
      if (C % cw == 0) { a.CW = cw; break; }
  a.npass = C / a.CW;
  a.split = (a.npass > 1 && a.tiles_m * a.npass <= num_sms()) ? 1 : 0;
  a.has_h0 = h0 != nullptr;
  a.nkb = ceil_div(C, TILE_K);
  const int stage_bytes = A_STAGE_BYTES + 4 * a.CW * TILE_K * 2;
  a.stages = std::max(2, std::min(4, (int)((224 * 1024 - 4 * LS_WARP_STAGE - 2048) / stage_bytes)));
  CUtensorMap mH, mH0, mW;
  LEOD_TRY(make_map3(&mH, h_all, C, M, L, C, (uint64_t)M * C, TILE_K, TILE_M));
  LEOD_TRY(make_map3(&mH0, h0 ? h0 : h_all, C, M, 1, C, (uint64_t)M * C, TILE_K, TILE_M));
  LEOD_TRY(make_map(&mW, Wh, C, 4 * C, ldw, TILE_K, a.CW));
  const size_t smem = (size_t)a.stages * stage_bytes + 4 * LS_WARP_STAGE + 1024 + (2 * a.stages + 4) * 8 + 64;
  LEOD_REQUIRE(smem <= 227 * 1024, "lstm_seq_fwd_tc: shared memory %zu", smem);
  static bool attr_set = false;
  if (!attr_set) {
    LEOD_CUDA(cudaFuncSetAttribute(lstm_seq_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
    attr_set = true;
  }
  LEOD_TRY(device_zero_u32(flags, (int64_t)a.tiles_m * L, st));
  const int grid = a.split ? a.tiles_m * a.npass : std::min(a.tiles_m, num_sms());
  LEOD_LAUNCH((lstm_seq_fwd_kernel), grid, LS_THREADS, smem, st, mH, mH0, mW, a);
  LEOD_LAUNCH_CHECK();
  return 0;
}

// Backward of the whole-window recurrence (lstm_seq_bwd_kernel).  WhT: prepared [C][ldw] = rows C..2C-1 of W_l^T (K = 4C).
int lstm_seq_bwd_tc(const void *gates, const void *c_all, const void *c0, const void *dh, const void *dc_last, void *dgates, void *dc_ws,
                    void *dc0, void *dh0, const void *WhT, int ldw, unsigned *flags, int M, int C, int L, cudaStream_t st) {
  LEOD_REQUIRE(C % 16 == 0 && M > 0 && L > 0, "lstm_seq_bwd_tc: bad shape M=%d C=%d L=%d", M, C, L);
  LstmBwdArgs a;
  a.gates = (const bf16 *)gates; a.c_all = (const bf16 *)c_all; a.c0 = (const bf16 *)c0; a.dh = (const bf16 *)dh;
  a.dc_last = (const bf16 *)dc_last; a.dgates = (bf16 *)dgates; a.dc_ws = (bf16 *)dc_ws; a.dc0 = (bf16 *)dc0; a.dh0 = (bf16 *)dh0;
  a.flags = flags; a.M = M; a.C = C; a.L = L;
  // output-channel pass: the widest multiple of 16 that divides C and keeps the B stage <= 24 KB (shared-memory budget);
  // stages with few token tiles are split into more, narrower passes so that ~120 CTAs share a timestep
  a.tiles_m = ceil_div(M, TILE_M);
  const int want_pass = a.tiles_m >= 40 ? 1 : ceil_div(120, a.tiles_m);
  a.CWb = 16;
  for (int cw = std::min(C, 192); cw >= 16; cw -= 16)
    if (C % cw == 0 && C / cw >= want_pass) { a.CWb = cw; break; }
  if (a.tiles_m * (C / a.CWb) > num_sms())
    for (int cw = std::min(C, 192); cw >= 16; cw -= 16)
      if (C % cw == 0) { a.CWb = cw; break; }
  a.npass = C / a.CWb;
  a.split = (a.npass > 1 && a.tiles_m * a.npass <= num_sms()) ? 1 : 0;
  a.nkb = ceil_div(4 * C, TILE_K);
  const int stage_bytes = A_STAGE_BYTES + (int)round_up(a.CWb * TILE_K * 2, 1024);
  a.stages = std::max(2, std::min(4, (int)((224 * 1024 - 4 * LB_WARP_STAGE - 2048) / stage_bytes)));
  CUtensorMap mG, mW;
  LEOD_TRY(make_map3(&mG, dgates, 4 * C, M, L, 4 * C, (uint64_t)M * 4 * C, TILE_K, TILE_M));
  LEOD_TRY(make_map(&mW, WhT, 4 * C, C, ldw, TILE_K, a.CWb));
  const size_t smem = (size_t)a.stages * stage_bytes + 4 * LB_WARP_STAGE + 1024 + (2 * a.stages + 4) * 8 + 64;
  LEOD_REQUIRE(smem <= 227 * 1024, "lstm_seq_bwd_tc: shared memory %zu", smem);
  static bool attr_set = false;
  if (!attr_set) {
    LEOD_CUDA(cudaFuncSetAttribute(lstm_seq_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
    attr_set = true;
  }
  LEOD_TRY(device_zero_u32(flags, (int64_t)a.tiles_m * L, st));
  const int grid = a.split ? a.tiles_m * a.npass : std::min(a.tiles_m, num_sms());
  LEOD_LAUNCH((lstm_seq_bwd_kernel), grid, LS_THREADS, smem, st, mG, mW, a);
  LEOD_LAUNCH_CHECK();
  return 0;
}

int gemm_tn_tc(const void *dY, int ldy, const void *X, int ldx, float *dW, int ldw, float *dbias, int M, int N, int K,
               cudaStream_t st, const ConvTaps *taps) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (ldy % 8 != 0 || ldx % 8 != 0 || (((uintptr_t)dY | (uintptr_t)X) & 15) != 0) {
    // TMA needs 16-byte aligned rows: odd shapes (never produced by the backbone) take the SIMT kernel
    return gemm_tn_simt(LEOD_BF16, dY, ldy, X, ldx, dW, ldw, dbias, M, N, K, st, taps);
  }
  TnArgs a;
  if (taps) a.taps = *taps;
  const int ntap = a.taps.n > 1 ? a.taps.n : 1;
  a.dW = dW; a.ldw = ldw; a.M = M; a.N = N; a.K = K;
  a.dbias = dbias;
  a.BKt = K >= 256 ? 256 : (int)round_up(K, 64);
  const int tn = ceil_div(N, TILE_M), tk = ceil_div(K, a.BKt);
  // each split ends in an atomic epilogue over the whole output tile, so splits are only worth it when
  // they still stream a few thousand rows each; small-M problems run unsplit (plain read-modify-write)
  a.tiles_k = tk;
  // Backbone shapes (M up to 860k rows): each split ends in an atomic epilogue, so it should stream a few thousand rows.
  // The tapped / small-M problems of the neck and head are latency chains of 64-row k-blocks instead: split them finely enough
  // that ~2 waves of CTAs share the reduction.
  const bool fine = taps != nullptr || M <= 32768;
  int splits = ceil_div(fine ? 2 * 148 : 148, tn * tk * ntap);
  const int max_splits = fine ? std::max(1, M / 256) : (M <= 4096 ? 1 : ceil_div(M, 2048));
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  a.rows_per_split = (int)round_up(ceil_div(M, splits), 64);
  splits = ceil_div(M, a.rows_per_split);
  const int nkb = a.rows_per_split / 64;
  a.stages = nkb < 4 ? nkb : 4;
  a.tmem_cols = 32;
  while (a.tmem_cols < a.BKt + (dbias ? 64 : 0)) a.tmem_cols *= 2;
  CUtensorMap mY, mX;
  LEOD_TRY(make_map(&mY, dY, N, M, ldy, 64, 64));
  LEOD_TRY(make_map(&mX, X, K, M, ldx, 64, 64));
  const int stage_bytes = (2 + a.BKt / 64) * TN_BOX_BYTES;
  const size_t smem = (size_t)a.stages * stage_bytes + TN_BOX_BYTES + 1024 + (2 * a.stages + 2) * 8;
  static bool attr_set = false;
  if (!attr_set) {
    LEOD_CUDA(cudaFuncSetAttribute(gemm_tn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(210 * 1024)));
    attr_set = true;
  }
  dim3 grid(tn, tk * ntap, splits);
  LEOD_LAUNCH((gemm_tn_tc_kernel), grid, NUM_THREADS, smem, st, mY, mX, a);
  LEOD_LAUNCH_CHECK();
  return 0;
}
