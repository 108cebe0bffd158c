// Implicit-GEMM stem (models/layers/maxvit/maxvit.py:143-182: bias-free 7x7 stride-4 pad-3 convolution of the event tensor) for
// sm_100a, forward and weight gradient, WITHOUT a patch matrix.
//
// The explicit path (im2col_nchw + GEMM) writes a 1.9 GB bf16 patch matrix per training step of BASELINE configs[1] and reads it
// twice (forward GEMM, weight-gradient GEMM) — for a layer whose compulsory input is the 245 MB uint8 event tensor.  Here a CTA owns
// tiles of 8 x 16 output pixels.  Its 35 x 68 x Cin uint8 input window is copied into shared memory with cp.async (zero fill outside
// the frame = the zero padding of utils/padding.py:33-58 and of the convolution), double buffered across tiles; eight builder warps
// expand it into swizzle-128B bf16 operand tiles (k = (cin*7 + ky)*8 + slot, slot 0 <-> kx = -1 with a zero weight, so one 16-byte
// chunk is 8 consecutive input bytes) that tcgen05.mma consumes directly:
//   forward :  D[128 px, C]   = A[128 px, 64 k] (K-major)      x W[C, 64 k]^T          per k-block, accumulated in TMEM over 18 k-blocks
//   wgrad   :  D[128 k, C]   += A[128 px, 2 x 64 k] (MN-major) x dY[128 px, C] (MN-major) per tile, accumulated in TMEM over ALL tiles of
//              the CTA, one red.global.add pass at the end.  The k range is split over two CTA groups (9 k-blocks = channels 0..11
//              and 11..19), so each stages only its channels.
// Warps: 0 = TMA (weights / dY), 1 = MMA issuer, 2..5 = epilogue, 6..13 = input staging + operand builders.
#include <algorithm>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int ST_THREADS = 448;
constexpr int ST_BUILDERS = 256;
constexpr int ROWS_IN = 35, PITCH_IN = 68, WORDS_IN = 17;   // input window: 8*4+3 rows, 16*4+4 bytes per row
constexpr int A_TILE = 128 * 128;                            // 128 pixels x 64 bf16

struct StemArgs {
  const uint8_t *x;
  bf16 *y;          // forward output [nimg*Ho*Wo, C]
  float *dW;        // wgrad output [C, ldw] fp32, accumulated
  int ldw;
  int nimg, Cin, xh, xw, Ho, Wo, C, BN;
  int tiles_y, tiles_x, num_tiles;
  int nkb, stages, in_bytes;
};

__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src, bool valid) {
  const int sz = valid ? 4 : 0;   // src-size 0 -> zero fill
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void builders_sync() { asm volatile("bar.sync 1, %0;" ::"n"(ST_BUILDERS) : "memory"); }

// Stage the input window of `tile` (channels c0 .. c0+nc-1) into `buf`: [nc][35][68] bytes, zero outside the frame.
// thread -> (column word, row group of 3 rows): no divisions or multiplications in the loops; 255 of the 256 builders take part.
__device__ __forceinline__ void stage_input(const StemArgs &a, int tile, int c0, int nc, uint8_t *buf, int bt) {
  const int col = bt % WORDS_IN, rg = bt / WORDS_IN;
  if (rg >= 15) return;
  const int per_img = a.tiles_y * a.tiles_x;
  const int img = tile / per_img, tr = tile - img * per_img;
  const int ty = tr / a.tiles_x, tx = tr - ty * a.tiles_x;
  const int ix = tx * 64 - 4 + col * 4;
  const bool xok = ix >= 0 && ix + 3 < a.xw;
  const int plane = a.xh * a.xw;
  bool ok[3];
  int goff[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int row = rg + 15 * i, iy = ty * 32 - 3 + row;
    ok[i] = xok && row < ROWS_IN && iy >= 0 && iy < a.xh;
    goff[i] = ok[i] ? iy * a.xw + ix : 0;
  }
  const uint8_t *src = a.x + ((size_t)img * a.Cin + c0) * plane;
  uint32_t dst = smem_u32(buf) + rg * PITCH_IN + col * 4;
  for (int c = 0; c < nc; ++c, src += plane, dst += ROWS_IN * PITCH_IN) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
      if (rg + 15 * i < ROWS_IN) cp_async4(dst + 15 * i * PITCH_IN, src + goff[i], ok[i]);
  }
}
// One 16-byte chunk (8 pixels of one (cin, ky) row) of pixel row r = (py, px) into the swizzle-128B tile: chunk c of row r at
// physical chunk c ^ (r & 7).  F16 = true (forward): the operand is FP16 — counts 0..255 are exact, byte b -> half 0x6400 | b =
// 1024 + b through one PRMT per pair, minus 1024 with one HSUB2 (the weights are then fed as FP16 copies of their bf16 values: both
// operands of a kind::f16 MMA must have the same format).  F16 = false (weight gradient, whose other operand is a bf16 gradient
// that could underflow FP16): integer -> float -> bf16, exact as well.
__device__ __forceinline__ uint32_t bytes_to_half2(uint32_t w, uint32_t selector) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(0x64646464u), "r"(selector));
  const __half2 h = __hsub2(*reinterpret_cast<const __half2 *>(&d), __halves2half2(__ushort_as_half(0x6400), __ushort_as_half(0x6400)));
  return *reinterpret_cast<const uint32_t *>(&h);
}
__device__ __forceinline__ uint32_t bytes_to_bf162(uint32_t w, int shift) {
  const __nv_bfloat162 v = __floats2bfloat162_rn((float)((w >> shift) & 0xffu), (float)((w >> (shift + 8)) & 0xffu));
  return *reinterpret_cast<const uint32_t *>(&v);
}
template <bool F16>
__device__ __forceinline__ void build_chunk(const uint8_t *src, bool valid, uint8_t *dst) {
  uint32_t w0 = 0, w1 = 0;
  if (valid) {
    w0 = reinterpret_cast<const uint32_t *>(src)[0];
    w1 = reinterpret_cast<const uint32_t *>(src)[1];
  }
  uint4 v;
  if (F16) {
    v.x = bytes_to_half2(w0, 0x4140u); v.y = bytes_to_half2(w0, 0x4342u);
    v.z = bytes_to_half2(w1, 0x4140u); v.w = bytes_to_half2(w1, 0x4342u);
  } else {
    v.x = bytes_to_bf162(w0, 0); v.y = bytes_to_bf162(w0, 16);
    v.z = bytes_to_bf162(w1, 0); v.w = bytes_to_bf162(w1, 16);
  }
  *reinterpret_cast<uint4 *>(dst) = v;
}
// Per-thread walk over the patch chunks ch = 8*kb + c of successive k-blocks: (cin, ky) of chunk c advance by (1, 1) per k-block
// (8 = 7 + 1), so the byte offset of the chunk's input row in the staged window moves by a constant, with a fix-up when ky wraps.
struct ChunkWalk {
  int off[4], ky[4];
  __device__ __forceinline__ void init(int kb, int cbase, int c0, int py, int px) {
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int ch = kb * 8 + cbase + cc;
      const int cin = ch / 7;
      ky[cc] = ch - cin * 7;
      off[cc] = ((cin - c0) * ROWS_IN + py * 4 + ky[cc]) * PITCH_IN + px * 4;
    }
  }
  __device__ __forceinline__ void next() {
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      ky[cc] += 1;
      off[cc] += (ROWS_IN + 1) * PITCH_IN;
      if (ky[cc] >= 7) {
        ky[cc] -= 7;
        off[cc] += (ROWS_IN - 7) * PITCH_IN;
      }
    }
  }
};
// instruction descriptor of kind::f16 with A = B = FP16 (format 0), D = FP32
__device__ __forceinline__ uint32_t idesc_f16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__global__ void bf16_to_f16_kernel(const bf16 *__restrict__ src, __half *__restrict__ dst, int64_t n) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = __float2half_rn(__bfloat162float(src[i]));
}

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(ST_THREADS, 1) stem_fwd_kernel(const __grid_constant__ CUtensorMap mapW, const StemArgs a) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int b_bytes = a.BN * 128;
  const int stage_bytes = A_TILE + b_bytes;
  uint8_t *ring = smem;
  uint8_t *inbuf = smem + (size_t)a.stages * stage_bytes;
  uint64_t *bars = (uint64_t *)(inbuf + 2 * (size_t)a.in_bytes);
  uint64_t *full_bar = bars, *empty_bar = bars + a.stages, *tmem_full = bars + 2 * a.stages, *tmem_empty = tmem_full + 2;
  uint32_t *tmem_slot = (uint32_t *)(tmem_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1 + ST_BUILDERS / 32);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tmem_full[b]), 1);
      mbar_init(smem_u32(&tmem_empty[b]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int acc_cols = (a.BN + 31) / 32 * 32;
  const int tmem_cols = acc_cols <= 32 ? 64 : 128;   // two accumulators
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x)
        for (int kb = 0; kb < a.nkb; ++kb, ++it) {
          const int s = it % a.stages;
          mbar_wait_relaxed(smem_u32(&empty_bar[s]), ((it / a.stages) & 1) ^ 1);
          const uint32_t fb = smem_u32(&full_bar[s]);
          mbar_expect_tx(fb, b_bytes);
          tma_load_2d(smem_u32(ring + (size_t)s * stage_bytes + A_TILE), &mapW, fb, kb * 64, 0);
        }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_f16(128, a.BN, 0, 0);
      int it = 0, j = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++j) {
        const int buf = j & 1;
        mbar_wait_relaxed(smem_u32(&tmem_empty[buf]), ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(buf * acc_cols);
        for (int kb = 0; kb < a.nkb; ++kb, ++it) {
          const int s = it % a.stages;
          mbar_wait_relaxed(smem_u32(&full_bar[s]), (it / a.stages) & 1, 32);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes), sb = sa + A_TILE;
          const uint64_t adesc = make_desc(sa, 16, 1024), bdesc = make_desc(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_bf16(acc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          tc_commit(smem_u32(&empty_bar[s]));
        }
        tc_commit(smem_u32(&tmem_full[buf]));
      }
    }
  } else if (warp < 6) {
    // epilogue: TMEM lane = pixel of the tile; every thread writes its pixel's C channels (C*2 contiguous bytes)
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane, py = r >> 4, px = r & 15;
    const int per_img = a.tiles_y * a.tiles_x;
    int j = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++j) {
      const int buf = j & 1;
      const int img = tile / per_img, tr = tile - img * per_img;
      const int ty = tr / a.tiles_x, tx = tr - ty * a.tiles_x;
      const size_t m = ((size_t)img * a.Ho + ty * 8 + py) * a.Wo + tx * 16 + px;
      mbar_wait_relaxed(smem_u32(&tmem_full[buf]), (j >> 1) & 1, 256);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * acc_cols);
      bf16 *dst = a.y + m * a.C;
      for (int c = 0; c < a.BN; c += 16) {
        uint32_t rr[16];
        tmem_ld16(trow + (uint32_t)c, rr);
        uint4 o[2];
        uint32_t *ow = reinterpret_cast<uint32_t *>(o);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const __nv_bfloat162 pr = __floats2bfloat162_rn(__uint_as_float(rr[2 * q]), __uint_as_float(rr[2 * q + 1]));
          ow[q] = *reinterpret_cast<const uint32_t *>(&pr);
        }
        if (c + 16 <= a.C) {
          reinterpret_cast<uint4 *>(dst + c)[0] = o[0];
          reinterpret_cast<uint4 *>(dst + c)[1] = o[1];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[buf]));
    }
  } else {
    // builders: stage the next tile's input with cp.async, expand the current one into the A tiles of the ring
    const int bt = threadIdx.x - 6 * 32;
    const int r = bt & 127, half = bt >> 7, py = r >> 4, px = r & 15;
    const int nch = a.Cin * 7;
    int dsto[4];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) dsto[cc] = r * 128 + (((half * 4 + cc) ^ (r & 7)) << 4);
    int it = 0, j = 0;
    stage_input(a, blockIdx.x, 0, a.Cin, inbuf, bt);
    cp_async_commit();
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++j) {
      const int next = tile + gridDim.x;
      if (next < a.num_tiles) stage_input(a, next, 0, a.Cin, inbuf + (size_t)((j + 1) & 1) * a.in_bytes, bt);
      cp_async_commit();
      cp_async_wait<1>();
      builders_sync();                       // every builder's share of this tile's window has landed
      const uint8_t *in = inbuf + (size_t)(j & 1) * a.in_bytes;
      ChunkWalk cw;
      cw.init(0, half * 4, 0, py, px);
      for (int kb = 0; kb < a.nkb; ++kb, ++it) {
        const int s = it % a.stages;
        mbar_wait(smem_u32(&empty_bar[s]), ((it / a.stages) & 1) ^ 1);
        uint8_t *tileA = ring + (size_t)s * stage_bytes;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) build_chunk<true>(in + cw.off[cc], kb * 8 + half * 4 + cc < nch, tileA + dsto[cc]);
        cw.next();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&full_bar[s]));
      }
      builders_sync();                       // the window buffer may be overwritten by the tile after next
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ weight gradient
// blockIdx.y = k half: 0 -> k-block pairs 0..4 (channels 0..11), 1 -> pairs 5..8 (channels 11..19).  dW[c, k] += sum_px dY[px, c] * patch[px, k].
__global__ void __launch_bounds__(ST_THREADS, 1) stem_wgrad_kernel(const __grid_constant__ CUtensorMap mapDY, const StemArgs a) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int half = blockIdx.y;
  const int npair_all = (a.nkb + 1) / 2;
  const int pair0 = half == 0 ? 0 : (npair_all + 1) / 2, pair1 = half == 0 ? (npair_all + 1) / 2 : npair_all;
  const int npair = pair1 - pair0;
  const int c0 = (pair0 * 16) / 7;                                             // first / last channel touched by k-blocks 2*pair0 .. 2*pair1-1
  const int c_last = min(a.Cin - 1, (pair1 * 16 - 1) / 7);
  const int nc = c_last - c0 + 1;
  const int pair_bytes = 2 * A_TILE;
  uint8_t *ring = smem;                                                        // [stages][2 tiles of 128 px x 64 k]
  uint8_t *dyb = ring + (size_t)a.stages * pair_bytes;                         // [2][128 px x 64 c]
  uint8_t *inbuf = dyb + 2 * A_TILE;
  uint64_t *bars = (uint64_t *)(inbuf + 2 * (size_t)a.in_bytes);
  uint64_t *full_bar = bars, *empty_bar = bars + a.stages, *dy_full = bars + 2 * a.stages, *dy_empty = dy_full + 2, *tmem_full = dy_empty + 2;
  uint32_t *tmem_slot = (uint32_t *)(tmem_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), ST_BUILDERS / 32);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&dy_full[b]), 1);
      mbar_init(smem_u32(&dy_empty[b]), 1);
    }
    mbar_init(smem_u32(tmem_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr int ACC = 64;        // accumulator column pitch (BN <= 64)
  const int tmem_cols = 512;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const int per_img = a.tiles_y * a.tiles_x;
  const bool any = (int)blockIdx.x < a.num_tiles;

  if (warp == 0) {
    if (lane == 0) {
      int j = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++j) {
        const int d = j & 1;
        mbar_wait_relaxed(smem_u32(&dy_empty[d]), ((j >> 1) & 1) ^ 1);
        const uint32_t fb = smem_u32(&dy_full[d]);
        mbar_expect_tx(fb, A_TILE);
        const int img = tile / per_img, tr = tile - img * per_img;
        const int ty = tr / a.tiles_x, tx = tr - ty * a.tiles_x;
        for (int seg = 0; seg < 8; ++seg) {      // 8 output rows of 16 consecutive pixels each
          const int m = (img * a.Ho + ty * 8 + seg) * a.Wo + tx * 16;
          tma_load_2d(smem_u32(dyb + (size_t)d * A_TILE + seg * 2048), &mapDY, fb, 0, m);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, a.BN, 1, 1);
      int it = 0, j = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++j) {
        const int d = j & 1;
        mbar_wait_relaxed(smem_u32(&dy_full[d]), (j >> 1) & 1, 32);
        tc_fence_after();
        const uint64_t bdesc = make_desc(smem_u32(dyb + (size_t)d * A_TILE), A_TILE, 1024);
        for (int p = 0; p < npair; ++p, ++it) {
          const int s = it % a.stages;
          mbar_wait_relaxed(smem_u32(&full_bar[s]), (it / a.stages) & 1, 32);
          tc_fence_after();
          const uint64_t adesc = make_desc(smem_u32(ring + (size_t)s * pair_bytes), A_TILE, 1024);
#pragma unroll
          for (int k = 0; k < 8; ++k)            // 16 pixels per MMA = 16 rows of 128 B (>>4: +128)
            tc_mma_bf16(tmem_base + (uint32_t)(p * ACC), adesc + 128 * k, bdesc + 128 * k, idesc, (j | k) != 0);
          tc_commit(smem_u32(&empty_bar[s]));
        }
        tc_commit(smem_u32(&dy_empty[d]));
      }
      tc_commit(smem_u32(tmem_full));
    }
  } else if (warp < 6) {
    if (any) {
      const int quarter = warp & 3;
      mbar_wait_relaxed(smem_u32(tmem_full), 0, 1000);
      tc_fence_after();
      const int K = a.Cin * 56;
      for (int p = 0; p < npair; ++p) {
        const int k = (pair0 + p) * 128 + quarter * 32 + lane;       // accumulator row = patch element
        for (int c = 0; c < a.BN; c += 16) {
          uint32_t rr[16];
          tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(p * ACC + c), rr);
          if (k < K)
#pragma unroll
            for (int q = 0; q < 16; ++q)
              if (c + q < a.C) atomicAdd(a.dW + (size_t)(c + q) * a.ldw + k, __uint_as_float(rr[q]));
        }
      }
    }
  } else {
    const int bt = threadIdx.x - 6 * 32;
    const int r = bt & 127, hf = bt >> 7, py = r >> 4, px = r & 15;
    const int nch = a.Cin * 7;
    int dsto[4];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) dsto[cc] = r * 128 + (((hf * 4 + cc) ^ (r & 7)) << 4);
    int it = 0, j = 0;
    if (any) stage_input(a, blockIdx.x, c0, nc, inbuf, bt);
    cp_async_commit();
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++j) {
      const int next = tile + gridDim.x;
      if (next < a.num_tiles) stage_input(a, next, c0, nc, inbuf + (size_t)((j + 1) & 1) * a.in_bytes, bt);
      cp_async_commit();
      cp_async_wait<1>();
      builders_sync();
      const uint8_t *in = inbuf + (size_t)(j & 1) * a.in_bytes;
      ChunkWalk cw;
      cw.init(pair0 * 2, hf * 4, c0, py, px);
      for (int p = 0; p < npair; ++p, ++it) {
        const int s = it % a.stages;
        mbar_wait(smem_u32(&empty_bar[s]), ((it / a.stages) & 1) ^ 1);
        uint8_t *pairA = ring + (size_t)s * pair_bytes;
#pragma unroll
        for (int t2 = 0; t2 < 2; ++t2) {
          const int kb = (pair0 + p) * 2 + t2;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) build_chunk<false>(in + cw.off[cc], kb * 8 + hf * 4 + cc < nch, pairA + t2 * A_TILE + dsto[cc]);
          cw.next();
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&full_bar[s]));
      }
      builders_sync();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

int num_sms_() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

int fill_args(StemArgs *a, const uint8_t *x, int nimg, int Cin, int xh, int xw, int Ho, int Wo, int C) {
  a->x = x; a->nimg = nimg; a->Cin = Cin; a->xh = xh; a->xw = xw; a->Ho = Ho; a->Wo = Wo; a->C = C;
  a->BN = (int)round_up(C, 16);
  a->tiles_y = Ho / 8; a->tiles_x = Wo / 16;
  a->num_tiles = nimg * a->tiles_y * a->tiles_x;
  a->nkb = ceil_div(Cin * 56, 64);
  return 0;
}

}  // namespace

bool stem_implicit_supported(int Cin, int xh, int xw, int Ho, int Wo, int C, const void *x) {
  return Ho % 8 == 0 && Wo % 16 == 0 && (xw & 3) == 0 && (((uintptr_t)x) & 3) == 0 && C % 16 == 0 && C <= 64 && Cin * 7 >= 16 && Cin <= 32 &&
         xh <= Ho * 4 && xw <= Wo * 4;
}

// FP16 copy of the prepared (bf16-valued) stem weight for the forward kernel
int stem_weight_to_f16(const void *W_bf16, void *W_f16, int64_t n, cudaStream_t st) {
  LEOD_LAUNCH((bf16_to_f16_kernel), (int)std::min<int64_t>((n + 255) / 256, 148 * 4), 256, 0, st, (const bf16 *)W_bf16, (__half *)W_f16, n);
  LEOD_LAUNCH_CHECK();
  return 0;
}

// y[nimg*Ho*Wo, C] = conv7x7s4p3(x) with W = prepared weight [C, Cin*56] (patch order, FP16 copy of the bf16 values)
int stem_fwd_tc(const uint8_t *x, int nimg, int Cin, int xh, int xw, int Ho, int Wo, int C, const void *W, int ldw, void *y, cudaStream_t st) {
  StemArgs a;
  fill_args(&a, x, nimg, Cin, xh, xw, Ho, Wo, C);
  a.y = (bf16 *)y; a.dW = nullptr; a.ldw = 0;
  a.in_bytes = (int)round_up((int64_t)Cin * ROWS_IN * PITCH_IN, 128);
  const int stage_bytes = A_TILE + a.BN * 128;
  a.stages = std::min(4, (int)((220 * 1024 - 2 * a.in_bytes - 2048) / stage_bytes));
  LEOD_REQUIRE(a.stages >= 2, "stem_fwd_tc: shared memory (Cin %d)", Cin);
  CUtensorMap mW;
  LEOD_TRY(tc_make_map_2d(&mW, W, (uint64_t)Cin * 56, C, ldw, 64, a.BN));
  const size_t smem = (size_t)a.stages * stage_bytes + 2 * (size_t)a.in_bytes + 1024 + (2 * a.stages + 4) * 8 + 64;
  static bool attr_set = false;
  if (!attr_set) {
    LEOD_CUDA(cudaFuncSetAttribute(stem_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
    attr_set = true;
  }
  const int waves = ceil_div(a.num_tiles, num_sms_());
  LEOD_LAUNCH((stem_fwd_kernel), ceil_div(a.num_tiles, waves), ST_THREADS, smem, st, mW, a);
  LEOD_LAUNCH_CHECK();
  return 0;
}

// dW[C, ldw] (fp32, patch order) += dY[nimg*Ho*Wo, C]^T * patches(x)
int stem_wgrad_tc(const uint8_t *x, int nimg, int Cin, int xh, int xw, int Ho, int Wo, int C, const void *dY, float *dW, int ldw, cudaStream_t st) {
  StemArgs a;
  fill_args(&a, x, nimg, Cin, xh, xw, Ho, Wo, C);
  a.y = nullptr; a.dW = dW; a.ldw = ldw;
  const int npair_all = (a.nkb + 1) / 2;
  LEOD_REQUIRE((npair_all + 1) / 2 * 64 <= 512, "stem_wgrad_tc: %d k-blocks exceed the TMEM accumulators", a.nkb);
  const int max_nc = std::min(Cin, ((npair_all + 1) / 2 * 16 + 6) / 7 + 1);
  a.in_bytes = (int)round_up((int64_t)max_nc * ROWS_IN * PITCH_IN, 128);
  a.stages = std::min(3, (int)((220 * 1024 - 2 * a.in_bytes - 2 * A_TILE - 2048) / (2 * A_TILE)));
  LEOD_REQUIRE(a.stages >= 2, "stem_wgrad_tc: shared memory (Cin %d)", Cin);
  CUtensorMap mDY;
  LEOD_TRY(tc_make_map_2d(&mDY, dY, C, (uint64_t)nimg * Ho * Wo, C, 64, 16));
  const size_t smem = (size_t)a.stages * 2 * A_TILE + 2 * A_TILE + 2 * (size_t)a.in_bytes + 1024 + (2 * a.stages + 5) * 8 + 64;
  static bool attr_set = false;
  if (!attr_set) {
    LEOD_CUDA(cudaFuncSetAttribute(stem_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
    attr_set = true;
  }
  const int per_half = std::max(1, std::min(a.num_tiles, num_sms_() / 2));
  LEOD_LAUNCH((stem_wgrad_kernel), dim3(per_half, 2), ST_THREADS, smem, st, mDY, a);
  LEOD_LAUNCH_CHECK();
  return 0;
}
