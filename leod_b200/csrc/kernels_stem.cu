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
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

// Warps: 0 = TMA, 1 = MMA issuer, 2..5 = epilogue, then G groups of four builder warps.  A builder warp is latency-bound (dependent
// PRMT -> HSUB2 -> STS chains, the proxy fence), so the kernels want many of them: each group builds every G-th k-block (pair).
constexpr int FWD_GROUPS = 4, WG_GROUPS = 3;
constexpr int FWD_THREADS = 192 + FWD_GROUPS * 128, WG_THREADS = 192 + WG_GROUPS * 128;
// input window of a tile: 8*4+3 rows; per row the 80 bytes x = tx*64-16 .. tx*64+63 (the left halo is 4 bytes, the window starts 16
// bytes early so that every row is five 16-byte cp.async pieces; an 80-byte pitch also makes the builders' 32-bit reads conflict-free:
// a warp covers two pixel rows = 4 window rows = 80 words apart = 16 banks)
constexpr int ROWS_IN = 35, PITCH_IN = 80, SEGS_IN = 5, X_HALO = 12;
constexpr int CH_BYTES = ROWS_IN * PITCH_IN;
constexpr int A_TILE = 128 * 128;                            // 128 pixels x 64 bf16
constexpr int MAX_CHUNKS = 256;                              // Cin <= 32: 224 (cin, ky) chunks + the padding of the last k-block

struct StemArgs {
  const uint8_t *x;
  bf16 *y;          // forward output [nimg*Ho*Wo, C]
  float *dW;        // wgrad output [C, ldw] fp32, accumulated
  int ldw;
  int nimg, Cin, xh, xw, Ho, Wo, C, BN;
  int tiles_y, tiles_x, num_tiles;
  int nkb, stages, stages_b, in_bytes;
  int dbg;          // LEOD_STEM_DEBUG bit mask (timing experiments only; results are wrong when set)
};

// byte offset of chunk ch = cin*7 + ky (8 consecutive input bytes per output pixel) inside the staged window
struct ChunkTab { int v[MAX_CHUNKS]; };
constexpr ChunkTab make_chunk_tab() {
  ChunkTab t{};
  for (int ch = 0; ch < MAX_CHUNKS; ++ch) t.v[ch] = ((ch / 7) * ROWS_IN + ch % 7) * PITCH_IN;
  return t;
}
__constant__ ChunkTab c_chunk = make_chunk_tab();

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, bool valid) {
  const int sz = valid ? 16 : 0;   // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
template <int G> __device__ __forceinline__ void builders_sync() { asm volatile("bar.sync 1, %0;" ::"n"(G * 128) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

// Stage the input window of `tile` (channels c0 .. c0+nc-1) into the shared buffer at `buf` (shared-space address): [nc][35][80]
// bytes, zero outside the frame.  One (row, 16-byte piece) per thread, the first 175 builders, one cp.async per channel.
__device__ __forceinline__ void stage_input(const StemArgs &a, int tile, int c0, int nc, uint32_t buf, int bt) {
  if (bt >= ROWS_IN * SEGS_IN) return;
  const int row = bt / SEGS_IN, seg = bt - row * SEGS_IN;
  const int per_img = a.tiles_y * a.tiles_x;
  const int img = tile / per_img, tr = tile - img * per_img;
  const int ty = tr / a.tiles_x, tx = tr - ty * a.tiles_x;
  const int ix = tx * 64 - 16 + seg * 16, iy = ty * 32 - 3 + row;
  const bool ok = ix >= 0 && ix + 15 < a.xw && iy >= 0 && iy < a.xh;
  const size_t plane = (size_t)a.xh * a.xw;
  const uint8_t *src = a.x + ((size_t)img * a.Cin + c0) * plane + (ok ? iy * a.xw + ix : 0);
  uint32_t dst = buf + row * PITCH_IN + seg * 16;
#pragma unroll 4
  for (int c = 0; c < nc; ++c, src += plane, dst += CH_BYTES) cp_async16(dst, src, ok);
}
// 8 consecutive input bytes (two words) -> one 16-byte chunk of the operand tile.
// Forward: the operand is FP16 — counts 0..255 are exact, byte b -> half 0x6400 | b = 1024 + b through one PRMT per pair, minus 1024
// with one HSUB2 (the weights are then fed as FP16 copies of their bf16 values: both operands of a kind::f16 MMA must have the same
// format).  Weight gradient (its other operand is a bf16 gradient that could underflow FP16): byte b -> float 0x4B000000 | b =
// 2^23 + b, minus 2^23 = b exactly, whose low 16 bits are zero, so the upper halves of two floats ARE the bf16 pair.
__device__ __forceinline__ uint32_t bytes_to_half2(uint32_t w, uint32_t selector) {
  const uint32_t d = prmt(w, 0x64646464u, selector);
  const __half2 h = __hsub2(*reinterpret_cast<const __half2 *>(&d), __halves2half2(__ushort_as_half(0x6400), __ushort_as_half(0x6400)));
  return *reinterpret_cast<const uint32_t *>(&h);
}
__device__ __forceinline__ uint32_t bytes_to_bf162(uint32_t w, uint32_t sel0, uint32_t sel1) {
  const float f0 = __uint_as_float(prmt(w, 0x4B000000u, sel0)) - 8388608.0f;
  const float f1 = __uint_as_float(prmt(w, 0x4B000000u, sel1)) - 8388608.0f;
  return prmt(__float_as_uint(f0), __float_as_uint(f1), 0x7632u);
}
template <bool F16>
__device__ __forceinline__ void build_chunk(uint32_t w0, uint32_t w1, uint32_t dst) {
  if (F16)
    sts128(dst, bytes_to_half2(w0, 0x4140u), bytes_to_half2(w0, 0x4342u), bytes_to_half2(w1, 0x4140u), bytes_to_half2(w1, 0x4342u));
  else
    sts128(dst, bytes_to_bf162(w0, 0x7540u, 0x7541u), bytes_to_bf162(w0, 0x7542u, 0x7543u), bytes_to_bf162(w1, 0x7540u, 0x7541u),
           bytes_to_bf162(w1, 0x7542u, 0x7543u));
}
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
// Two rings: the A tiles built on chip (16 KB each, `stages` deep) and the weight tiles fetched by TMA (BN x 128 B each, `stages_b`
// deep — deeper, because a 6 KB weight tile is consumed in ~0.15 us while a TMA round trip takes ~1 us).
__global__ void __launch_bounds__(FWD_THREADS, 1) stem_fwd_kernel(const __grid_constant__ CUtensorMap mapW, const StemArgs a) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int b_bytes = a.BN * 128;
  uint8_t *ringA = smem;
  uint8_t *ringB = ringA + (size_t)a.stages * A_TILE;
  uint8_t *inbuf = ringB + (size_t)a.stages_b * b_bytes;
  uint64_t *bars = (uint64_t *)(inbuf + 2 * (size_t)a.in_bytes);
  uint64_t *fullA = bars, *emptyA = fullA + a.stages, *fullB = emptyA + a.stages, *emptyB = fullB + a.stages_b;
  uint64_t *tmem_full = emptyB + a.stages_b, *tmem_empty = tmem_full + 2;
  uint32_t *tmem_slot = (uint32_t *)(tmem_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(smem_u32(&fullA[s]), 4);   // one builder group (4 warps) per k-block
      mbar_init(smem_u32(&emptyA[s]), 1);
    }
    for (int s = 0; s < a.stages_b; ++s) {
      mbar_init(smem_u32(&fullB[s]), 1);
      mbar_init(smem_u32(&emptyB[s]), 1);
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
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x)
        for (int kb = 0; kb < a.nkb; ++kb) {
          if (!(a.dbg & 128)) mbar_wait_relaxed(smem_u32(&emptyB[s]), ph ^ 1, 64);
          const uint32_t fb = smem_u32(&fullB[s]);
          if (a.dbg & 4) {
            mbar_arrive(fb);
          } else {
            mbar_expect_tx(fb, b_bytes);
            tma_load_2d(smem_u32(ringB + (size_t)s * b_bytes), &mapW, fb, kb * 64, 0);
          }
          if (++s == a.stages_b) { s = 0; ph ^= 1; }
        }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_f16(128, a.BN, 0, 0);
      int sa = 0, sb = 0, j = 0;
      uint32_t pa = 0, pb = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++j) {
        const int buf = j & 1;
        mbar_wait_relaxed(smem_u32(&tmem_empty[buf]), ((j >> 1) & 1) ^ 1, 32);
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(buf * acc_cols);
        for (int kb = 0; kb < a.nkb; ++kb) {
          mbar_wait(smem_u32(&fullB[sb]), pb);
          mbar_wait(smem_u32(&fullA[sa]), pa);
          tc_fence_after();
          const uint64_t adesc = make_desc(smem_u32(ringA + (size_t)sa * A_TILE), 16, 1024);
          const uint64_t bdesc = make_desc(smem_u32(ringB + (size_t)sb * b_bytes), 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (!(a.dbg & 2)) tc_mma_bf16(acc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          if (!(a.dbg & 256)) tc_commit(smem_u32(&emptyA[sa]));
          if (!(a.dbg & 128)) tc_commit(smem_u32(&emptyB[sb]));
          if (++sa == a.stages) { sa = 0; pa ^= 1; }
          if (++sb == a.stages_b) { sb = 0; pb ^= 1; }
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
      mbar_wait_relaxed(smem_u32(&tmem_full[buf]), (j >> 1) & 1, 512);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * acc_cols);
      bf16 *dst = a.y + m * a.C;
      for (int c = 0; c < a.BN; c += 16) {
        uint32_t rr[16];
        if (!(a.dbg & 512)) tmem_ld16(trow + (uint32_t)c, rr);
        uint4 o[2];
        uint32_t *ow = reinterpret_cast<uint32_t *>(o);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const __nv_bfloat162 pr = __floats2bfloat162_rn(__uint_as_float(rr[2 * q]), __uint_as_float(rr[2 * q + 1]));
          ow[q] = *reinterpret_cast<const uint32_t *>(&pr);
        }
        if (c + 16 <= a.C && !(a.dbg & 32)) {
          reinterpret_cast<uint4 *>(dst + c)[0] = o[0];
          reinterpret_cast<uint4 *>(dst + c)[1] = o[1];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[buf]));
    }
  } else {
    // builders: stage the next tile's input with cp.async, expand the current one into the A tiles of the ring.  Group g builds the
    // k-blocks whose running index is g modulo FWD_GROUPS (a thread builds the whole 128-byte row of its pixel), so that many k-blocks
    // are in flight and the store -> proxy fence -> arrive chain of one overlaps the loads and conversions of the others.  Chunks beyond Cin*7 (the tail
    // of the last k-block) read whatever follows the window in shared memory: finite FP16 values that meet the zero weights of the
    // TMA out-of-bounds fill.
    const int bt = threadIdx.x - 6 * 32;
    const int r = bt & 127, grp = bt >> 7, py = r >> 4, px = r & 15;
    const uint32_t in_s = smem_u32(inbuf), ring_s = smem_u32(ringA);
    const uint32_t thread_off = (uint32_t)(py * 4 * PITCH_IN + X_HALO + px * 4);
    const uint32_t row_off = (uint32_t)(r * 128), sw = (uint32_t)(r & 7);
    int s = grp % a.stages, j = 0, it0 = 0;   // it0: running index of the tile's first k-block
    uint32_t ph = (uint32_t)(grp / a.stages) & 1u;
    stage_input(a, blockIdx.x, 0, a.Cin, in_s, bt);
    cp_async_commit();
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++j, it0 += a.nkb) {
      const int next = tile + gridDim.x;
      if (next < a.num_tiles && !(a.dbg & 16)) stage_input(a, next, 0, a.Cin, in_s + (uint32_t)(((j + 1) & 1) * a.in_bytes), bt);
      cp_async_commit();
      cp_async_wait<1>();
      builders_sync<FWD_GROUPS>();           // every builder's share of this tile's window has landed
      const uint32_t in = in_s + (uint32_t)((j & 1) * a.in_bytes) + thread_off;
      for (int kb = (grp - it0) & (FWD_GROUPS - 1); kb < a.nkb; kb += FWD_GROUPS) {
        const int *tab = c_chunk.v + kb * 8;
        uint32_t w[16];
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {     // the window is read-only here: load before waiting for the ring slot
          const uint32_t p = in + (uint32_t)tab[cc];
          if (a.dbg & 64) {
            w[2 * cc] = p; w[2 * cc + 1] = p;
          } else {
            w[2 * cc] = lds32(p);
            w[2 * cc + 1] = lds32(p + 4);
          }
        }
        if (!(a.dbg & 256)) mbar_wait(smem_u32(&emptyA[s]), ph ^ 1);
        const uint32_t rowA = ring_s + (uint32_t)(s * A_TILE) + row_off;
        if (!(a.dbg & 1)) {
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) build_chunk<true>(w[2 * cc], w[2 * cc + 1], rowA + (((uint32_t)cc ^ sw) << 4));
        }
        if (!(a.dbg & 8)) fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&fullA[s]));
        s += FWD_GROUPS;
        while (s >= a.stages) { s -= a.stages; ph ^= 1; }
      }
      builders_sync<FWD_GROUPS>();           // the window buffer may be overwritten by the tile after next
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
__global__ void __launch_bounds__(WG_THREADS, 1) stem_wgrad_kernel(const __grid_constant__ CUtensorMap mapDY, const StemArgs a) {
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
      mbar_init(smem_u32(&full_bar[s]), 4);   // one builder group (4 warps) per k-block pair
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
        mbar_wait_relaxed(smem_u32(&dy_empty[d]), ((j >> 1) & 1) ^ 1, 256);
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
      int s = 0, j = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++j) {
        const int d = j & 1;
        mbar_wait(smem_u32(&dy_full[d]), (j >> 1) & 1);
        tc_fence_after();
        const uint64_t bdesc = make_desc(smem_u32(dyb + (size_t)d * A_TILE), A_TILE, 1024);
        for (int p = 0; p < npair; ++p) {
          mbar_wait(smem_u32(&full_bar[s]), ph);
          tc_fence_after();
          const uint64_t adesc = make_desc(smem_u32(ring + (size_t)s * pair_bytes), A_TILE, 1024);
#pragma unroll
          for (int k = 0; k < 8; ++k)            // 16 pixels per MMA = 16 rows of 128 B (>>4: +128)
            tc_mma_bf16(tmem_base + (uint32_t)(p * ACC), adesc + 128 * k, bdesc + 128 * k, idesc, (j | k) != 0);
          tc_commit(smem_u32(&empty_bar[s]));
          if (++s == a.stages) { s = 0; ph ^= 1; }
        }
        tc_commit(smem_u32(&dy_empty[d]));
      }
      tc_commit(smem_u32(tmem_full));
    }
  } else if (warp < 6) {
    if (any) {
      const int quarter = warp & 3;
      mbar_wait_relaxed(smem_u32(tmem_full), 0, 2000);
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
    // builders (see the forward kernel: the two groups build alternate k-block PAIRS); accumulator rows k >= Cin*56 hold products
    // with whatever follows the window and are dropped by the epilogue
    const int bt = threadIdx.x - 6 * 32;
    const int r = bt & 127, grp = bt >> 7, py = r >> 4, px = r & 15;
    const uint32_t in_s = smem_u32(inbuf), ring_s = smem_u32(ring);
    const uint32_t thread_off = (uint32_t)(py * 4 * PITCH_IN + X_HALO + px * 4) - (uint32_t)(c0 * CH_BYTES);
    const uint32_t row_off = (uint32_t)(r * 128), sw = (uint32_t)(r & 7);
    int s = grp % a.stages, j = 0, it0 = 0;
    uint32_t ph = (uint32_t)(grp / a.stages) & 1u;
    if (any) stage_input(a, blockIdx.x, c0, nc, in_s, bt);
    cp_async_commit();
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++j, it0 += npair) {
      const int next = tile + gridDim.x;
      if (next < a.num_tiles) stage_input(a, next, c0, nc, in_s + (uint32_t)(((j + 1) & 1) * a.in_bytes), bt);
      cp_async_commit();
      cp_async_wait<1>();
      builders_sync<WG_GROUPS>();
      const uint32_t in = in_s + (uint32_t)((j & 1) * a.in_bytes) + thread_off;
      for (int p = (grp + WG_GROUPS - it0 % WG_GROUPS) % WG_GROUPS; p < npair; p += WG_GROUPS) {
        const int *tab = c_chunk.v + (pair0 + p) * 16;
        const uint32_t rowA = ring_s + (uint32_t)(s * pair_bytes) + row_off;
        bool waited = false;
#pragma unroll
        for (int t2 = 0; t2 < 2; ++t2) {
          uint32_t w[16];
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) {
            const uint32_t pa = in + (uint32_t)tab[t2 * 8 + cc];
            w[2 * cc] = lds32(pa);
            w[2 * cc + 1] = lds32(pa + 4);
          }
          if (!waited) {
            mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
            waited = true;
          }
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) build_chunk<false>(w[2 * cc], w[2 * cc + 1], rowA + (uint32_t)(t2 * A_TILE) + (((uint32_t)cc ^ sw) << 4));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&full_bar[s]));
        s += WG_GROUPS;
        while (s >= a.stages) { s -= a.stages; ph ^= 1; }
      }
      builders_sync<WG_GROUPS>();
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
  static int dbg = -1;
  if (dbg < 0) {
    const char *e = getenv("LEOD_STEM_DEBUG");
    dbg = e ? atoi(e) : 0;
  }
  a->dbg = dbg;
  return 0;
}



constexpr int SMEM_BUDGET = 225 * 1024;
constexpr int TAIL_SLACK = 4096;   // the chunks past Cin*7 read up to ~2.6 KB beyond the second window buffer

// shared-memory plan of the forward kernel: A ring 3-4 deep, weight ring as deep as fits (<= 8)
bool fwd_plan(StemArgs *a, size_t *smem) {
  a->in_bytes = (int)round_up((int64_t)a->Cin * CH_BYTES, 128);
  const int b_bytes = a->BN * 128;
  const int64_t fixed = 2 * (int64_t)a->in_bytes + TAIL_SLACK + 1024 + 512;
  for (int sa = 4; sa >= 2; --sa) {
    const int64_t left = SMEM_BUDGET - fixed - (int64_t)sa * A_TILE;
    const int sb = (int)std::min<int64_t>(8, left / b_bytes);
    if (sb >= sa) {
      a->stages = sa; a->stages_b = sb;
      *smem = (size_t)(fixed + (int64_t)sa * A_TILE + (int64_t)sb * b_bytes);
      return true;
    }
  }
  return false;
}
bool wgrad_plan(StemArgs *a, size_t *smem) {
  const int npair_all = (a->nkb + 1) / 2;
  if ((npair_all + 1) / 2 * 64 > 512) return false;                       // TMEM accumulators
  const int max_nc = std::min(a->Cin, ((npair_all + 1) / 2 * 16 + 6) / 7 + 1);
  a->in_bytes = (int)round_up((int64_t)max_nc * CH_BYTES, 128);
  const int64_t fixed = 2 * (int64_t)a->in_bytes + 2 * A_TILE + TAIL_SLACK + 1024 + 512;
  a->stages = (int)std::min<int64_t>(3, (SMEM_BUDGET - fixed) / (2 * A_TILE));
  a->stages_b = 0;
  *smem = (size_t)(fixed + (int64_t)a->stages * 2 * A_TILE);
  return a->stages >= 2;
}

}  // namespace

bool stem_implicit_supported(int Cin, int xh, int xw, int Ho, int Wo, int C, const void *x) {
  if (!(Ho % 8 == 0 && Wo % 16 == 0 && (xw & 15) == 0 && (((uintptr_t)x) & 15) == 0 && C % 16 == 0 && C <= 64 && Cin * 7 >= 16 && Cin <= 32 &&
        xh <= Ho * 4 && xw <= Wo * 4))
    return false;
  StemArgs a;
  size_t smem;
  fill_args(&a, (const uint8_t *)x, 1, Cin, xh, xw, Ho, Wo, C);
  return fwd_plan(&a, &smem) && wgrad_plan(&a, &smem);
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
  size_t smem = 0;
  LEOD_REQUIRE(fwd_plan(&a, &smem), "stem_fwd_tc: shared memory (Cin %d)", Cin);
  CUtensorMap mW;
  LEOD_TRY(tc_make_map_2d(&mW, W, (uint64_t)Cin * 56, C, ldw, 64, a.BN));
  static bool attr_set = false;
  if (!attr_set) {
    LEOD_CUDA(cudaFuncSetAttribute(stem_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
    attr_set = true;
  }
  const int waves = ceil_div(a.num_tiles, num_sms_());
  LEOD_LAUNCH((stem_fwd_kernel), ceil_div(a.num_tiles, waves), FWD_THREADS, smem, st, mW, a);
  LEOD_LAUNCH_CHECK();
  return 0;
}

// dW[C, ldw] (fp32, patch order) += dY[nimg*Ho*Wo, C]^T * patches(x)
int stem_wgrad_tc(const uint8_t *x, int nimg, int Cin, int xh, int xw, int Ho, int Wo, int C, const void *dY, float *dW, int ldw, cudaStream_t st) {
  StemArgs a;
  fill_args(&a, x, nimg, Cin, xh, xw, Ho, Wo, C);
  a.y = nullptr; a.dW = dW; a.ldw = ldw;
  size_t smem = 0;
  LEOD_REQUIRE(wgrad_plan(&a, &smem), "stem_wgrad_tc: shared memory / TMEM (Cin %d)", Cin);
  CUtensorMap mDY;
  LEOD_TRY(tc_make_map_2d(&mDY, dY, C, (uint64_t)nimg * Ho * Wo, C, 64, 16));
  static bool attr_set = false;
  if (!attr_set) {
    LEOD_CUDA(cudaFuncSetAttribute(stem_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
    attr_set = true;
  }
  const int per_half = std::max(1, std::min(a.num_tiles, num_sms_() / 2));
  LEOD_LAUNCH((stem_wgrad_kernel), dim3(per_half, 2), WG_THREADS, smem, st, mDY, a);
  LEOD_LAUNCH_CHECK();
  return 0;
}
