// Building blocks of the YOLOX PAFPN neck and decoupled head (models/detection/yolox/models/network_blocks.py:29-142,
// models/detection/yolox_extension/models/yolo_pafpn.py:109-140, models/detection/yolox/models/yolo_head.py:195-222).
//
// Layout ("padded flat"): a feature map [B, h, w, C] is stored as a row-major matrix with (h+2)*(w+2) rows per image —
// pixel (y, x) lives in row (y+1)*(w+2) + (x+1) and the one-pixel border rows are ZERO.  With that border a k=3, s=1,
// p=1 convolution is a sum over nine taps of the SAME matrix shifted by a constant number of rows,
//     out[r] = sum_tap  in[r + (ky-1)*(w+2) + (kx-1)] * W_tap^T ,
// i.e. an implicit GEMM whose A tiles are plain 2D TMA boxes at a shifted row coordinate (kernels_gemm_tc.cu,
// ConvTaps) — no patch matrix.  Rows shifted outside the matrix are zero-filled by TMA.  Border rows of a conv output
// hold garbage; every consumer below reads interior rows only and BN-apply re-establishes the zero border.
//
// BatchNorm runs on batch statistics in training (sum / sum of squares / count per channel, accumulated in double so
// that E[x^2] - mean^2 keeps fp32 accuracy) and on the running statistics in eval.
#include "common.cuh"

namespace {

struct V8 {
  float v[8];
};
template <typename T> __device__ __forceinline__ V8 load8(const T *p);
template <> __device__ __forceinline__ V8 load8<float>(const float *p) {
  V8 r;
  const float4 a = reinterpret_cast<const float4 *>(p)[0], b = reinterpret_cast<const float4 *>(p)[1];
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
template <> __device__ __forceinline__ V8 load8<bf16>(const bf16 *p) {
  V8 r;
  const uint4 u = *reinterpret_cast<const uint4 *>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    r.v[2 * j] = __uint_as_float(w[j] << 16);
    r.v[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
  }
  return r;
}
template <typename T> __device__ __forceinline__ void store8(T *p, const V8 &r);
template <> __device__ __forceinline__ void store8<float>(float *p, const V8 &r) {
  reinterpret_cast<float4 *>(p)[0] = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  reinterpret_cast<float4 *>(p)[1] = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
template <> __device__ __forceinline__ void store8<bf16>(bf16 *p, const V8 &r) {
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 pr = __floats2bfloat162_rn(r.v[2 * j], r.v[2 * j + 1]);
    w[j] = *reinterpret_cast<const uint32_t *>(&pr);
  }
  *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ V8 zero8() {
  V8 r;
#pragma unroll
  for (int j = 0; j < 8; ++j) r.v[j] = 0.f;
  return r;
}

// row inside one padded image -> pixel; false on the border
__device__ __forceinline__ bool pad_pixel(int q, const PadGeom &g, int &y, int &x) {
  y = q / g.w2 - 1;
  x = q % g.w2 - 1;
  return y >= 0 && y < g.h && x >= 0 && x < g.w;
}

// ------------------------------------------------------------------ layout changes
template <typename T>
__global__ void pad_gather_kernel(const T *__restrict__ src, T *__restrict__ dst, int ldd, PadGeom g, int C) {
  pdl_prologue();
  const int nc = C >> 3;
  const int64_t total = g.R * nc;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / nc;
    const int c = (int)(i % nc) * 8;
    const int b = (int)(r / g.P), q = (int)(r % g.P);
    int y, x;
    V8 v = zero8();
    if (pad_pixel(q, g, y, x)) v = load8<T>(src + (((int64_t)b * g.h + y) * g.w + x) * C + c);
    store8<T>(dst + r * ldd + c, v);
  }
}
// interior rows of a padded matrix -> dense NHWC
template <typename T>
__global__ void pad_scatter_kernel(const T *__restrict__ src, int lds, T *__restrict__ dst, PadGeom g, int C) {
  pdl_prologue();
  const int nc = C >> 3;
  const int64_t total = (int64_t)g.B * g.h * g.w * nc;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / nc;
    const int c = (int)(i % nc) * 8;
    const int x = (int)(p % g.w), y = (int)((p / g.w) % g.h), b = (int)(p / ((int64_t)g.w * g.h));
    const int64_t r = (int64_t)b * g.P + (y + 1) * g.w2 + (x + 1);
    store8<T>(dst + p * C + c, load8<T>(src + r * lds + c));
  }
}
// nearest-exact x2 (F.interpolate(scale_factor=2, mode='nearest-exact'), yolo_pafpn.py:49): dst(y, x) = src(y/2, x/2)
template <typename T>
__global__ void upsample2x_kernel(const T *__restrict__ src, int lds, PadGeom gs, T *__restrict__ dst, int ldd, PadGeom gd, int C) {
  pdl_prologue();
  const int nc = C >> 3;
  const int64_t total = gd.R * nc;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / nc;
    const int c = (int)(i % nc) * 8;
    const int b = (int)(r / gd.P), q = (int)(r % gd.P);
    int y, x;
    V8 v = zero8();
    if (pad_pixel(q, gd, y, x)) v = load8<T>(src + ((int64_t)b * gs.P + (y / 2 + 1) * gs.w2 + (x / 2 + 1)) * lds + c);
    store8<T>(dst + r * ldd + c, v);
  }
}
template <typename T>
__global__ void upsample2x_bwd_kernel(const T *__restrict__ ddst, int ldd, PadGeom gd, T *__restrict__ dsrc, int lds, PadGeom gs, int C,
                                      int accumulate) {
  pdl_prologue();
  const int nc = C >> 3;
  const int64_t total = gs.R * nc;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / nc;
    const int c = (int)(i % nc) * 8;
    const int b = (int)(r / gs.P), q = (int)(r % gs.P);
    int y, x;
    if (!pad_pixel(q, gs, y, x)) continue;   // border rows of a gradient matrix are never read
    V8 s = accumulate ? load8<T>(dsrc + r * lds + c) : zero8();
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const V8 v = load8<T>(ddst + ((int64_t)b * gd.P + (2 * y + dy + 1) * gd.w2 + (2 * x + dx + 1)) * ldd + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) s.v[j] += v.v[j];
      }
    store8<T>(dsrc + r * lds + c, s);
  }
}

// stride-2 3x3 patch matrix in the OUTPUT level's padded layout: col[r_out, tap*cinp + c] (border rows and channel pads zero)
template <typename T>
__global__ void im2col_pad_s2_kernel(const T *__restrict__ src, int lds, PadGeom gs, T *__restrict__ col, int ldcol, PadGeom go, int cin,
                                     int cinp) {
  pdl_prologue();
  const int nc = cinp >> 3;
  const int64_t total = go.R * 9 * nc;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % nc) * 8;
    const int tap = (int)((i / nc) % 9);
    const int64_t r = i / (9 * nc);
    const int b = (int)(r / go.P), q = (int)(r % go.P);
    int oy, ox;
    V8 v = zero8();
    if (c < cin && pad_pixel(q, go, oy, ox)) {
      const int iy = 2 * oy + tap / 3 - 1, ix = 2 * ox + tap % 3 - 1;   // in [-1, h]: always inside the padded source
      v = load8<T>(src + ((int64_t)b * gs.P + (iy + 1) * gs.w2 + (ix + 1)) * lds + c);
    }
    store8<T>(col + r * ldcol + tap * cinp + c, v);
  }
}
template <typename T>
__global__ void col2im_pad_s2_kernel(const T *__restrict__ dcol, int ldcol, PadGeom go, T *__restrict__ dsrc, int lds, PadGeom gs, int cin,
                                     int cinp, int accumulate) {
  pdl_prologue();
  const int nc = cin >> 3;
  const int64_t total = gs.R * nc;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / nc;
    const int c = (int)(i % nc) * 8;
    const int b = (int)(r / gs.P), q = (int)(r % gs.P);
    int iy, ix;
    if (!pad_pixel(q, gs, iy, ix)) continue;
    V8 s = accumulate ? load8<T>(dsrc + r * lds + c) : zero8();
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int oy2 = iy + 1 - ky;
      if (oy2 < 0 || (oy2 & 1) || (oy2 >> 1) >= go.h) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ox2 = ix + 1 - kx;
        if (ox2 < 0 || (ox2 & 1) || (ox2 >> 1) >= go.w) continue;
        const V8 v = load8<T>(dcol + ((int64_t)b * go.P + ((oy2 >> 1) + 1) * go.w2 + (ox2 >> 1) + 1) * ldcol + (ky * 3 + kx) * cinp + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) s.v[j] += v.v[j];
      }
    }
    store8<T>(dsrc + r * lds + c, s);
  }
}

// ------------------------------------------------------------------ BatchNorm (+SiLU)
// Per convolution: `stats` (double) = [0, C) sum, [C, 2C) sum of squares, [2C] element count, [2C+1] block ticket;
// `fin` (float) = [0, C) mean, [C, 2C) 1/sqrt(var + eps), written once per forward by the LAST block of the statistics
// kernel (or by bn_finalize_kernel after the data-parallel exchange), so the apply kernels never touch fp64.
// Backward likewise: `dstat` (double) = [0, C) sum dyhat, [C, 2C) sum dyhat*xhat, [2C] ticket; `dfin` (float) = the two means.
constexpr int BN_THREADS = 256;

struct BnFin {       // everything the finalisation needs
  double *stats;     // forward sums (+ count) — global in data-parallel mode
  float *fin;
  BnSeg s0, s1;
  int cseg;
  float eps, momentum;
  int update_running;
};

__device__ __forceinline__ void bn_finalize_channels(const BnFin &f, int C, int tid, int nthreads) {
  const double n = f.stats[2 * C];
  const double inv = 1.0 / n;
  for (int c = tid; c < C; c += nthreads) {
    const double m = f.stats[c] * inv;
    const double v = fmax(f.stats[C + c] * inv - m * m, 0.0);
    f.fin[c] = (float)m;
    f.fin[C + c] = rsqrtf((float)v + f.eps);
    if (f.update_running) {   // torch.nn.BatchNorm2d: momentum 0.1, unbiased variance (network_blocks.py:46)
      const BnSeg &sg = c < f.cseg ? f.s0 : f.s1;
      const int cl = c < f.cseg ? c : c - f.cseg;
      if (sg.rmean) {
        sg.rmean[cl] = (1.f - f.momentum) * sg.rmean[cl] + f.momentum * (float)m;
        sg.rvar[cl] = (1.f - f.momentum) * sg.rvar[cl] + f.momentum * (float)(v * (n / fmax(n - 1.0, 1.0)));
      }
      if (cl == 0 && sg.nbt) sg.nbt[0] += 1;
    }
  }
}
// true in every thread of the last block to arrive (all other blocks' atomics are then visible)
__device__ __forceinline__ bool last_block(double *ticket_slot) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long t = atomicAdd(reinterpret_cast<unsigned long long *>(ticket_slot), 1ULL);
    s_last = (t == (unsigned long long)gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

template <typename T>
__global__ void __launch_bounds__(BN_THREADS) bn_stats_kernel(const T *__restrict__ Y, int ldy, PadGeom g, int C, BnFin f, int finalize,
                                                              int rows_per_block) {
  pdl_prologue();
  __shared__ float red[BN_THREADS][17];
  double *stats = f.stats;
  const int nc = C >> 3;
  const int tid = threadIdx.x;
  const int rl = tid / nc, chunk = tid - rl * nc;
  const int lanes = blockDim.x / nc;
  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(g.R, r0 + rows_per_block);
  for (int64_t rb = r0 + rl; rb < r1; rb += 4 * lanes) {   // four rows in flight per thread
    V8 v[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t r = rb + (int64_t)u * lanes;
      int y, x;
      ok[u] = r < r1 && pad_pixel((int)(r % g.P), g, y, x);
      if (ok[u]) v[u] = load8<T>(Y + r * ldy + chunk * 8);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!ok[u]) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += v[u].v[j];
        ss[j] = fmaf(v[u].v[j], v[u].v[j], ss[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[tid][j] = s[j];
    red[tid][8 + j] = ss[j];
  }
  __syncthreads();
  for (int i = tid; i < nc * 16; i += blockDim.x) {
    const int c = i >> 4, j = i & 15;
    double t = 0.0;
    for (int l = 0; l < lanes; ++l) t += (double)red[l * nc + c][j];
    atomicAdd(&stats[(j >> 3) * C + c * 8 + (j & 7)], t);
  }
  if (blockIdx.x == 0 && tid == 0) atomicAdd(&stats[2 * C], (double)g.B * g.h * g.w);
  if (finalize && last_block(&stats[2 * C + 1])) bn_finalize_channels(f, C, tid, blockDim.x);
}
__global__ void __launch_bounds__(BN_THREADS) bn_finalize_kernel(BnFin f, int C) {
  pdl_prologue(); bn_finalize_channels(f, C, threadIdx.x, blockDim.x); }

// 8 consecutive floats as two 16-byte loads
__device__ __forceinline__ V8 ldg8(const float *p) {
  V8 r;
  const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}

// z = silu(gamma * (y - mean) * rstd + beta) on interior rows, 0 on the border.  training: batch statistics from `fin`;
// eval: the running statistics.  Two parameter segments serve the fused twin convolutions.  One (row, 8-channel chunk) per
// thread: these matrices are a few MB, so the kernel is a latency chain and wants every load in flight at once.
template <typename T>
__global__ void __launch_bounds__(BN_THREADS) bn_apply_silu_kernel(const T *__restrict__ Y, int ldy, PadGeom g, int C, const float *__restrict__ fin,
                                                                   BnSeg s0, BnSeg s1, int cseg, T *__restrict__ Z, int ldz, float eps, int training) {
  pdl_prologue();
  const int nc = C >> 3;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.R * nc) return;
  const int64_t r = i / nc;
  const int c0 = (int)(i - r * nc) * 8;
  int y, x;
  V8 o = zero8();
  if (pad_pixel((int)(r % g.P), g, y, x)) {
    const V8 v = load8<T>(Y + r * ldy + c0);
    const BnSeg &sg = c0 < cseg ? s0 : s1;
    const int cl = c0 < cseg ? c0 : c0 - cseg;
    const V8 ga = ldg8(sg.gamma + cl), be = ldg8(sg.beta + cl);
    const V8 mean = training ? ldg8(fin + c0) : ldg8(sg.rmean + cl);
    V8 rstd;
    if (training) {
      rstd = ldg8(fin + C + c0);
    } else {
      rstd = ldg8(sg.rvar + cl);
#pragma unroll
      for (int j = 0; j < 8; ++j) rstd.v[j] = rsqrtf(rstd.v[j] + eps);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sc = ga.v[j] * rstd.v[j];
      const float u = fmaf(v.v[j] - mean.v[j], sc, be.v[j]);
      o.v[j] = u / (1.f + __expf(-u));
    }
  }
  store8<T>(Z + r * ldz + c0, o);
}

struct BnBwdFin {
  double *dloc, *dglob;   // local sums / sums over all ranks (the same array in a single process)
  const double *stats;    // forward statistics (for the element count)
  float *dfin;
  BnSeg s0, s1;
  int cseg;
};
// means over the (global) batch for the input gradient; dgamma += LOCAL sum dyhat*xhat, dbeta += LOCAL sum dyhat (SyncBatchNorm
// semantics: the parameter gradients stay local, the data-parallel average happens in the flat-gradient all-reduce)
__device__ __forceinline__ void bn_bwd_finalize_channels(const BnBwdFin &f, int C, int tid, int nthreads) {
  const double inv = 1.0 / f.stats[2 * C];
  for (int c = tid; c < C; c += nthreads) {
    f.dfin[c] = (float)(f.dglob[c] * inv);
    f.dfin[C + c] = (float)(f.dglob[C + c] * inv);
    const BnSeg &sg = c < f.cseg ? f.s0 : f.s1;
    const int cl = c < f.cseg ? c : c - f.cseg;
    if (sg.dgamma) {
      sg.dgamma[cl] += (float)f.dloc[C + c];
      sg.dbeta[cl] += (float)f.dloc[c];
    }
  }
}

// dyhat = dz * silu'(u);  dstat[0,C) += sum dyhat,  dstat[C,2C) += sum dyhat * xhat
template <typename T>
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_reduce_kernel(const T *__restrict__ dZ, int lddz, const T *__restrict__ Y, int ldy, PadGeom g, int C,
                                                                   const float *__restrict__ fin, BnBwdFin f, int finalize, int rows_per_block) {
  pdl_prologue();
  __shared__ float red[BN_THREADS][17];
  const int nc = C >> 3;
  const int tid = threadIdx.x;
  const int rl = tid / nc, chunk = tid - rl * nc;
  const int lanes = blockDim.x / nc;
  float a1[8], a2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a1[j] = a2[j] = 0.f;
  {
    const int c0 = chunk * 8;
    const BnSeg &sg = c0 < f.cseg ? f.s0 : f.s1;
    const int cl = c0 < f.cseg ? c0 : c0 - f.cseg;
    const V8 mean = ldg8(fin + c0), rstd = ldg8(fin + C + c0), ga = ldg8(sg.gamma + cl), be = ldg8(sg.beta + cl);
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = min(g.R, r0 + rows_per_block);
    for (int64_t rb = r0 + rl; rb < r1; rb += 2 * lanes) {   // two rows (four loads) in flight per thread
      V8 yv[2], dz[2];
      bool ok[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t r = rb + (int64_t)u * lanes;
        int y, x;
        ok[u] = r < r1 && pad_pixel((int)(r % g.P), g, y, x);
        if (ok[u]) {
          yv[u] = load8<T>(Y + r * ldy + c0);
          dz[u] = load8<T>(dZ + r * lddz + c0);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (!ok[u]) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (yv[u].v[j] - mean.v[j]) * rstd.v[j];
          const float uu = fmaf(ga.v[j], xh, be.v[j]);
          const float sg_ = 1.f / (1.f + __expf(-uu));
          const float dyh = dz[u].v[j] * sg_ * (1.f + uu * (1.f - sg_));
          a1[j] += dyh;
          a2[j] = fmaf(dyh, xh, a2[j]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[tid][j] = a1[j];
    red[tid][8 + j] = a2[j];
  }
  __syncthreads();
  for (int i = tid; i < nc * 16; i += blockDim.x) {
    const int c = i >> 4, j = i & 15;
    double t = 0.0;
    for (int l = 0; l < lanes; ++l) t += (double)red[l * nc + c][j];
    atomicAdd(&f.dloc[(j >> 3) * C + c * 8 + (j & 7)], t);
  }
  if (finalize && last_block(&f.dloc[2 * C])) bn_bwd_finalize_channels(f, C, tid, blockDim.x);
}
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_finalize_kernel(BnBwdFin f, int C) {
  pdl_prologue(); bn_bwd_finalize_channels(f, C, threadIdx.x, blockDim.x); }

// dy = gamma * rstd * (dyhat - mean(dyhat) - xhat * mean(dyhat * xhat)), 0 on the border
template <typename T>
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_apply_kernel(const T *__restrict__ dZ, int lddz, const T *__restrict__ Y, int ldy, PadGeom g, int C,
                                                                  const float *__restrict__ fin, const float *__restrict__ dfin, BnSeg s0, BnSeg s1,
                                                                  int cseg, T *__restrict__ dY, int lddy) {
  pdl_prologue();
  const int nc = C >> 3;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.R * nc) return;
  const int64_t r = i / nc;
  const int c0 = (int)(i - r * nc) * 8;
  int y, x;
  V8 o = zero8();
  if (pad_pixel((int)(r % g.P), g, y, x)) {
    const V8 yv = load8<T>(Y + r * ldy + c0);
    const V8 dz = load8<T>(dZ + r * lddz + c0);
    const BnSeg &sg = c0 < cseg ? s0 : s1;
    const int cl = c0 < cseg ? c0 : c0 - cseg;
    const V8 ga = ldg8(sg.gamma + cl), be = ldg8(sg.beta + cl), mean = ldg8(fin + c0), rstd = ldg8(fin + C + c0), m1 = ldg8(dfin + c0),
             m2 = ldg8(dfin + C + c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (yv.v[j] - mean.v[j]) * rstd.v[j];
      const float u = fmaf(ga.v[j], xh, be.v[j]);
      const float sg_ = 1.f / (1.f + __expf(-u));
      const float dyh = dz.v[j] * sg_ * (1.f + u * (1.f - sg_));
      o.v[j] = ga.v[j] * rstd.v[j] * (dyh - m1.v[j] - xh * m2.v[j]);
    }
  }
  store8<T>(dY + r * lddy + c0, o);
}

template <typename T> __global__ void fill_zero_kernel(T *p, int64_t n) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = T(0);
}

inline int grid_for(int64_t items, int threads = 256) {
  int64_t b = (items + threads - 1) / threads;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (int)b;
}

// BN kernels: threads = largest multiple of (C/8) <= 256; rows per block chosen so that ~2 waves of CTAs cover the matrix
inline void bn_launch_shape(int C, int64_t R, int *threads, int *rows_per_block, int *blocks) {
  const int nc = C >> 3;
  const int lanes = BN_THREADS / nc > 0 ? BN_THREADS / nc : 1;
  *threads = lanes * nc;
  int64_t rpb = (R + 148 * 2 - 1) / (148 * 2);
  rpb = (rpb + lanes - 1) / lanes * lanes;
  if (rpb < lanes) rpb = lanes;
  *rows_per_block = (int)rpb;
  *blocks = (int)((R + rpb - 1) / rpb);
}

}  // namespace

#define DISPATCH_T(dtype, ...)                 \
  do {                                         \
    if ((dtype) == LEOD_F32) {                 \
      typedef float T;                         \
      __VA_ARGS__;                             \
    } else {                                   \
      typedef bf16 T;                          \
      __VA_ARGS__;                             \
    }                                          \
  } while (0)

PadGeom make_pad_geom(int B, int h, int w) {
  PadGeom g;
  g.B = B; g.h = h; g.w = w; g.w2 = w + 2; g.P = (h + 2) * (w + 2); g.R = (int64_t)B * g.P;
  return g;
}

int pad_gather(int dtype, const void *src, void *dst, int ldd, const PadGeom &g, int C, cudaStream_t st) {
  LEOD_REQUIRE(C % 8 == 0 && ldd % 8 == 0, "pad_gather: C %d / ld %d must be multiples of 8", C, ldd);
  DISPATCH_T(dtype, (LEOD_LAUNCH((pad_gather_kernel<T>), grid_for(g.R * (C / 8)), 256, 0, st, (const T *)src, (T *)dst, ldd, g, C)));
  LEOD_LAUNCH_CHECK();
  return 0;
}
int pad_scatter(int dtype, const void *src, int lds, void *dst, const PadGeom &g, int C, cudaStream_t st) {
  LEOD_REQUIRE(C % 8 == 0 && lds % 8 == 0, "pad_scatter: C %d / ld %d must be multiples of 8", C, lds);
  DISPATCH_T(dtype, (LEOD_LAUNCH((pad_scatter_kernel<T>), grid_for((int64_t)g.B * g.h * g.w * (C / 8)), 256, 0, st, (const T *)src, lds, (T *)dst, g, C)));
  LEOD_LAUNCH_CHECK();
  return 0;
}
int upsample2x(int dtype, const void *src, int lds, const PadGeom &gs, void *dst, int ldd, const PadGeom &gd, int C, cudaStream_t st) {
  LEOD_REQUIRE(gd.h == 2 * gs.h && gd.w == 2 * gs.w && gd.B == gs.B && C % 8 == 0, "upsample2x: shape mismatch");
  DISPATCH_T(dtype, (LEOD_LAUNCH((upsample2x_kernel<T>), grid_for(gd.R * (C / 8)), 256, 0, st, (const T *)src, lds, gs, (T *)dst, ldd, gd, C)));
  LEOD_LAUNCH_CHECK();
  return 0;
}
int upsample2x_bwd(int dtype, const void *ddst, int ldd, const PadGeom &gd, void *dsrc, int lds, const PadGeom &gs, int C, int accumulate,
                   cudaStream_t st) {
  LEOD_REQUIRE(gd.h == 2 * gs.h && gd.w == 2 * gs.w && gd.B == gs.B && C % 8 == 0, "upsample2x_bwd: shape mismatch");
  DISPATCH_T(dtype, (LEOD_LAUNCH((upsample2x_bwd_kernel<T>), grid_for(gs.R * (C / 8)), 256, 0, st, (const T *)ddst, ldd, gd, (T *)dsrc, lds, gs, C, accumulate)));
  LEOD_LAUNCH_CHECK();
  return 0;
}
int im2col_pad_s2(int dtype, const void *src, int lds, const PadGeom &gs, void *col, int ldcol, const PadGeom &go, int cin, int cinp,
                  cudaStream_t st) {
  LEOD_REQUIRE(gs.h == 2 * go.h && gs.w == 2 * go.w && cin % 8 == 0 && cinp % 8 == 0, "im2col_pad_s2: shape mismatch");
  DISPATCH_T(dtype, (LEOD_LAUNCH((im2col_pad_s2_kernel<T>), grid_for(go.R * 9 * (cinp / 8)), 256, 0, st, (const T *)src, lds, gs, (T *)col, ldcol, go, cin, cinp)));
  LEOD_LAUNCH_CHECK();
  return 0;
}
int col2im_pad_s2(int dtype, const void *dcol, int ldcol, const PadGeom &go, void *dsrc, int lds, const PadGeom &gs, int cin, int cinp,
                  int accumulate, cudaStream_t st) {
  LEOD_REQUIRE(gs.h == 2 * go.h && gs.w == 2 * go.w && cin % 8 == 0, "col2im_pad_s2: shape mismatch");
  DISPATCH_T(dtype, (LEOD_LAUNCH((col2im_pad_s2_kernel<T>), grid_for(gs.R * (cin / 8)), 256, 0, st, (const T *)dcol, ldcol, go, (T *)dsrc, lds, gs, cin, cinp, accumulate)));
  LEOD_LAUNCH_CHECK();
  return 0;
}

// finalize != 0: the last block also turns the sums into mean / rstd (single process); 0: bn_finalize() does it after the exchange
int bn_stats(int dtype, const void *Y, int ldy, const PadGeom &g, const BnLayer &l, int finalize, cudaStream_t st) {
  LEOD_REQUIRE(l.C % 8 == 0 && l.cseg % 8 == 0 && l.C <= 8 * BN_THREADS, "bn_stats: C = %d", l.C);
  int threads, rpb, blocks;
  bn_launch_shape(l.C, g.R, &threads, &rpb, &blocks);
  BnFin f{l.stats, l.fin, l.s0, l.s1, l.cseg, l.eps, l.momentum, 1};
  DISPATCH_T(dtype, (LEOD_LAUNCH((bn_stats_kernel<T>), blocks, threads, 0, st, (const T *)Y, ldy, g, l.C, f, finalize, rpb)));
  LEOD_LAUNCH_CHECK();
  return 0;
}
int bn_finalize(const BnLayer &l, cudaStream_t st) {
  BnFin f{l.stats, l.fin, l.s0, l.s1, l.cseg, l.eps, l.momentum, 1};
  LEOD_LAUNCH((bn_finalize_kernel), 1, BN_THREADS, 0, st, f, l.C);
  LEOD_LAUNCH_CHECK();
  return 0;
}
int bn_apply_silu(int dtype, const void *Y, int ldy, const PadGeom &g, const BnLayer &l, void *Z, int ldz, int training, cudaStream_t st) {
  const int blocks = ceil_div(g.R * (l.C / 8), BN_THREADS);
  DISPATCH_T(dtype, (LEOD_LAUNCH((bn_apply_silu_kernel<T>), blocks, BN_THREADS, 0, st, (const T *)Y, ldy, g, l.C, l.fin, l.s0, l.s1, l.cseg, (T *)Z, ldz, l.eps,
                                                                            training)));
  LEOD_LAUNCH_CHECK();
  return 0;
}
int bn_bwd_reduce(int dtype, const void *dZ, int lddz, const void *Y, int ldy, const PadGeom &g, const BnLayer &l, int finalize, cudaStream_t st) {
  int threads, rpb, blocks;
  bn_launch_shape(l.C, g.R, &threads, &rpb, &blocks);
  BnBwdFin f{l.dloc, l.dglob, l.stats, l.dfin, l.s0, l.s1, l.cseg};
  DISPATCH_T(dtype, (LEOD_LAUNCH((bn_bwd_reduce_kernel<T>), blocks, threads, 0, st, (const T *)dZ, lddz, (const T *)Y, ldy, g, l.C, l.fin, f, finalize, rpb)));
  LEOD_LAUNCH_CHECK();
  return 0;
}
int bn_bwd_finalize(const BnLayer &l, cudaStream_t st) {
  BnBwdFin f{l.dloc, l.dglob, l.stats, l.dfin, l.s0, l.s1, l.cseg};
  LEOD_LAUNCH((bn_bwd_finalize_kernel), 1, BN_THREADS, 0, st, f, l.C);
  LEOD_LAUNCH_CHECK();
  return 0;
}
int bn_bwd_apply(int dtype, const void *dZ, int lddz, const void *Y, int ldy, const PadGeom &g, const BnLayer &l, void *dY, int lddy, cudaStream_t st) {
  const int blocks = ceil_div(g.R * (l.C / 8), BN_THREADS);
  DISPATCH_T(dtype, (LEOD_LAUNCH((bn_bwd_apply_kernel<T>), blocks, BN_THREADS, 0, st, (const T *)dZ, lddz, (const T *)Y, ldy, g, l.C, l.fin, l.dfin, l.s0, l.s1,
                                                                           l.cseg, (T *)dY, lddy)));
  LEOD_LAUNCH_CHECK();
  return 0;
}

int device_zero_bytes(void *p, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return 0;
  LEOD_REQUIRE(bytes % 4 == 0, "device_zero_bytes: %zu not a multiple of 4", bytes);
  LEOD_LAUNCH((fill_zero_kernel<uint32_t>), grid_for((int64_t)(bytes / 4)), 256, 0, st, (uint32_t *)p, (int64_t)(bytes / 4));
  LEOD_LAUNCH_CHECK();
  return 0;
}
