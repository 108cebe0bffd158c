// HBM-bound kernels of the backbone: LayerNorm, ConvLSTM gate math, patch gather/scatter for the
// strided convolutions, weight preparation.  All are templated on the activation storage type and
// compute in fp32.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int LN_MAX_C = 512;

// ------------------------------------------------------------------ LayerNorm (one warp per token)
template <typename T, int LN_MAXV>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const T *__restrict__ x, const float *__restrict__ w,
                                                     const float *__restrict__ b, T *__restrict__ y, int M, int C, float eps) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const T *xr = x + (size_t)row * C;
  float v[LN_MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int c = lane + 32 * i;
    v[i] = (c < C) ? to_f<T>(xr[c]) : 0.f;
    s += v[i];
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int c = lane + 32 * i;
    const float d = (c < C) ? v[i] - mean : 0.f;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  T *yr = y + (size_t)row * C;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < C) yr[c] = from_f<T>((v[i] - mean) * rstd * w[c] + b[c]);
  }
}

template <typename T, int LN_MAXV>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const T *__restrict__ x, const float *__restrict__ w,
                                                     const T *__restrict__ dy, const T *__restrict__ dres, T *__restrict__ dx,
                                                     float *__restrict__ dw, float *__restrict__ db, int M, int C, float eps) {
  pdl_prologue();
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  float dwacc[LN_MAXV], dbacc[LN_MAXV];
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) dwacc[i] = dbacc[i] = 0.f;
  for (int row = blockIdx.x * nwarp + warp; row < M; row += gridDim.x * nwarp) {
    const T *xr = x + (size_t)row * C;
    const T *gr = dy + (size_t)row * C;
    float v[LN_MAXV], g[LN_MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = lane + 32 * i;
      v[i] = (c < C) ? to_f<T>(xr[c]) : 0.f;
      g[i] = (c < C) ? to_f<T>(gr[c]) : 0.f;
      s += v[i];
    }
    const float mean = warp_sum(s) / C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = lane + 32 * i;
      const float d = (c < C) ? v[i] - mean : 0.f;
      q += d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) / C + eps);
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < C) {
        const float xh = (v[i] - mean) * rstd;
        dwacc[i] += g[i] * xh;
        dbacc[i] += g[i];
        const float gw = g[i] * w[c];
        v[i] = xh;   // keep xhat
        g[i] = gw;   // keep dy*w
        c1 += gw;
        c2 += gw * xh;
      }
    }
    c1 = warp_sum(c1) / C;
    c2 = warp_sum(c2) / C;
    T *dxr = dx + (size_t)row * C;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < C) {
        float o = rstd * (g[i] - c1 - v[i] * c2);
        if (dres) o += to_f<T>(dres[(size_t)row * C + c]);
        dxr[c] = from_f<T>(o);
      }
    }
  }
  // cross-warp reduction of the affine gradients, one channel slot at a time
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    if (32 * i >= C) break;
    for (int pass = 0; pass < 2; ++pass) {
      red[warp][lane] = pass == 0 ? dwacc[i] : dbacc[i];
      __syncthreads();
      if (warp == 0) {
        float t = 0.f;
        for (int k = 0; k < nwarp; ++k) t += red[k][lane];
        const int c = lane + 32 * i;
        if (c < C) atomicAdd(pass == 0 ? &dw[c] : &db[c], t);
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------ LayerNorm, bf16 fast path (C % 8 == 0)
// LPR lanes share a token; every lane owns NCH chunks of 8 channels (one 16-byte load each), so a warp reads
// 32/LPR whole rows with full-width accesses instead of 2-byte strided ones.
__device__ __forceinline__ void bf16x8_to_f(const uint4 &u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(w[j] << 16);
    f[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 f_to_bf16x8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    w[j] = *reinterpret_cast<const uint32_t *>(&p);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int LPR, int NCH>
__global__ void __launch_bounds__(256) ln_fwd_bf16_kernel(const bf16 *__restrict__ x, const float *__restrict__ w,
                                                          const float *__restrict__ b, bf16 *__restrict__ y, int M, int C, float eps) {
  pdl_prologue();
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane % LPR;
  const int row = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LPR;
  const int nchunk = C >> 3;
  const bool row_ok = row < M;
  float v[NCH][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int ch = sub + LPR * i;
    if (row_ok && ch < nchunk) {
      bf16x8_to_f(*reinterpret_cast<const uint4 *>(x + (size_t)row * C + ch * 8), v[i]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[i][j];
  }
  const float mean = group_sum<LPR>(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    if (sub + LPR * i < nchunk) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        q += d * d;
      }
    }
  }
  const float rstd = rsqrtf(group_sum<LPR>(q) / C + eps);
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int ch = sub + LPR * i;
    if (row_ok && ch < nchunk) {
      float wv[8], bv[8], o[8];
      *reinterpret_cast<float4 *>(wv) = __ldg(reinterpret_cast<const float4 *>(w + ch * 8));
      *reinterpret_cast<float4 *>(wv + 4) = __ldg(reinterpret_cast<const float4 *>(w + ch * 8 + 4));
      *reinterpret_cast<float4 *>(bv) = __ldg(reinterpret_cast<const float4 *>(b + ch * 8));
      *reinterpret_cast<float4 *>(bv + 4) = __ldg(reinterpret_cast<const float4 *>(b + ch * 8 + 4));
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * wv[j] + bv[j];
      *reinterpret_cast<uint4 *>(y + (size_t)row * C + ch * 8) = f_to_bf16x8(o);
    }
  }
}

template <int LPR, int NCH>
__global__ void __launch_bounds__(256) ln_bwd_bf16_kernel(const bf16 *__restrict__ x, const float *__restrict__ w,
                                                          const bf16 *__restrict__ dy, const bf16 *__restrict__ dres,
                                                          bf16 *__restrict__ dx, float *__restrict__ dw, float *__restrict__ db, int M,
                                                          int C, float eps) {
  pdl_prologue();
  constexpr int RPW = 32 / LPR;
  __shared__ float sdw[LN_MAX_C], sdb[LN_MAX_C];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sdw[i] = sdb[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, sub = lane % LPR;
  const int nchunk = C >> 3;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  float wv[NCH][8], dwacc[NCH][8], dbacc[NCH][8];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int ch = sub + LPR * i;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      wv[i][j] = ch < nchunk ? w[ch * 8 + j] : 0.f;
      dwacc[i][j] = dbacc[i][j] = 0.f;
    }
  }
  for (int row = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LPR;; row += warps_total * RPW) {
    // all lanes of a warp leave together: the first row of the warp decides
    if (row - lane / LPR >= M) break;
    const bool row_ok = row < M;
    float v[NCH][8], g[NCH][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int ch = sub + LPR * i;
      if (row_ok && ch < nchunk) {
        bf16x8_to_f(*reinterpret_cast<const uint4 *>(x + (size_t)row * C + ch * 8), v[i]);
        bf16x8_to_f(*reinterpret_cast<const uint4 *>(dy + (size_t)row * C + ch * 8), g[i]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = g[i][j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    }
    const float mean = group_sum<LPR>(s) / C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      if (sub + LPR * i < nchunk) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[i][j] - mean;
          q += d * d;
        }
      }
    }
    const float rstd = rsqrtf(group_sum<LPR>(q) / C + eps);
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      if (row_ok && sub + LPR * i < nchunk) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (v[i][j] - mean) * rstd;
          dwacc[i][j] += g[i][j] * xh;
          dbacc[i][j] += g[i][j];
          const float gw = g[i][j] * wv[i][j];
          v[i][j] = xh;
          g[i][j] = gw;
          c1 += gw;
          c2 += gw * xh;
        }
      }
    }
    c1 = group_sum<LPR>(c1) / C;
    c2 = group_sum<LPR>(c2) / C;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int ch = sub + LPR * i;
      if (row_ok && ch < nchunk) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (g[i][j] - c1 - v[i][j] * c2);
        if (dres) {
          float r[8];
          bf16x8_to_f(*reinterpret_cast<const uint4 *>(dres + (size_t)row * C + ch * 8), r);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += r[j];
        }
        *reinterpret_cast<uint4 *>(dx + (size_t)row * C + ch * 8) = f_to_bf16x8(o);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int ch = sub + LPR * i;
    if (ch < nchunk) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&sdw[ch * 8 + j], dwacc[i][j]);
        atomicAdd(&sdb[ch * 8 + j], dbacc[i][j]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(&dw[i], sdw[i]);
    atomicAdd(&db[i], sdb[i]);
  }
}

// ------------------------------------------------------------------ ConvLSTM gate math (rnn.py:57-68)
template <typename T>
__global__ void lstm_fwd_kernel(T *__restrict__ gates, const T *__restrict__ c_prev, T *__restrict__ h_out,
                                T *__restrict__ c_out, int M, int C) {
  pdl_prologue();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)M * C) return;
  const int m = (int)(idx / C), c = (int)(idx % C);
  T *gr = gates + (size_t)m * 4 * C;
  const float f = sigmoid_f(to_f<T>(gr[c]));
  const float i = sigmoid_f(to_f<T>(gr[C + c]));
  const float o = sigmoid_f(to_f<T>(gr[2 * C + c]));
  const float g = tanhf(to_f<T>(gr[3 * C + c]));
  const float cp = c_prev ? to_f<T>(c_prev[idx]) : 0.f;
  const float cn = f * cp + i * g;
  gr[c] = from_f<T>(f);
  gr[C + c] = from_f<T>(i);
  gr[2 * C + c] = from_f<T>(o);
  gr[3 * C + c] = from_f<T>(g);
  c_out[idx] = from_f<T>(cn);
  h_out[idx] = from_f<T>(o * tanhf(cn));
}

template <typename T>
__global__ void lstm_bwd_kernel(const T *__restrict__ gates, const T *__restrict__ c_prev, const T *__restrict__ c_out,
                                const T *__restrict__ dh, const T *__restrict__ dh2, const T *__restrict__ dh3,
                                const T *dc, T *__restrict__ dgates, T *dc_prev, int M, int C) {
  pdl_prologue();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)M * C) return;
  const int m = (int)(idx / C), c = (int)(idx % C);
  const T *gr = gates + (size_t)m * 4 * C;
  const float f = to_f<T>(gr[c]), i = to_f<T>(gr[C + c]), o = to_f<T>(gr[2 * C + c]), g = to_f<T>(gr[3 * C + c]);
  const float cp = c_prev ? to_f<T>(c_prev[idx]) : 0.f;
  const float tc = tanhf(to_f<T>(c_out[idx]));
  float dhv = dh ? to_f<T>(dh[idx]) : 0.f;
  if (dh2) dhv += to_f<T>(dh2[idx]);
  if (dh3) dhv += to_f<T>(dh3[idx]);
  float dcv = dc ? to_f<T>(dc[idx]) : 0.f;
  dcv += dhv * o * (1.f - tc * tc);
  T *dg = dgates + (size_t)m * 4 * C;
  dg[c] = from_f<T>(dcv * cp * f * (1.f - f));
  dg[C + c] = from_f<T>(dcv * g * i * (1.f - i));
  dg[2 * C + c] = from_f<T>(dhv * tc * o * (1.f - o));
  dg[3 * C + c] = from_f<T>(dcv * i * (1.f - g * g));
  dc_prev[idx] = from_f<T>(dcv * f);
}

template <typename T>
__global__ void add_kernel(const T *__restrict__ a, const T *__restrict__ b, T *__restrict__ out, int64_t n) {
  pdl_prologue();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) out[idx] = from_f<T>(to_f<T>(a[idx]) + to_f<T>(b[idx]));
}

// ------------------------------------------------------------------ patch gather for the strided convs
// stem: x [B,Cin,xh,xw] channel-first (u8 / f32 / bf16), implicit zero pad to (Hp,Wp) and conv padding.
// Patch layout: k = (cin*ksz + ky)*8 + slot, slot 0 <-> kx = -1 (its weight is zero), slots 1..7 <-> kx = 0..6, so one
// 16-byte chunk of a patch row is 8 consecutive input pixels starting at a 4-aligned column (stride 4, pad 3):
// two 32-bit loads for uint8 input instead of eight byte gathers.  One thread = one chunk.
template <typename TI, typename T>
__global__ void __launch_bounds__(256) im2col_nchw_kernel(const TI *__restrict__ x, T *__restrict__ col, int B, int Cin, int xh,
                                                          int xw, int Ho, int Wo, int ksz, int stride, int pad, int ldcol) {
  pdl_prologue();
  const int chunks = Cin * ksz;   // == ldcol / 8
  const int64_t total = (int64_t)B * Ho * Wo * chunks;
  const bool word_path = sizeof(TI) == 1 && (xw & 3) == 0 && ((stride * 1 - pad - 1) & 3) == 0 && (stride & 3) == 0 &&
                         ((uintptr_t)x & 3) == 0;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(idx % chunks);
    const int64_t m = idx / chunks;
    const int ox = (int)(m % Wo), oy = (int)((m / Wo) % Ho), b = (int)(m / ((int64_t)Wo * Ho));
    const int cin = ch / ksz, ky = ch - cin * ksz;
    const int iy = oy * stride - pad + ky, ix0 = ox * stride - pad - 1;
    __align__(16) T vals[8];
    if (iy < 0 || iy >= xh) {
#pragma unroll
      for (int j = 0; j < 8; ++j) vals[j] = from_f<T>(0.f);
    } else {
      const TI *row = x + ((size_t)b * Cin + cin) * xh * xw + (size_t)iy * xw;
      if (word_path) {
        uint32_t w0 = 0, w1 = 0;
        if (ix0 >= 0 && ix0 + 3 < xw) w0 = *reinterpret_cast<const uint32_t *>(row + ix0);
        if (ix0 + 4 >= 0 && ix0 + 7 < xw) w1 = *reinterpret_cast<const uint32_t *>(row + ix0 + 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          vals[j] = from_f<T>((float)((w0 >> (8 * j)) & 0xffu));
          vals[4 + j] = from_f<T>((float)((w1 >> (8 * j)) & 0xffu));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int ix = ix0 + j;
          vals[j] = from_f<T>((j > 0 && j <= ksz && ix >= 0 && ix < xw) ? to_f<TI>(row[ix]) : 0.f);
        }
      }
    }
    T *dst = col + (size_t)m * ldcol + ch * 8;
    if (sizeof(T) == 2) {
      *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(vals);
    } else {
      reinterpret_cast<uint4 *>(dst)[0] = reinterpret_cast<const uint4 *>(vals)[0];
      reinterpret_cast<uint4 *>(dst)[1] = reinterpret_cast<const uint4 *>(vals)[1];
    }
  }
}

// stages 2-4: x [B,Hi,Wi,Cin] channels-last; 8 channels (16 bytes in bf16) per thread
template <typename T>
__global__ void __launch_bounds__(256) im2col_nhwc_kernel(const T *__restrict__ x, T *__restrict__ col, int B, int Hi, int Wi,
                                                          int Cin, int Ho, int Wo, int ksz, int stride, int pad, int K, int ldcol) {
  pdl_prologue();
  const int cchunks = Cin / 8, chunks = ldcol / 8;
  const int64_t total = (int64_t)B * Ho * Wo * chunks;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(idx % chunks);
    const int64_t m = idx / chunks;
    T *dst = col + (size_t)m * ldcol + ch * 8;
    const int tap = ch / cchunks, c8 = ch % cchunks;
    bool ok = ch * 8 < K;
    const T *src = nullptr;
    if (ok) {
      const int kx = tap % ksz, ky = tap / ksz;
      const int ox = (int)(m % Wo), oy = (int)((m / Wo) % Ho), b = (int)(m / ((int64_t)Wo * Ho));
      const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
      ok = iy >= 0 && iy < Hi && ix >= 0 && ix < Wi;
      src = x + (((size_t)b * Hi + iy) * Wi + ix) * Cin + c8 * 8;
    }
    constexpr int NV = sizeof(T) * 8 / 16;
#pragma unroll
    for (int v = 0; v < NV; ++v)
      reinterpret_cast<uint4 *>(dst)[v] = ok ? reinterpret_cast<const uint4 *>(src)[v] : make_uint4(0, 0, 0, 0);
  }
}

// gather form of the transposed patch scatter: dx[b,iy,ix,c] = dres + sum over the (<= 4) output
// positions whose window covers (iy,ix)
template <typename T>
__global__ void __launch_bounds__(256) col2im_nhwc_kernel(const T *__restrict__ dcol, int ldcol, const T *__restrict__ dres,
                                                          T *__restrict__ dx, int B, int Hi, int Wi, int Cin, int Ho, int Wo, int ksz,
                                                          int stride, int pad) {
  pdl_prologue();
  const int cchunks = Cin / 8;
  const int64_t total = (int64_t)B * Hi * Wi * cchunks;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % cchunks);
    const int ix = (int)((idx / cchunks) % Wi), iy = (int)((idx / ((int64_t)cchunks * Wi)) % Hi),
              b = (int)(idx / ((int64_t)cchunks * Wi * Hi));
    float acc[8];
    const size_t o = (((size_t)b * Hi + iy) * Wi + ix) * Cin + c8 * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = dres ? to_f<T>(dres[o + j]) : 0.f;
    for (int ky = 0; ky < ksz; ++ky) {
      const int ty = iy + pad - ky;
      if (ty < 0 || ty % stride) continue;
      const int oy = ty / stride;
      if (oy >= Ho) continue;
      for (int kx = 0; kx < ksz; ++kx) {
        const int tx = ix + pad - kx;
        if (tx < 0 || tx % stride) continue;
        const int ox = tx / stride;
        if (ox >= Wo) continue;
        const T *src = dcol + (((size_t)b * Ho + oy) * Wo + ox) * ldcol + (ky * ksz + kx) * Cin + c8 * 8;
        __align__(16) T v[8];
        constexpr int NV = sizeof(T) * 8 / 16;
#pragma unroll
        for (int q = 0; q < NV; ++q) reinterpret_cast<uint4 *>(v)[q] = reinterpret_cast<const uint4 *>(src)[q];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += to_f<T>(v[j]);
      }
    }
    __align__(16) T outv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) outv[j] = from_f<T>(acc[j]);
    constexpr int NV = sizeof(T) * 8 / 16;
#pragma unroll
    for (int q = 0; q < NV; ++q) reinterpret_cast<uint4 *>(dx + o)[q] = reinterpret_cast<const uint4 *>(outv)[q];
  }
}

// ------------------------------------------------------------------ weights
template <typename T>
__global__ void prep_weight_kernel(const float *__restrict__ src, const float *__restrict__ scale, T *__restrict__ dst, int ldd,
                                   T *__restrict__ dstT, int lddT, int N, int K, int perm, int Cin, int ksz) {
  pdl_prologue();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * K) return;
  const int n = (int)(idx / K), k = (int)(idx % K);
  size_t si = idx;
  bool zero = false;
  if (perm == 1) {
    const int cin = k % Cin, kx = (k / Cin) % ksz, ky = k / (Cin * ksz);
    si = (((size_t)n * Cin + cin) * ksz + ky) * ksz + kx;
  } else if (perm == 2) {   // stem: k = (cin*ksz + ky)*8 + slot, slot 0 unused, slots 1..ksz <-> kx
    const int slot = k & 7, r = k >> 3;
    zero = slot == 0 || slot > ksz;
    si = (size_t)n * Cin * ksz * ksz + (size_t)r * ksz + (slot - 1);
  }
  float v = zero ? 0.f : src[si];
  if (scale) v *= scale[n];
  const T t = from_f<T>(v);
  if (dst) dst[(size_t)n * ldd + k] = t;
  if (dstT) dstT[(size_t)k * lddT + n] = t;
}

__global__ void scaled_bias_kernel(const float *b, const float *scale, float *out, int N) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) out[i] = b[i] * scale[i];
}

// y = x + gamma * (W a + b) was computed with folded weights; G = dY^T a and s = colsum(dY) are the
// unscaled accumulators.  dW += gamma*G, db += gamma*s, dgamma += rowsum(W.G) + b*s; G,s reset.
__global__ void ls_finalize_kernel(float *__restrict__ G, float *__restrict__ s, const float *__restrict__ W,
                                   const float *__restrict__ b, const float *__restrict__ gamma, float *__restrict__ dW,
                                   float *__restrict__ db, float *__restrict__ dgamma, int N, int K) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const float gm = gamma[n];
  float dot = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float g = G[(size_t)n * K + k];
    dot += W[(size_t)n * K + k] * g;
    dW[(size_t)n * K + k] += gm * g;
    G[(size_t)n * K + k] = 0.f;
  }
  dot = warp_sum(dot);
  if (lane == 0) {
    dgamma[n] += dot + b[n] * s[n];
    db[n] += gm * s[n];
    s[n] = 0.f;
  }
}

// ---- batched variants: one launch for many small tensors (the per-tensor launches above cost more in launch gaps than in work:
// 57 weight/bias tensors are prepared after every optimizer step, 16 LayerScale gradients are finalised after every backward)
template <typename T>
__global__ void prep_batch_kernel(const __grid_constant__ PrepBatch pb) {
  pdl_prologue();
  int it = 0;
  while (it + 1 < pb.n && (int)blockIdx.x >= pb.it[it + 1].first_block) ++it;
  const PrepItem &d = pb.it[it];
  const int64_t idx = (int64_t)(blockIdx.x - d.first_block) * blockDim.x + threadIdx.x;
  if (d.bias_out) {   // scaled bias: out = b * scale
    if (idx < d.N) d.bias_out[idx] = d.src[idx] * d.scale[idx];
    return;
  }
  if (idx >= (int64_t)d.N * d.K) return;
  const int n = (int)(idx / d.K), k = (int)(idx % d.K);
  size_t si = idx;
  bool zero = false;
  if (d.perm == 1) {
    const int cin = k % d.Cin, kx = (k / d.Cin) % d.ksz, ky = k / (d.Cin * d.ksz);
    si = (((size_t)n * d.Cin + cin) * d.ksz + ky) * d.ksz + kx;
  } else if (d.perm == 2) {
    const int slot = k & 7, r = k >> 3;
    zero = slot == 0 || slot > d.ksz;
    si = (size_t)n * d.Cin * d.ksz * d.ksz + (size_t)r * d.ksz + (slot - 1);
  }
  float v = zero ? 0.f : d.src[si];
  if (d.scale) v *= d.scale[n];
  const T t = from_f<T>(v);
  if (d.dst) reinterpret_cast<T *>(d.dst)[(size_t)n * d.ldd + k] = t;
  if (d.dstT) reinterpret_cast<T *>(d.dstT)[(size_t)k * d.lddT + n] = t;
}

__global__ void ls_finalize_batch_kernel(const __grid_constant__ LsBatch lb) {
  pdl_prologue();
  int it = 0;
  while (it + 1 < lb.n && (int)blockIdx.x >= lb.it[it + 1].first_block) ++it;
  const LsItem &d = lb.it[it];
  const int lane = threadIdx.x & 31;
  const int n = (blockIdx.x - d.first_block) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= d.N) return;
  const float gm = d.gamma[n];
  float dot = 0.f;
  for (int k = lane; k < d.K; k += 32) {
    const float g = d.G[(size_t)n * d.K + k];
    dot += d.W[(size_t)n * d.K + k] * g;
    d.dW[(size_t)n * d.K + k] += gm * g;
    d.G[(size_t)n * d.K + k] = 0.f;
  }
  dot = warp_sum(dot);
  if (lane == 0) {
    d.dgamma[n] += dot + d.b[n] * d.s[n];
    d.db[n] += gm * d.s[n];
    d.s[n] = 0.f;
  }
}

inline int blocks_for(int64_t n, int bs) { return (int)((n + bs - 1) / bs); }

// SM-side copy / clear for small buffers on the compute stream.  cudaMemcpyAsync / cudaMemsetAsync would go to a copy
// engine and queue behind whatever bulk host->device upload is in flight there (the next step's input batch).
__global__ void copy_u4_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, int64_t n16, const unsigned char *__restrict__ src_tail,
                               unsigned char *__restrict__ dst_tail, int ntail) {
  pdl_prologue();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n16) dst[i] = src[i];
  if (i < ntail) dst_tail[i] = src_tail[i];
}
__global__ void zero_u32_kernel(unsigned *__restrict__ p, int64_t n) {
  pdl_prologue();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0u;
}

}  // namespace

#define DISPATCH_T(dtype, ...)            \
  if ((dtype) == LEOD_F32) {              \
    typedef float T;                      \
    __VA_ARGS__;                          \
  } else {                                \
    typedef bf16 T;                       \
    __VA_ARGS__;                          \
  }

int layernorm_fwd(int dtype, const void *x, const float *w, const float *b, void *y, int M, int C, float eps, cudaStream_t st) {
  ProfScope ps(PK_LAYERNORM, 8.0 * M * C, 2.0 * M * C * dtype_size(dtype), st, M, C, 0);
  LEOD_REQUIRE(C <= LN_MAX_C, "layernorm: C=%d > %d", C, LN_MAX_C);
  if (dtype == LEOD_BF16 && C % 8 == 0 && ((((uintptr_t)x) | ((uintptr_t)y) | ((uintptr_t)w) | ((uintptr_t)b)) & 15) == 0) {
    const int nchunk = C / 8;
#define LN_FWD_FAST(LPR, NCH)                                                                                         \
  LEOD_LAUNCH((ln_fwd_bf16_kernel<LPR, NCH>), ceil_div(M, 8 * (32 / LPR)), 256, 0, st, (const bf16 *)x, w, b, (bf16 *)y, M, C, eps)
    if (nchunk <= 8) { LN_FWD_FAST(8, 1); } else if (nchunk <= 16) { LN_FWD_FAST(16, 1); } else if (nchunk <= 32) { LN_FWD_FAST(32, 1); }
    else { LN_FWD_FAST(32, 2); }
#undef LN_FWD_FAST
    LEOD_LAUNCH_CHECK();
    return 0;
  }
#define LN_FWD(NV) DISPATCH_T(dtype, (LEOD_LAUNCH((ln_fwd_kernel<T, NV>), ceil_div(M, 8), 256, 0, st, (const T *)x, w, b, (T *)y, M, C, eps)))
  const int nv = ceil_div(C, 32);
  if (nv <= 1) { LN_FWD(1); } else if (nv <= 2) { LN_FWD(2); } else if (nv <= 3) { LN_FWD(3); } else if (nv <= 4) { LN_FWD(4); }
  else if (nv <= 6) { LN_FWD(6); } else if (nv <= 8) { LN_FWD(8); } else if (nv <= 12) { LN_FWD(12); } else { LN_FWD(16); }
#undef LN_FWD
  LEOD_LAUNCH_CHECK();
  return 0;
}

int layernorm_bwd(int dtype, const void *x, const float *w, const void *dy, const void *dres, void *dx, float *dw, float *db,
                  int M, int C, float eps, cudaStream_t st) {
  ProfScope ps(PK_LAYERNORM, 16.0 * M * C, (dres ? 4.0 : 3.0) * M * C * dtype_size(dtype), st, M, C, 1);
  LEOD_REQUIRE(C <= LN_MAX_C, "layernorm: C=%d > %d", C, LN_MAX_C);
  if (dtype == LEOD_BF16 && C % 8 == 0 &&
      ((((uintptr_t)x) | ((uintptr_t)dy) | ((uintptr_t)dres) | ((uintptr_t)dx)) & 15) == 0) {
    const int nchunk = C / 8;
#define LN_BWD_FAST(LPR, NCH)                                                                                          \
  LEOD_LAUNCH((ln_bwd_bf16_kernel<LPR, NCH>), std::max(1, std::min(ceil_div(M, 8 * (32 / LPR)), 148 * 4)), 256, 0, st, \
      (const bf16 *)x, w, (const bf16 *)dy, (const bf16 *)dres, (bf16 *)dx, dw, db, M, C, eps)
    if (nchunk <= 8) { LN_BWD_FAST(8, 1); } else if (nchunk <= 16) { LN_BWD_FAST(16, 1); } else if (nchunk <= 32) { LN_BWD_FAST(32, 1); }
    else { LN_BWD_FAST(32, 2); }
#undef LN_BWD_FAST
    LEOD_LAUNCH_CHECK();
    return 0;
  }
  int blocks = ceil_div(M, 8);          // one row per warp until the grid covers the GPU a few times over
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
#define LN_BWD(NV)                                                                                                             \
  DISPATCH_T(dtype, (LEOD_LAUNCH((ln_bwd_kernel<T, NV>), blocks, 256, 0, st, (const T *)x, w, (const T *)dy, (const T *)dres, (T *)dx, dw, db, \
                                                                  M, C, eps)))
  const int nv = ceil_div(C, 32);
  if (nv <= 1) { LN_BWD(1); } else if (nv <= 2) { LN_BWD(2); } else if (nv <= 3) { LN_BWD(3); } else if (nv <= 4) { LN_BWD(4); }
  else if (nv <= 6) { LN_BWD(6); } else if (nv <= 8) { LN_BWD(8); } else if (nv <= 12) { LN_BWD(12); } else { LN_BWD(16); }
#undef LN_BWD
  LEOD_LAUNCH_CHECK();
  return 0;
}

int lstm_pointwise_fwd(int dtype, void *gates, const void *c_prev, void *h_out, void *c_out, int M, int C, cudaStream_t st) {
  ProfScope ps(PK_LSTM, 30.0 * M * C, 11.0 * M * C * dtype_size(dtype), st);
  const int64_t n = (int64_t)M * C;
  DISPATCH_T(dtype, (LEOD_LAUNCH((lstm_fwd_kernel<T>), blocks_for(n, 256), 256, 0, st, (T *)gates, (const T *)c_prev, (T *)h_out, (T *)c_out, M, C)));
  LEOD_LAUNCH_CHECK();
  return 0;
}

int lstm_pointwise_bwd(int dtype, const void *gates, const void *c_prev, const void *c_out, const void *dh, const void *dh2,
                       const void *dc, void *dgates, void *dc_prev, int M, int C, cudaStream_t st, const void *dh3) {
  ProfScope ps(PK_LSTM, 30.0 * M * C, 13.0 * M * C * dtype_size(dtype), st);
  const int64_t n = (int64_t)M * C;
  DISPATCH_T(dtype, (LEOD_LAUNCH((lstm_bwd_kernel<T>), blocks_for(n, 256), 256, 0, st, (const T *)gates, (const T *)c_prev, (const T *)c_out,
                                                                           (const T *)dh, (const T *)dh2, (const T *)dh3,
                                                                           (const T *)dc, (T *)dgates, (T *)dc_prev, M, C)));
  LEOD_LAUNCH_CHECK();
  return 0;
}

int add_tensors(int dtype, const void *a, const void *b, void *out, int64_t n, cudaStream_t st) {
  ProfScope ps(PK_OTHER, 0.0, 3.0 * n * dtype_size(dtype), st);
  DISPATCH_T(dtype, (LEOD_LAUNCH((add_kernel<T>), blocks_for(n, 256), 256, 0, st, (const T *)a, (const T *)b, (T *)out, n)));
  LEOD_LAUNCH_CHECK();
  return 0;
}

int im2col_nchw(int x_dtype, int dtype, const void *x, void *col, int B, int Cin, int xh, int xw, int Hp, int Wp, int ksz,
                int stride, int pad, int ldcol, cudaStream_t st) {
  ProfScope ps(PK_PATCH, 0.0, (double)B * (Hp / stride) * (Wp / stride) * ldcol * dtype_size(dtype) + (double)B * Cin * xh * xw * dtype_size(x_dtype), st, B, Cin, 0);
  const int Ho = (Hp + 2 * pad - ksz) / stride + 1, Wo = (Wp + 2 * pad - ksz) / stride + 1;
  LEOD_REQUIRE(ksz <= 7 && ldcol == Cin * ksz * 8, "im2col_nchw: patch pitch %d != Cin*ksz*8 (ksz=%d)", ldcol, ksz);
  const int64_t n = (int64_t)B * Ho * Wo * (ldcol / 8);
  int nb = blocks_for(n, 256);
  if (nb > 148 * 32) nb = 148 * 32;
#define IM2COL_CASE(TI)                                                                                                     \
  DISPATCH_T(dtype, (LEOD_LAUNCH((im2col_nchw_kernel<TI, T>), nb, 256, 0, st, (const TI *)x, (T *)col, B, Cin, xh, xw, Ho, Wo, ksz, stride, \
                                                                 pad, ldcol)))
  if (x_dtype == LEOD_U8) {
    IM2COL_CASE(uint8_t);
  } else if (x_dtype == LEOD_BF16) {
    IM2COL_CASE(bf16);
  } else {
    IM2COL_CASE(float);
  }
#undef IM2COL_CASE
  LEOD_LAUNCH_CHECK();
  return 0;
}

int im2col_nhwc(int dtype, const void *x, void *col, int B, int Hi, int Wi, int Cin, int ksz, int stride, int pad, int ldcol,
                cudaStream_t st) {
  ProfScope ps(PK_PATCH, 0.0, ((double)B * (Hi / stride) * (Wi / stride) * ldcol + (double)B * Hi * Wi * Cin) * dtype_size(dtype), st, B, Cin, 1);
  const int Ho = (Hi + 2 * pad - ksz) / stride + 1, Wo = (Wi + 2 * pad - ksz) / stride + 1;
  const int K = Cin * ksz * ksz;
  LEOD_REQUIRE(Cin % 8 == 0 && ldcol % 8 == 0, "im2col_nhwc: Cin=%d / pitch %d must be multiples of 8", Cin, ldcol);
  const int64_t n = (int64_t)B * Ho * Wo * (ldcol / 8);
  DISPATCH_T(dtype, (LEOD_LAUNCH((im2col_nhwc_kernel<T>), std::min(blocks_for(n, 256), 148 * 16), 256, 0, st, (const T *)x, (T *)col, B, Hi, Wi, Cin, Ho, Wo, ksz,
                                                                              stride, pad, K, ldcol)));
  LEOD_LAUNCH_CHECK();
  return 0;
}

int col2im_nhwc(int dtype, const void *dcol, int ldcol, const void *dres, void *dx, int B, int Hi, int Wi, int Cin, int ksz,
                int stride, int pad, cudaStream_t st) {
  ProfScope ps(PK_PATCH, 0.0, ((double)B * (Hi / stride) * (Wi / stride) * ldcol + (double)B * Hi * Wi * Cin) * dtype_size(dtype), st, B, Cin, 2);
  const int Ho = (Hi + 2 * pad - ksz) / stride + 1, Wo = (Wi + 2 * pad - ksz) / stride + 1;
  LEOD_REQUIRE(Cin % 8 == 0 && ldcol % 8 == 0, "col2im_nhwc: Cin=%d / pitch %d must be multiples of 8", Cin, ldcol);
  const int64_t n = (int64_t)B * Hi * Wi * (Cin / 8);
  DISPATCH_T(dtype, (LEOD_LAUNCH((col2im_nhwc_kernel<T>), std::min(blocks_for(n, 256), 148 * 16), 256, 0, st, (const T *)dcol, ldcol, (const T *)dres, (T *)dx, B,
                                                                              Hi, Wi, Cin, Ho, Wo, ksz, stride, pad)));
  LEOD_LAUNCH_CHECK();
  return 0;
}

int prep_weight(int dtype, const float *src, const float *scale, void *dst, int ldd, void *dstT, int lddT, int N, int K, int perm,
                int Cin, int ksz, cudaStream_t st) {
  const int64_t n = (int64_t)N * K;
  DISPATCH_T(dtype, (LEOD_LAUNCH((prep_weight_kernel<T>), blocks_for(n, 256), 256, 0, st, src, scale, (T *)dst, ldd, (T *)dstT, lddT, N, K, perm,
                                                                              Cin, ksz)));
  LEOD_LAUNCH_CHECK();
  return 0;
}

int prep_scaled_bias(const float *b, const float *scale, float *out, int N, cudaStream_t st) {
  LEOD_LAUNCH((scaled_bias_kernel), blocks_for(N, 128), 128, 0, st, b, scale, out, N);
  LEOD_LAUNCH_CHECK();
  return 0;
}

int prep_batch_add(PrepBatch &pb, int &blocks, const float *src, const float *scale, void *dst, int ldd, void *dstT, int lddT, int N, int K,
                   int perm, int Cin, int ksz, float *bias_out) {
  if (pb.n >= PREP_BATCH_MAX) return -1;
  PrepItem &d = pb.it[pb.n++];
  d.src = src; d.scale = scale; d.dst = dst; d.dstT = dstT; d.bias_out = bias_out;
  d.ldd = ldd; d.lddT = lddT; d.N = N; d.K = K; d.perm = perm; d.Cin = Cin; d.ksz = ksz;
  d.first_block = blocks;
  blocks += (int)(((int64_t)N * (bias_out ? 1 : K) + 255) / 256);
  return 0;
}
int prep_batch_launch(int dtype, const PrepBatch &pb, int blocks, cudaStream_t st) {
  if (pb.n == 0) return 0;
  DISPATCH_T(dtype, (LEOD_LAUNCH((prep_batch_kernel<T>), blocks, 256, 0, st, pb)));
  LEOD_LAUNCH_CHECK();
  return 0;
}
int ls_batch_add(LsBatch &lb, int &blocks, float *G, float *s, const float *W, const float *b, const float *gamma, float *dW, float *db,
                 float *dgamma, int N, int K) {
  if (lb.n >= LS_BATCH_MAX) return -1;
  LsItem &d = lb.it[lb.n++];
  d.G = G; d.s = s; d.W = W; d.b = b; d.gamma = gamma; d.dW = dW; d.db = db; d.dgamma = dgamma; d.N = N; d.K = K;
  d.first_block = blocks;
  blocks += ceil_div(N, 8);
  return 0;
}
int ls_batch_launch(const LsBatch &lb, int blocks, cudaStream_t st) {
  if (lb.n == 0) return 0;
  LEOD_LAUNCH((ls_finalize_batch_kernel), blocks, 256, 0, st, lb);
  LEOD_LAUNCH_CHECK();
  return 0;
}

int layerscale_grad_finalize(const float *G, const float *s, const float *W, const float *b, const float *gamma, float *dW,
                             float *db, float *dgamma, int N, int K, cudaStream_t st) {
  LEOD_LAUNCH((ls_finalize_kernel), ceil_div(N, 8), 256, 0, st, (float *)G, (float *)s, W, b, gamma, dW, db, dgamma, N, K);
  LEOD_LAUNCH_CHECK();
  return 0;
}

int device_copy(void *dst, const void *src, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return 0;
  if ((((uintptr_t)dst | (uintptr_t)src) & 15) != 0) {   // unaligned: leave it to the runtime
    LEOD_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  const int64_t n16 = (int64_t)(bytes / 16);
  const int ntail = (int)(bytes % 16);
  LEOD_LAUNCH((copy_u4_kernel), blocks_for(std::max<int64_t>(n16, ntail), 256), 256, 0, st, (const uint4 *)src, (uint4 *)dst, n16,
                                                                                (const unsigned char *)src + n16 * 16,
                                                                                (unsigned char *)dst + n16 * 16, ntail);
  LEOD_LAUNCH_CHECK();
  return 0;
}

int device_zero_u32(unsigned *p, int64_t n, cudaStream_t st) {
  if (n <= 0) return 0;
  LEOD_LAUNCH((zero_u32_kernel), blocks_for(n, 256), 256, 0, st, p, n);
  LEOD_LAUNCH_CHECK();
  return 0;
}
