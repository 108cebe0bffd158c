// Input-side kernels of the step: small host->device uploads that stay off the copy engines, and the spatial / temporal data
// augmentation of the event tensors and their labels (data/utils/augmentor.py:125-478, data/genx_utils/labels.py:327-509,
// data/genx_utils/sequence_base.py:208-225) applied to a whole [L, B] batch that is already resident in HBM.
#include <string.h>

#include <algorithm>

#include "common.cuh"

// ------------------------------------------------------------------ small uploads
// The label / index tensors of a training step are a few KB.  As cudaMemcpyAsync they queue on the host->device copy engine behind
// the next batch's bulk upload (hundreds of MB) and stall the compute stream for milliseconds.  Pinned host memory is mapped into
// the device address space (UVA), so an SM kernel reads it over PCIe directly and the copy engines never see it.
__global__ void upload_small_kernel(const uint32_t *__restrict__ src_host, uint32_t *__restrict__ dst, int64_t nwords) {
  pdl_wait();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nwords; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src_host[i];
  pdl_launch_dependents();
}

extern "C" int leod_upload_small(void *dst, const void *src_pinned_host, int64_t nbytes, void *stream) {
  LEOD_REQUIRE(dst && src_pinned_host && nbytes >= 0, "leod_upload_small: null operand");
  LEOD_REQUIRE(nbytes % 4 == 0 && ((uintptr_t)dst & 3) == 0 && ((uintptr_t)src_pinned_host & 3) == 0,
               "leod_upload_small: buffers must be 4-byte aligned and a multiple of 4 bytes long (%lld)", (long long)nbytes);
  LEOD_REQUIRE(nbytes <= (64 << 20), "leod_upload_small: %lld bytes is a bulk transfer, use a copy stream", (long long)nbytes);
  if (nbytes == 0) return 0;
  cudaPointerAttributes at;
  LEOD_CUDA(cudaPointerGetAttributes(&at, src_pinned_host));
  LEOD_REQUIRE(at.type == cudaMemoryTypeHost && at.devicePointer != nullptr, "leod_upload_small: the source must be pinned (page-locked) host memory");
  const int64_t nw = nbytes / 4;
  const int grid = (int)std::min<int64_t>((nw + 255) / 256, 64);
  ProfScope ps(PK_OTHER, 0.0, 2.0 * nbytes, (cudaStream_t)stream);
  LEOD_LAUNCH(upload_small_kernel, grid, 256, 0, (cudaStream_t)stream, (const uint32_t *)at.devicePointer, (uint32_t *)dst, nw);
  LEOD_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ event-tensor augmentation
// One CTA per (frame, channel) plane.  The x index map of the sequence is built once per CTA in shared memory, rows are produced
// 16 output bytes per thread.  nearest-exact resampling as torch.nn.functional.interpolate computes it (fp32):
//   src = min(floor((dst + 0.5f) * (float(in) / float(out))), in - 1)
struct AugmBatch {
  leod_augm_state s[LEOD_AUGM_MAX_SEQ];
};

__device__ __forceinline__ int nearest_exact(int dst, float scale, int in_size) {
  const int v = (int)floorf(__fmul_rn((float)dst + 0.5f, scale));
  return v < in_size - 1 ? v : in_size - 1;
}

// source coordinate of output coordinate `o` along one axis, or -1 for the zero border of a zoom-out canvas
__device__ __forceinline__ int augm_src(int o, int size, int mode, int z0, int win, bool flip) {
  int s;
  if (mode == 1) {
    s = z0 + nearest_exact(o, (float)win / (float)size, win);
  } else if (mode == 2) {
    const int r = o - z0;
    if (r < 0 || r >= win) return -1;
    s = nearest_exact(r, (float)size / (float)win, size);
  } else {
    s = o;
  }
  return flip ? size - 1 - s : s;
}

__global__ void __launch_bounds__(256) augment_ev_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int L, int B, int C,
                                                         int H, int W, int b_base, const __grid_constant__ AugmBatch ab) {
  extern __shared__ int16_t xmap[];     // [W] source column (or -1)
  const int plane = blockIdx.x;         // ((l * Bc) + bl) * C + c over the sequences of this launch
  const int Bc = min(B - b_base, LEOD_AUGM_MAX_SEQ);
  const int c = plane % C;
  const int bl = (plane / C) % Bc;
  const int l = plane / (C * Bc);
  const leod_augm_state &s = ab.s[bl];
  const int b = b_base + bl;
  const int ls = s.t_flip ? L - 1 - l : l;
  const int cs = s.t_flip ? C - 1 - c : c;
  const uint8_t *src = in + (((int64_t)ls * B + b) * C + cs) * (int64_t)H * W;
  uint8_t *dst = out + (((int64_t)l * B + b) * C + c) * (int64_t)H * W;
  pdl_wait();
  for (int x = threadIdx.x; x < W; x += blockDim.x) xmap[x] = (int16_t)augm_src(x, W, s.zoom_mode, s.x0, s.win_w, s.h_flip != 0);
  __syncthreads();
  const bool vec = (W % 16 == 0) && (((uintptr_t)out & 15) == 0);
  if (vec) {
    const int cpr = W / 16;   // 16-byte chunks per row
    for (int i = threadIdx.x; i < H * cpr; i += blockDim.x) {
      const int y = i / cpr, x0 = (i - y * cpr) * 16;
      const int ys = augm_src(y, H, s.zoom_mode, s.y0, s.win_h, false);
      uint32_t w4[4] = {0u, 0u, 0u, 0u};
      if (ys >= 0) {
        const uint8_t *row = src + (int64_t)ys * W;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int xs = xmap[x0 + k];
          const uint32_t v = xs >= 0 ? (uint32_t)__ldg(row + xs) : 0u;
          w4[k >> 2] |= v << (8 * (k & 3));
        }
      }
      *reinterpret_cast<uint4 *>(dst + (int64_t)y * W + x0) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
    }
  } else {
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
      const int y = i / W, x = i - y * W;
      const int ys = augm_src(y, H, s.zoom_mode, s.y0, s.win_h, false);
      const int xs = xmap[x];
      dst[i] = (ys >= 0 && xs >= 0) ? __ldg(src + (int64_t)ys * W + xs) : (uint8_t)0;
    }
  }
  pdl_launch_dependents();
}

extern "C" int leod_augment_ev_repr(const void *in, void *out, int L, int B, int C, int H, int W, const leod_augm_state *states, void *stream) {
  LEOD_REQUIRE(in && out && states, "leod_augment_ev_repr: null operand");
  LEOD_REQUIRE(in != out, "leod_augment_ev_repr: in-place augmentation is not supported (the maps are gathers)");
  LEOD_REQUIRE(L > 0 && B > 0 && C > 0 && H > 0 && W > 0 && W <= 32767 && H <= 32767, "leod_augment_ev_repr: bad shape [%d,%d,%d,%d,%d]", L, B, C, H, W);
  for (int b = 0; b < B; ++b) {
    const leod_augm_state &s = states[b];
    LEOD_REQUIRE(s.zoom_mode >= 0 && s.zoom_mode <= 2, "leod_augment_ev_repr: sequence %d: zoom_mode %d", b, s.zoom_mode);
    if (s.zoom_mode != 0)   // augmentor.py:241-244, :322-326: the window lies inside the canvas
      LEOD_REQUIRE(s.x0 >= 0 && s.y0 >= 0 && s.win_h > 0 && s.win_w > 0 && s.x0 + s.win_w <= W && s.y0 + s.win_h <= H,
                   "leod_augment_ev_repr: sequence %d: zoom window (%d,%d)+(%d,%d) outside %dx%d", b, s.x0, s.y0, s.win_w, s.win_h, W, H);
  }
  const double bytes = 2.0 * L * B * C * H * W;
  ProfScope ps(PK_PATCH, 0.0, bytes, (cudaStream_t)stream, L * B, C, 3);
  for (int b0 = 0; b0 < B; b0 += LEOD_AUGM_MAX_SEQ) {
    AugmBatch ab;
    const int Bc = std::min(B - b0, (int)LEOD_AUGM_MAX_SEQ);
    memset(&ab, 0, sizeof(ab));
    for (int i = 0; i < Bc; ++i) ab.s[i] = states[b0 + i];
    LEOD_LAUNCH(augment_ev_kernel, L * Bc * C, 256, W * sizeof(int16_t), (cudaStream_t)stream, (const uint8_t *)in, (uint8_t *)out, L, B, C, H, W, b0,
                ab);
    LEOD_LAUNCH_CHECK();
  }
  return 0;
}

// ------------------------------------------------------------------ label augmentation
// One thread per ObjectLabels row (t, x, y, w, h, cls, cls_conf, obj).  Every operation is the fp32 operation torch performs, in the
// same order, with explicit round-to-nearest intrinsics (no FMA contraction), so the rows are bit-identical to the reference's:
//   flip_lr_ (labels.py:499-502)  ->  zoom_in_and_rescale_ (:371-411)  |  zoom_out_and_rescale_ (:437-459)  with scale_ (:482-497)
// keep[i] = 0 marks a row removed by remove_flat_labels_ (:67-69); removed rows keep their last computed values.
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

__device__ __forceinline__ bool label_scale(float &x, float &y, float &w, float &h, float mul, float cap_x, float cap_y) {
  const float x1 = fminf(__fmul_rn(__fadd_rn(x, w), mul), cap_x);
  const float y1 = fminf(__fmul_rn(__fadd_rn(y, h), mul), cap_y);
  x = __fmul_rn(x, mul);
  y = __fmul_rn(y, mul);
  w = __fsub_rn(x1, x);
  h = __fsub_rn(y1, y);
  return w > 0.f && h > 0.f;
}

__global__ void augment_labels_kernel(float *__restrict__ rows, const int32_t *__restrict__ row_seq, int64_t n, int b_base,
                                      const __grid_constant__ AugmBatch ab, uint8_t *__restrict__ keep) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  pdl_wait();
  if (i < n) {
    const int bl = row_seq[i] - b_base;
    if (bl >= 0 && bl < LEOD_AUGM_MAX_SEQ) {
      const leod_augm_state &s = ab.s[bl];
      float *r = rows + i * 8;
      float x = r[1], y = r[2], w = r[3], h = r[4];
      bool ok = true;
      if (s.h_flip) x = __fsub_rn(__fsub_rn(s.flip_c, x), w);
      if (s.zoom_mode == 1) {
        const float xa = clampf(x, s.lo_x, s.hi_x), ya = clampf(y, s.lo_y, s.hi_y);
        const float xb = clampf(__fadd_rn(x, w), s.lo_x, s.hi_x), yb = clampf(__fadd_rn(y, h), s.lo_y, s.hi_y);
        x = __fsub_rn(xa, s.lo_x);
        y = __fsub_rn(ya, s.lo_y);
        w = __fsub_rn(xb, xa);
        h = __fsub_rn(yb, ya);
        ok = w > 0.f && h > 0.f;
        if (ok) ok = label_scale(x, y, w, h, s.mul, s.cap_x, s.cap_y);
      } else if (s.zoom_mode == 2) {
        ok = label_scale(x, y, w, h, s.mul, s.cap_x, s.cap_y);
        if (ok) {
          x = __fadd_rn(x, (float)s.x0);
          y = __fadd_rn(y, (float)s.y0);
        }
      }
      r[1] = x; r[2] = y; r[3] = w; r[4] = h;
      keep[i] = ok ? 1 : 0;
    }
  }
  pdl_launch_dependents();
}

extern "C" int leod_augment_labels(float *rows, const int32_t *row_seq, int64_t n, int B, const leod_augm_state *states, uint8_t *keep,
                                   void *stream) {
  LEOD_REQUIRE(states && B > 0 && n >= 0, "leod_augment_labels: bad arguments");
  if (n == 0) return 0;
  LEOD_REQUIRE(rows && row_seq && keep, "leod_augment_labels: null operand");
  ProfScope ps(PK_OTHER, 0.0, 70.0 * n, (cudaStream_t)stream);
  for (int b0 = 0; b0 < B; b0 += LEOD_AUGM_MAX_SEQ) {
    AugmBatch ab;
    const int Bc = std::min(B - b0, (int)LEOD_AUGM_MAX_SEQ);
    memset(&ab, 0, sizeof(ab));
    for (int i = 0; i < Bc; ++i) ab.s[i] = states[b0 + i];
    LEOD_LAUNCH(augment_labels_kernel, ceil_div(n, 256), 256, 0, (cudaStream_t)stream, rows, row_seq, n, b0, ab, keep);
    LEOD_LAUNCH_CHECK();
  }
  return 0;
}
