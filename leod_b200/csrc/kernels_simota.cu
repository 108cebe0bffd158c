// YOLOX head output decode, SimOTA label assignment and the IoU / objectness / class losses with their gradients.
// Reference: models/detection/yolox/models/yolo_head.py:289-332 (grids, decode), :382-401 (_ignore_bbox),
// :403-597 / :776-972 (losses), :606-774 / :974-1148 (assignment, geometry constraint, dynamic-k matching),
// models/detection/yolox/models/losses.py:18-44 (IoU loss), models/detection/yolox/utils/boxes.py:89-113 (pairwise IoU).
//
// The reference loops over images and ground-truth boxes in Python with host synchronisation; here the same
// assignment runs as fixed-shape kernels with no host round trip:
//   head_decode        one thread per (image, anchor): decode + sigmoid, candidate / ignore flags (geometry constraint)
//   simota_match       one CTA per (image, gt): pairwise IoU + cost over all anchors in shared memory, dynamic k from the
//                      top-10 IoUs, k smallest costs (ties -> lower anchor index)
//   simota_resolve     one thread per (image, anchor): anchors claimed by several gts go to the cheapest one; targets,
//                      the three loss sums and the foreground count
//   loss_finalize      the six scalars of the reference's loss dict
//   loss_bwd           gradient w.r.t. the RAW head outputs, written straight into the padded-flat matrices the
//                      prediction convolutions' backward GEMMs read
// All arithmetic is fp32 (the reference disables autocast here, yolo_head.py:640).
#include <math.h>

#include "common.cuh"

namespace {

constexpr float COST_INF = __builtin_huge_valf();

struct Label {
  float cls, cx, cy, w, h;
  bool nonzero, valid;
};
// One zero-padded label row (cls, cx, cy, w, h, obj_conf, cls_conf): `nonzero` as yolo_head.py:410, the confidence
// thresholds of _ignore_bbox (:382-401) turn low-confidence pseudo labels into ignore boxes.
__device__ __forceinline__ Label load_label(const float *row, const SimotaCfg &c) {
  Label l;
  l.cls = row[0]; l.cx = row[1]; l.cy = row[2]; l.w = row[3]; l.h = row[4];
  const float sum = row[0] + row[1] + row[2] + row[3] + row[4] + row[5] + row[6];
  l.nonzero = sum > 0.f;
  bool ign = false;
  for (int i = 0; i < c.n_thresh; ++i) ign |= (l.cls == (float)i) && (row[5] < c.thresh[i] || row[6] < c.thresh[i]);
  if (ign && l.nonzero) l.cls = c.ignore_label;
  l.valid = l.nonzero && l.cls != c.ignore_label;
  return l;
}

__device__ __forceinline__ void anchor_pos(const HeadGeom &g, int a, int &lev, int &y, int &x) {
  lev = a < g.a0[1] ? 0 : (a < g.a0[2] ? 1 : 2);
  const int la = a - g.a0[lev];
  y = la / g.w[lev];
  x = la - y * g.w[lev];
}
__device__ __forceinline__ int64_t anchor_row(const HeadGeom &g, int b, int lev, int y, int x) {
  return (int64_t)b * g.P[lev] + (y + 1) * (g.w[lev] + 2) + (x + 1);
}
// centre-radius geometry constraint (yolo_head.py:702-732): anchor centre strictly inside the 1.5-stride box around the gt centre
__device__ __forceinline__ bool in_radius(const Label &l, float gx, float gy, float s) {
  const float cx = (gx + 0.5f) * s, cy = (gy + 0.5f) * s, r = s * 1.5f;
  const float d0 = cx - (l.cx - r), d1 = cy - (l.cy - r), d2 = (l.cx + r) - cx, d3 = (l.cy + r) - cy;
  return fminf(fminf(d0, d1), fminf(d2, d3)) > 0.f;
}
// boxes.py:89-113 (cxcywh): inter / (area_a + area_b - inter)
__device__ __forceinline__ float pair_iou(const Label &l, const float *p) {
  const float tlx = fmaxf(l.cx - l.w / 2, p[0] - p[2] / 2), tly = fmaxf(l.cy - l.h / 2, p[1] - p[3] / 2);
  const float brx = fminf(l.cx + l.w / 2, p[0] + p[2] / 2), bry = fminf(l.cy + l.h / 2, p[1] + p[3] / 2);
  const float en = (tlx < brx && tly < bry) ? 1.f : 0.f;
  const float inter = (brx - tlx) * (bry - tly) * en;
  return inter / (l.w * l.h + p[2] * p[3] - inter);
}
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }
// yolo_head.py:652-670: BCE(sqrt(sigmoid(cls) * sigmoid(obj)), onehot) summed over classes (log clamped at -100 as
// F.binary_cross_entropy) + 3 * -log(iou + 1e-8) + 1e6 * (outside the geometry constraint)
__device__ __forceinline__ float match_cost(const Label &l, const float *t, int C, float iou, bool inside) {
  const float so = sigmoid_acc(t[4]);
  const int gc = min(max((int)l.cls, 0), C - 1);
  float cc = 0.f;
  for (int j = 0; j < C; ++j) {
    const float p = sqrtf(sigmoid_acc(t[5 + j]) * so);
    cc += j == gc ? fmaxf(logf(p), -100.f) : fmaxf(logf(1.f - p), -100.f);
  }
  return (-cc + 3.0f * (-logf(iou + 1e-8f))) + (inside ? 0.f : 1e6f);
}

// ------------------------------------------------------------------ decode (+ geometry flags in training)
// raw[lev]: fp32 [R_lev, 8] padded-flat rows (reg x4, obj logit, cls logits).  preds [B, A, 5+C]: decoded, sigmoid scores
// (yolo_head.py:249-251, 310-332).  tout [B, A, 8]: decoded boxes + logits (training, :303-308).  flags [B, A]: bit 0 =
// inside the radius of a VALID gt (candidate), bit 1 = inside the radius of any gt.
__global__ void __launch_bounds__(256) head_decode_kernel(HeadPtrs rp, HeadGeom g, SimotaCfg cfg, float *__restrict__ preds, float *__restrict__ tout,
                                                         const float *__restrict__ labels, int nmax, uint8_t *__restrict__ flags,
                                                         int *__restrict__ match_cnt, int *__restrict__ match_gt) {
  pdl_prologue();
  extern __shared__ float s_lab[];   // [nmax][4]: cx, cy, flags (1 nonzero | 2 valid), unused
  const int b = blockIdx.y;
  if (labels) {
    for (int n = threadIdx.x; n < nmax; n += blockDim.x) {
      const Label l = load_label(labels + ((size_t)b * nmax + n) * 7, cfg);
      s_lab[n * 4 + 0] = l.cx; s_lab[n * 4 + 1] = l.cy;
      s_lab[n * 4 + 2] = (float)((l.nonzero ? 1 : 0) | (l.valid ? 2 : 0));
    }
    __syncthreads();
  }
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= g.A) return;
  int lev, y, x;
  anchor_pos(g, a, lev, y, x);
  const float *r = rp.raw[lev] + anchor_row(g, b, lev, y, x) * 8;
  const float4 r0 = reinterpret_cast<const float4 *>(r)[0], r1 = reinterpret_cast<const float4 *>(r)[1];
  const float s = (float)g.stride[lev];
  const float cx = (r0.x + (float)x) * s, cy = (r0.y + (float)y) * s, w = expf(r0.z) * s, h = expf(r0.w) * s;
  const float lg[4] = {r1.x, r1.y, r1.z, r1.w};
  if (preds) {
    float *po = preds + ((size_t)b * g.A + a) * (5 + g.C);
    po[0] = cx; po[1] = cy; po[2] = w; po[3] = h;
    for (int j = 0; j <= g.C; ++j) po[4 + j] = sigmoid_acc(lg[j]);
  }
  if (!tout) return;
  float4 *to = reinterpret_cast<float4 *>(tout + ((size_t)b * g.A + a) * 8);
  to[0] = make_float4(cx, cy, w, h);
  to[1] = r1;
  int f = 0;
  Label l;
  for (int n = 0; n < nmax; ++n) {
    const int lf = (int)s_lab[n * 4 + 2];
    if (!(lf & 1)) continue;
    l.cx = s_lab[n * 4 + 0]; l.cy = s_lab[n * 4 + 1];
    if (in_radius(l, (float)x, (float)y, s)) f |= (lf & 2) ? 3 : 2;
  }
  flags[(size_t)b * g.A + a] = (uint8_t)f;
  match_cnt[(size_t)b * g.A + a] = 0;
  match_gt[(size_t)b * g.A + a] = -1;
}

// block-wide arg-extreme over a shared array: smallest (MIN) / largest value, ties -> lower index.  All threads return the result.
template <bool MIN>
__device__ __forceinline__ void block_arg(const float *arr, int n, float &val, int &idx, float *s_val, int *s_idx) {
  float bv = MIN ? COST_INF : -COST_INF;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = arr[i];
    if (MIN ? (v < bv) : (v > bv)) { bv = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if ((MIN ? (ov < bv) : (ov > bv)) || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();   // previous round's readers are done with s_val / s_idx
  if (lane == 0) { s_val[warp] = bv; s_idx[warp] = bi; }
  __syncthreads();
  bv = s_val[0]; bi = s_idx[0];
  for (int w2 = 1; w2 < (int)(blockDim.x >> 5); ++w2) {
    const float ov = s_val[w2];
    const int oi = s_idx[w2];
    if ((MIN ? (ov < bv) : (ov > bv)) || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  val = bv; idx = bi;
}

// one CTA per (gt n, image b): dynamic-k matching of yolo_head.py:734-774
__global__ void __launch_bounds__(256) simota_match_kernel(HeadGeom g, SimotaCfg cfg, const float *__restrict__ tout, const float *__restrict__ labels,
                                                          int nmax, const uint8_t *__restrict__ flags, int *__restrict__ match_cnt,
                                                          int *__restrict__ match_gt) {
  pdl_prologue();
  extern __shared__ float s_buf[];
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  const int n = blockIdx.x, b = blockIdx.y;
  const Label l = load_label(labels + ((size_t)b * nmax + n) * 7, cfg);
  if (!l.valid) return;
  float *s_iou = s_buf, *s_cost = s_buf + g.A;
  for (int a = threadIdx.x; a < g.A; a += blockDim.x) {
    const int f = flags[(size_t)b * g.A + a];
    float iou = 0.f, cost = COST_INF;
    if (f & 1) {
      int lev, y, x;
      anchor_pos(g, a, lev, y, x);
      const float *t = tout + ((size_t)b * g.A + a) * 8;
      iou = pair_iou(l, t);
      cost = match_cost(l, t, g.C, iou, in_radius(l, (float)x, (float)y, (float)g.stride[lev]));
    }
    s_iou[a] = iou;
    s_cost[a] = cost;
  }
  __syncthreads();
  float sum = 0.f;
  const int kk = min(10, g.A);
  for (int k = 0; k < kk; ++k) {
    float v; int i;
    block_arg<false>(s_iou, g.A, v, i, s_val, s_idx);
    sum += v;
    if (threadIdx.x == 0) s_iou[i] = -1.f;
    __syncthreads();
  }
  const int dyn_k = max(1, (int)sum);
  for (int k = 0; k < dyn_k; ++k) {
    float c; int i;
    block_arg<true>(s_cost, g.A, c, i, s_val, s_idx);
    if (!(c < COST_INF)) break;
    if (threadIdx.x == 0) {
      s_cost[i] = COST_INF;
      atomicAdd(&match_cnt[(size_t)b * g.A + i], 1);
      atomicMax(&match_gt[(size_t)b * g.A + i], n);
    }
    __syncthreads();
  }
}

__device__ __forceinline__ float bce_logits(float x, float t) { return fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x))); }

// IoU of losses.py:18-44 (union + 1e-16) and its partial derivatives w.r.t. the predicted (cx, cy, w, h)
__device__ __forceinline__ float loss_iou(const float *p, const float *t, float *d /* null or [4] */) {
  const float ptlx = p[0] - p[2] / 2, ptly = p[1] - p[3] / 2, pbrx = p[0] + p[2] / 2, pbry = p[1] + p[3] / 2;
  const float gtlx = t[0] - t[2] / 2, gtly = t[1] - t[3] / 2, gbrx = t[0] + t[2] / 2, gbry = t[1] + t[3] / 2;
  const float tlx = fmaxf(ptlx, gtlx), tly = fmaxf(ptly, gtly), brx = fminf(pbrx, gbrx), bry = fminf(pbry, gbry);
  const float en = (tlx < brx && tly < bry) ? 1.f : 0.f;
  const float dx = brx - tlx, dy = bry - tly;
  const float I = dx * dy * en;
  const float Ap = p[2] * p[3], Ag = t[2] * t[3];
  const float U = Ap + Ag - I + 1e-16f;
  const float iou = I / U;
  if (d) {
    // torch.max / torch.min route the gradient to the selected operand (split evenly on exact ties)
    const float mtlx = ptlx > gtlx ? 1.f : (ptlx == gtlx ? 0.5f : 0.f), mtly = ptly > gtly ? 1.f : (ptly == gtly ? 0.5f : 0.f);
    const float mbrx = pbrx < gbrx ? 1.f : (pbrx == gbrx ? 0.5f : 0.f), mbry = pbry < gbry ? 1.f : (pbry == gbry ? 0.5f : 0.f);
    const float dI_dI = 1.f / U + I / (U * U);   // d iou / d I  (U contains -I)
    const float dI_dAp = -I / (U * U);
    const float Ix = dy * en, Iy = dx * en;      // d I / d (brx - tlx), d I / d (bry - tly)
    d[0] = dI_dI * Ix * (mbrx - mtlx);
    d[1] = dI_dI * Iy * (mbry - mtly);
    d[2] = dI_dI * Ix * 0.5f * (mbrx + mtlx) + dI_dAp * p[3];
    d[3] = dI_dI * Iy * 0.5f * (mbry + mtly) + dI_dAp * p[2];
  }
  return iou;
}

// conflict resolution (yolo_head.py:757-766) + targets + loss sums.  sums (double): [0] sum (1 - iou^2) over fg,
// [1] sum obj BCE over non-ignored anchors, [2] sum cls BCE over fg, [3] foreground anchors, [4] valid gts
__global__ void __launch_bounds__(256) simota_resolve_kernel(HeadGeom g, SimotaCfg cfg, const float *__restrict__ tout, const float *__restrict__ labels,
                                                            int nmax, const uint8_t *__restrict__ flags, const int *__restrict__ match_cnt,
                                                            const int *__restrict__ match_gt, int *__restrict__ assign, float *__restrict__ miou_out,
                                                            double *__restrict__ sums) {
  pdl_prologue();
  extern __shared__ float s_lab[];   // [nmax][7] raw rows
  __shared__ float s_red[8][4];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < nmax * 7; i += blockDim.x) s_lab[i] = labels[(size_t)b * nmax * 7 + i];
  __syncthreads();
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (a < g.A) {
    const size_t ia = (size_t)b * g.A + a;
    const float *t = tout + ia * 8;
    const int cnt = match_cnt[ia], f = flags[ia];
    int lev, y, x;
    anchor_pos(g, a, lev, y, x);
    int ns = -1;
    if (cnt == 1) {
      ns = match_gt[ia];
    } else if (cnt > 1) {
      float best = COST_INF;
      for (int n = 0; n < nmax; ++n) {
        const Label l = load_label(s_lab + n * 7, cfg);
        if (!l.valid) continue;
        const float c = match_cost(l, t, g.C, pair_iou(l, t), in_radius(l, (float)x, (float)y, (float)g.stride[lev]));
        if (c < best) { best = c; ns = n; }
      }
    }
    float miou = 0.f;
    const bool ignore = (f & 2) && !(f & 1);
    if (!ignore) acc[1] = bce_logits(t[4], ns >= 0 ? 1.f : 0.f);
    if (ns >= 0) {
      const Label l = load_label(s_lab + ns * 7, cfg);
      miou = pair_iou(l, t);
      const float tb[4] = {l.cx, l.cy, l.w, l.h};
      const float iou = loss_iou(t, tb, nullptr);
      acc[0] = 1.f - iou * iou;
      const int gc = min(max((int)l.cls, 0), g.C - 1);
      for (int j = 0; j < g.C; ++j) acc[2] += bce_logits(t[5 + j], j == gc ? miou : 0.f);
      acc[3] = 1.f;
    }
    assign[ia] = ns;
    miou_out[ia] = miou;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) acc[k] = warp_sum(acc[k]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
    for (int k = 0; k < 4; ++k) s_red[warp][k] = acc[k];
  __syncthreads();
  if (threadIdx.x < 4) {
    double tsum = 0.0;
    for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) tsum += (double)s_red[w2][threadIdx.x];
    if (tsum != 0.0) atomicAdd(&sums[threadIdx.x], tsum);
  }
  if (blockIdx.x == 0 && threadIdx.x == 32) {
    int nv = 0;
    for (int n = 0; n < nmax; ++n) nv += load_label(s_lab + n * 7, cfg).valid ? 1 : 0;
    if (nv) atomicAdd(&sums[4], (double)nv);
  }
}

// yolo_head.py:563-597: loss = reg_w * sum(1 - iou^2)/num_fg + obj_w * sum(bce_obj)/num_fg + cls_w * sum(bce_cls)/num_fg
__global__ void loss_finalize_kernel(const double *__restrict__ sums, SimotaCfg cfg, float *__restrict__ out) {
  pdl_prologue();
  const double nfg = fmax(sums[3], 1.0);
  const float li = (float)(cfg.reg_w * sums[0] / nfg), lo = (float)(cfg.obj_w * sums[1] / nfg), lc = (float)(cfg.cls_w * sums[2] / nfg);
  out[0] = li + lo + lc;
  out[1] = li; out[2] = lo; out[3] = lc; out[4] = 0.f;
  out[5] = (float)(nfg / fmax(sums[4], 1.0));
}

template <typename T>
__global__ void __launch_bounds__(256) loss_bwd_kernel(HeadGeom g, SimotaCfg cfg, const float *__restrict__ tout, const float *__restrict__ labels, int nmax,
                                                      const uint8_t *__restrict__ flags, const int *__restrict__ assign,
                                                      const float *__restrict__ miou_in, const double *__restrict__ sums,
                                                      const float *__restrict__ gscale, HeadGradPtrs dp) {
  pdl_prologue();
  const int b = blockIdx.y;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= g.A) return;
  const size_t ia = (size_t)b * g.A + a;
  const float *t = tout + ia * 8;
  const float k = (gscale ? gscale[0] : 1.f) / (float)fmax(sums[3], 1.0);
  const int ns = assign[ia], f = flags[ia];
  float d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const bool ignore = (f & 2) && !(f & 1);
  if (!ignore) d[4] = cfg.obj_w * k * (sigmoid_acc(t[4]) - (ns >= 0 ? 1.f : 0.f));
  int lev, y, x;
  anchor_pos(g, a, lev, y, x);
  if (ns >= 0) {
    const Label l = load_label(labels + ((size_t)b * nmax + ns) * 7, cfg);
    const float tb[4] = {l.cx, l.cy, l.w, l.h};
    float di[4];
    const float iou = loss_iou(t, tb, di);
    const float c = -2.f * iou * cfg.reg_w * k;
    const float s = (float)g.stride[lev];
    d[0] = c * di[0] * s;        // cx = (raw + grid) * stride
    d[1] = c * di[1] * s;
    d[2] = c * di[2] * t[2];     // w = exp(raw) * stride
    d[3] = c * di[3] * t[3];
    const int gc = min(max((int)l.cls, 0), g.C - 1);
    const float miou = miou_in[ia];
    for (int j = 0; j < g.C; ++j) d[5 + j] = cfg.cls_w * k * (sigmoid_acc(t[5 + j]) - (j == gc ? miou : 0.f));
  }
  T *o = (T *)dp.draw[lev] + anchor_row(g, b, lev, y, x) * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = from_f<T>(d[j]);
}

// [B, A, 8] fp32 <-> the padded-flat per-level raw-gradient matrices (diagnostics / custom losses)
template <typename T, bool SET>
__global__ void __launch_bounds__(256) raw_grad_copy_kernel(HeadGeom g, float *__restrict__ flat, HeadGradPtrs dp) {
  pdl_prologue();
  const int b = blockIdx.y;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= g.A) return;
  int lev, y, x;
  anchor_pos(g, a, lev, y, x);
  T *o = (T *)dp.draw[lev] + anchor_row(g, b, lev, y, x) * 8;
  float *f = flat + ((size_t)b * g.A + a) * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (SET) o[j] = from_f<T>(f[j]);
    else f[j] = to_f<T>(o[j]);
  }
}

}  // namespace

int raw_grad_copy(int dtype, const HeadGeom &g, float *flat, const HeadGradPtrs &dp, int set, cudaStream_t st) {
  dim3 grid(ceil_div(g.A, 256), g.B);
  if (dtype == LEOD_F32) {
    if (set) LEOD_LAUNCH((raw_grad_copy_kernel<float, true>), grid, 256, 0, st, g, flat, dp);
    else LEOD_LAUNCH((raw_grad_copy_kernel<float, false>), grid, 256, 0, st, g, flat, dp);
  } else {
    if (set) LEOD_LAUNCH((raw_grad_copy_kernel<bf16, true>), grid, 256, 0, st, g, flat, dp);
    else LEOD_LAUNCH((raw_grad_copy_kernel<bf16, false>), grid, 256, 0, st, g, flat, dp);
  }
  LEOD_LAUNCH_CHECK();
  return 0;
}

int head_decode(const HeadPtrs &rp, const HeadGeom &g, const SimotaCfg &cfg, float *preds, float *tout, const float *labels, int nmax,
                uint8_t *flags, int *match_cnt, int *match_gt, cudaStream_t st) {
  LEOD_REQUIRE(g.C >= 1 && g.C <= 3, "head_decode: num_classes %d (1..3 supported: 5 + C <= 8 output columns)", g.C);
  dim3 grid(ceil_div(g.A, 256), g.B);
  const size_t smem = tout ? (size_t)nmax * 4 * sizeof(float) : 0;
  LEOD_REQUIRE(smem <= 48 * 1024, "head_decode: %d labels per image exceed the shared-memory table", nmax);
  LEOD_LAUNCH((head_decode_kernel), grid, 256, smem, st, rp, g, cfg, preds, tout, tout ? labels : nullptr, nmax, flags, match_cnt, match_gt);
  LEOD_LAUNCH_CHECK();
  return 0;
}

int simota_loss_fwd(const HeadGeom &g, const SimotaCfg &cfg, const float *tout, const float *labels, int nmax, const uint8_t *flags,
                    int *match_cnt, int *match_gt, int *assign, float *miou, double *sums, float *losses, cudaStream_t st) {
  LEOD_REQUIRE(nmax >= 1 && (size_t)nmax * 7 * sizeof(float) <= 48 * 1024, "simota: %d labels per image", nmax);
  LEOD_TRY(device_zero_bytes(sums, 8 * sizeof(double), st));
  const size_t smem_m = (size_t)2 * g.A * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    LEOD_CUDA(cudaFuncSetAttribute(simota_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  LEOD_REQUIRE(smem_m <= 200 * 1024, "simota: %d anchors exceed the shared-memory cost table", g.A);
  LEOD_LAUNCH((simota_match_kernel), dim3(nmax, g.B), 256, smem_m, st, g, cfg, tout, labels, nmax, flags, match_cnt, match_gt);
  LEOD_LAUNCH_CHECK();
  LEOD_LAUNCH((simota_resolve_kernel), dim3(ceil_div(g.A, 256), g.B), 256, (size_t)nmax * 7 * sizeof(float), st, g, cfg, tout, labels, nmax, flags, match_cnt,
                                                                                                      match_gt, assign, miou, sums);
  LEOD_LAUNCH_CHECK();
  LEOD_LAUNCH((loss_finalize_kernel), 1, 1, 0, st, sums, cfg, losses);
  LEOD_LAUNCH_CHECK();
  return 0;
}

int simota_loss_bwd(int dtype, const HeadGeom &g, const SimotaCfg &cfg, const float *tout, const float *labels, int nmax, const uint8_t *flags,
                    const int *assign, const float *miou, const double *sums, const float *gscale, const HeadGradPtrs &dp, cudaStream_t st) {
  dim3 grid(ceil_div(g.A, 256), g.B);
  if (dtype == LEOD_F32)
    LEOD_LAUNCH((loss_bwd_kernel<float>), grid, 256, 0, st, g, cfg, tout, labels, nmax, flags, assign, miou, sums, gscale, dp);
  else
    LEOD_LAUNCH((loss_bwd_kernel<bf16>), grid, 256, 0, st, g, cfg, tout, labels, nmax, flags, assign, miou, sums, gscale, dp);
  LEOD_LAUNCH_CHECK();
  return 0;
}
