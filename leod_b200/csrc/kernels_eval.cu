// Detection evaluation on the device: Prophesee box filter + COCO bounding-box matching / accumulation.
// Replaces utils/evaluation/prophesee/io/box_filtering.py:18-36, utils/evaluation/prophesee/evaluation.py:5-42,
// utils/evaluation/prophesee/metrics/coco_eval.py:32-120 and pycocotools.cocoeval.COCOeval (evaluate / accumulate, iouType bbox)
// for per-frame buffers (utils/evaluation/prophesee/evaluator.py:73-110: one entry per labelled frame).
//
//   eval_match_kernel    one CTA per (frame, category): filter, stable descending-score order (rank by counting), top 100,
//                        IoU in fp64, greedy matching for 4 area ranges x 10 IoU thresholds (one thread each), results as two 64-bit
//                        masks per detection (bit a*10+t: matched / ignored) + a sort key (descending score, then slot index)
//   eval_sort_*          bitonic sort of the keys per category (unique keys -> the stable merge order of COCOeval.accumulate)
//   eval_accum_kernel    one CTA per (category, area range, maxDets): cumulative TP/FP over the sorted detections, precision
//                        envelope (suffix maximum), precision at the 101 recall thresholds, recall — all in fp64 like numpy
// The 12 summary numbers are means over the small precision/recall arrays and are taken on the host (numpy's pairwise summation).
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace {
constexpr int EV_T = 10, EV_R = 101, EV_A = 4, EV_M = 3, EV_MAXDET = 100;
constexpr int EV_MAXG = 128;     // ground-truth boxes of one category in one frame
constexpr int EV_MAXD = 2048;    // detections of one category in one frame before the top-100 cut
constexpr unsigned long long EV_SENTINEL = ~0ull;

struct EvalParams {
  double iou_thrs[EV_T];
  double rec_thrs[EV_R];
  double area_lo[EV_A], area_hi[EV_A];
  int max_dets[EV_M];
  long long skip_ts;
  float min_diag_sq, min_side;
  int only_class;
};

__device__ __forceinline__ bool box_kept(long long t, float w, float h, const EvalParams &p) {
  // box_filtering.py:31-36, fp32 like the structured array fields
  const float dsq = __fadd_rn(__fmul_rn(w, w), __fmul_rn(h, h));
  return t > p.skip_ts && dsq >= p.min_diag_sq && w >= p.min_side && h >= p.min_side;
}

struct MatchSmem {
  double gbox[EV_MAXG][4];
  double garea[EV_MAXG];
  double dbox[EV_MAXDET][4];
  double darea[EV_MAXDET];
  float dscore[EV_MAXDET];
  int dsel[EV_MAXDET];
  unsigned long long dtm[EV_MAXDET], dig[EV_MAXDET];
  unsigned char gord[EV_A][EV_MAXG];
  unsigned char gig[EV_A][EV_MAXG];
  int cand[EV_MAXD];
  int ng, nd_all, nd, valid, overflow;
};

__global__ void __launch_bounds__(128) eval_match_kernel(const long long *__restrict__ gt_t, const float *__restrict__ gt_xywh,
                                                         const int *__restrict__ gt_cls, const int *__restrict__ gt_ptr,
                                                         const long long *__restrict__ dt_t, const float *__restrict__ dt_xywh,
                                                         const int *__restrict__ dt_cls, const float *__restrict__ dt_score,
                                                         const int *__restrict__ dt_ptr, int F, int K, const __grid_constant__ EvalParams p,
                                                         unsigned long long *__restrict__ keys, int Npad, unsigned long long *__restrict__ dtm_out,
                                                         unsigned long long *__restrict__ dig_out, int *__restrict__ npig, int *__restrict__ counts) {
  extern __shared__ __align__(16) unsigned char smraw[];
  MatchSmem &s = *reinterpret_cast<MatchSmem *>(smraw);
  double *iou = reinterpret_cast<double *>(smraw + ((sizeof(MatchSmem) + 15) & ~size_t(15)));     // [EV_MAXDET][EV_MAXG]
  const int f = blockIdx.x, k = blockIdx.y, tid = threadIdx.x;
  pdl_wait();
  const int g0 = gt_ptr[f], g1 = gt_ptr[f + 1], d0 = dt_ptr[f], d1 = dt_ptr[f + 1];
  // the frame is an image only if a ground-truth box survives the filter (coco_eval.py:55: np.unique(gt_boxes['t']))
  int any = 0;
  for (int i = g0 + tid; i < g1; i += blockDim.x)
    if ((p.only_class < 0 || gt_cls[i] == p.only_class) && box_kept(gt_t[i], gt_xywh[4 * i + 2], gt_xywh[4 * i + 3], p)) any = 1;
  any = __syncthreads_or(any);
  unsigned long long *kk = keys + (size_t)k * Npad + (size_t)f * EV_MAXDET;
  if (!any) {
    for (int d = tid; d < EV_MAXDET; d += blockDim.x) kk[d] = EV_SENTINEL;
    return;
  }
  if (tid == 0) {
    int ng = 0, nd = 0, over = 0;
    for (int i = g0; i < g1; ++i)
      if (gt_cls[i] == k && (p.only_class < 0 || k == p.only_class) && box_kept(gt_t[i], gt_xywh[4 * i + 2], gt_xywh[4 * i + 3], p)) {
        if (ng < EV_MAXG) {
          for (int c = 0; c < 4; ++c) s.gbox[ng][c] = (double)gt_xywh[4 * i + c];
          s.garea[ng] = (double)__fmul_rn(gt_xywh[4 * i + 2], gt_xywh[4 * i + 3]);     // coco_eval.py:166 area = w * h (fp32)
          ++ng;
        } else {
          over = 1;
        }
      }
    for (int i = d0; i < d1; ++i)
      if (dt_cls[i] == k && (p.only_class < 0 || k == p.only_class) && box_kept(dt_t[i], dt_xywh[4 * i + 2], dt_xywh[4 * i + 3], p)) {
        if (nd < EV_MAXD) s.cand[nd++] = i; else over = 1;
      }
    s.ng = ng; s.nd_all = nd; s.nd = min(nd, EV_MAXDET); s.overflow = over;
    if (k == 0) atomicAdd(&counts[0], 1);     // images
  }
  for (int d = tid; d < EV_MAXDET; d += blockDim.x) { s.dtm[d] = 0ull; s.dig[d] = 0ull; }
  __syncthreads();
  if (s.overflow && tid == 0) atomicExch(&counts[2], 1);
  const int ng = s.ng, nda = s.nd_all, nd = s.nd;
  // stable descending-score rank (np.argsort(-score, kind='mergesort')) by counting; the first 100 ranks are kept
  for (int i = tid; i < nda; i += blockDim.x) {
    const float si = dt_score[s.cand[i]];
    int rank = 0;
    for (int j = 0; j < nda; ++j) {
      const float sj = dt_score[s.cand[j]];
      rank += (sj > si) || (sj == si && j < i);
    }
    if (rank < EV_MAXDET) s.dsel[rank] = s.cand[i];
  }
  __syncthreads();
  for (int d = tid; d < nd; d += blockDim.x) {
    const int i = s.dsel[d];
    for (int c = 0; c < 4; ++c) s.dbox[d][c] = (double)dt_xywh[4 * i + c];
    s.darea[d] = (double)__fmul_rn(dt_xywh[4 * i + 2], dt_xywh[4 * i + 3]);          // COCO.loadRes: bb[2] * bb[3] (fp32)
    s.dscore[d] = dt_score[i];
  }
  // ground-truth order per area range: not-ignored first, stable (np.argsort(gtIg, kind='mergesort'))
  if (tid < EV_A) {
    const int a = tid;
    int n = 0, cnt = 0;
    for (int pass = 0; pass < 2; ++pass)
      for (int g = 0; g < ng; ++g) {
        const bool ig = s.garea[g] < p.area_lo[a] || s.garea[g] > p.area_hi[a];
        if ((int)ig == pass) {
          s.gord[a][n] = (unsigned char)g;
          s.gig[a][n] = ig;
          ++n;
          cnt += !ig;
        }
      }
    atomicAdd(&npig[k * EV_A + a], cnt);
  }
  __syncthreads();
  // maskApi.c bbIou (iscrowd 0), fp64
  for (int q = tid; q < nd * ng; q += blockDim.x) {
    const int d = q / ng, g = q - d * ng;
    const double *D = s.dbox[d], *G = s.gbox[g];
    double o = 0.0;
    const double w = fmin(D[2] + D[0], G[2] + G[0]) - fmax(D[0], G[0]);
    if (w > 0) {
      const double h = fmin(D[3] + D[1], G[3] + G[1]) - fmax(D[1], G[1]);
      if (h > 0) {
        const double in = __dmul_rn(w, h);
        o = in / (__dmul_rn(D[2], D[3]) + __dmul_rn(G[2], G[3]) - in);
      }
    }
    iou[d * EV_MAXG + g] = o;
  }
  __syncthreads();
  // COCOeval.evaluateImg: one thread per (area range, IoU threshold)
  if (tid < EV_A * EV_T) {
    const int a = tid / EV_T, t = tid % EV_T;
    unsigned int gtm[EV_MAXG / 32] = {0u, 0u, 0u, 0u};
    const double thr = fmin(p.iou_thrs[t], 1 - 1e-10);
    for (int d = 0; d < nd; ++d) {
      double best = thr;
      int m = -1;
      for (int gi = 0; gi < ng; ++gi) {
        if (gtm[gi >> 5] >> (gi & 31) & 1u) continue;
        if (m > -1 && !s.gig[a][m] && s.gig[a][gi]) break;
        const double v = iou[d * EV_MAXG + s.gord[a][gi]];
        if (v < best) continue;
        best = v;
        m = gi;
      }
      bool ig;
      if (m >= 0) {
        gtm[m >> 5] |= 1u << (m & 31);
        atomicOr(&s.dtm[d], 1ull << tid);
        ig = s.gig[a][m];
      } else {
        ig = s.darea[d] < p.area_lo[a] || s.darea[d] > p.area_hi[a];
      }
      if (ig) atomicOr(&s.dig[d], 1ull << tid);
    }
  }
  __syncthreads();
  const size_t slot0 = ((size_t)k * F + f) * EV_MAXDET;
  for (int d = tid; d < EV_MAXDET; d += blockDim.x) {
    if (d < nd) {
      const unsigned int bits = __float_as_uint(s.dscore[d]);
      const unsigned int ord = (bits & 0x80000000u) ? bits : ~bits;       // ascending key <=> descending score (negative scores last)
      kk[d] = ((unsigned long long)ord << 32) | (unsigned int)(f * EV_MAXDET + d);
      dtm_out[slot0 + d] = s.dtm[d];
      dig_out[slot0 + d] = s.dig[d];
    } else {
      kk[d] = EV_SENTINEL;
    }
  }
  if (tid == 0 && nd > 0) atomicAdd(&counts[1], nd);
  pdl_launch_dependents();
}

// ---------------------------------------------------------------- bitonic sort of uint64 keys, one array of Npad keys per blockIdx.y
constexpr int SORT_TILE = 2048;   // keys sorted inside one CTA's shared memory

__device__ __forceinline__ void cmpswap(unsigned long long &a, unsigned long long &b, bool up) {
  if ((a > b) == up) { const unsigned long long t = a; a = b; b = t; }
}

// phase 0: fully sort each tile (direction alternates so that the tiles form bitonic pairs); phase 1: finish a merge stage `kstage`
// whose remaining strides are < SORT_TILE
__global__ void __launch_bounds__(SORT_TILE / 2) eval_sort_local_kernel(unsigned long long *__restrict__ keys, int Npad, int kstage, int phase) {
  __shared__ unsigned long long sm[SORT_TILE];
  unsigned long long *a = keys + (size_t)blockIdx.y * Npad + (size_t)blockIdx.x * SORT_TILE;
  const int tid = threadIdx.x;
  pdl_wait();
  sm[tid] = a[tid];
  sm[tid + SORT_TILE / 2] = a[tid + SORT_TILE / 2];
  __syncthreads();
  const int base = blockIdx.x * SORT_TILE;
  if (phase == 0) {
    for (int k = 2; k <= SORT_TILE; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        const int i = 2 * tid - (tid & (j - 1));
        cmpswap(sm[i], sm[i + j], ((base + i) & k) == 0);
        __syncthreads();
      }
  } else {
    for (int j = SORT_TILE >> 1; j > 0; j >>= 1) {
      const int i = 2 * tid - (tid & (j - 1));
      cmpswap(sm[i], sm[i + j], ((base + i) & kstage) == 0);
      __syncthreads();
    }
  }
  a[tid] = sm[tid];
  a[tid + SORT_TILE / 2] = sm[tid + SORT_TILE / 2];
  pdl_launch_dependents();
}

__global__ void eval_sort_global_kernel(unsigned long long *__restrict__ keys, int Npad, int kstage, int j) {
  unsigned long long *a = keys + (size_t)blockIdx.y * Npad;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_wait();
  if (t < Npad / 2) {
    const int i = 2 * t - (t & (j - 1));
    unsigned long long x = a[i], y = a[i + j];
    if ((x > y) == ((i & kstage) == 0)) { a[i] = y; a[i + j] = x; }
  }
  pdl_launch_dependents();
}

// ---------------------------------------------------------------- accumulate
constexpr int ACC_THREADS = 256;

__global__ void __launch_bounds__(ACC_THREADS) eval_accum_kernel(const unsigned long long *__restrict__ keys, int Npad, int F, int K,
                                                                 const unsigned long long *__restrict__ dtm, const unsigned long long *__restrict__ dig,
                                                                 const int *__restrict__ npig_all, const __grid_constant__ EvalParams p,
                                                                 int *__restrict__ scratch, double *__restrict__ precision, double *__restrict__ recall) {
  __shared__ int s_tp[ACC_THREADS], s_fp[ACC_THREADS];
  __shared__ double s_max[ACC_THREADS + 1];
  __shared__ int s_n;
  const int k = blockIdx.x, a = blockIdx.y, m = blockIdx.z, tid = threadIdx.x;
  const int max_det = p.max_dets[m];
  const unsigned long long *kk = keys + (size_t)k * Npad;
  pdl_wait();
  const int npig = npig_all[k * EV_A + a];
  if (npig == 0) return;                       // COCOeval.accumulate: `if npig == 0: continue` (entries stay -1)
  // number of real detections of this category = first sentinel (keys are sorted)
  if (tid == 0) {
    int lo = 0, hi = Npad;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (kk[mid] == EV_SENTINEL) hi = mid; else lo = mid + 1;
    }
    s_n = lo;
  }
  __syncthreads();
  const int n = s_n;
  const int len = (n + ACC_THREADS - 1) / ACC_THREADS;
  const int i0 = min(n, tid * len), i1 = min(n, i0 + len);
  int *tp_arr = scratch + ((size_t)(k * EV_A + a) * EV_M + m) * 2 * (size_t)Npad;
  int *fp_arr = tp_arr + Npad;
  const size_t slot_base = (size_t)k * F * EV_MAXDET;
  const double eps = 2.220446049250313e-16;    // np.spacing(1)
  for (int t = 0; t < EV_T; ++t) {
    const int bit = a * EV_T + t;
    int tp = 0, fp = 0;
    for (int i = i0; i < i1; ++i) {
      const unsigned int slot = (unsigned int)(kk[i] & 0xffffffffull);
      if ((int)(slot % EV_MAXDET) >= max_det) continue;
      const bool mt = dtm[slot_base + slot] >> bit & 1ull, ig = dig[slot_base + slot] >> bit & 1ull;
      tp += mt && !ig;
      fp += !mt && !ig;
    }
    s_tp[tid] = tp;
    s_fp[tid] = fp;
    __syncthreads();
    if (tid == 0) {     // exclusive scan over the chunks
      int ct = 0, cf = 0;
      for (int c = 0; c < ACC_THREADS; ++c) {
        const int x = s_tp[c], y = s_fp[c];
        s_tp[c] = ct; s_fp[c] = cf;
        ct += x; cf += y;
      }
      recall[((t * K + k) * EV_A + a) * EV_M + m] = n > 0 ? (double)ct / (double)npig : 0.0;
    }
    __syncthreads();
    tp = s_tp[tid];
    fp = s_fp[tid];
    double cmax = 0.0;
    for (int i = i0; i < i1; ++i) {
      const unsigned int slot = (unsigned int)(kk[i] & 0xffffffffull);
      if ((int)(slot % EV_MAXDET) < max_det) {
        const bool mt = dtm[slot_base + slot] >> bit & 1ull, ig = dig[slot_base + slot] >> bit & 1ull;
        tp += mt && !ig;
        fp += !mt && !ig;
      }
      tp_arr[i] = tp;
      fp_arr[i] = fp;
      cmax = fmax(cmax, (double)tp / ((double)fp + (double)tp + eps));
    }
    s_max[tid] = cmax;
    __syncthreads();
    if (tid == 0) {     // s_max[c] <- maximum over chunks c+1 ..
      double run = 0.0;
      for (int c = ACC_THREADS - 1; c >= 0; --c) {
        const double x = s_max[c];
        s_max[c] = run;
        run = fmax(run, x);
      }
    }
    __syncthreads();
    for (int r = tid; r < EV_R; r += ACC_THREADS) {
      const double thr = p.rec_thrs[r];
      int lo = 0, hi = n;     // np.searchsorted(rc, thr, side='left') on rc = tp / npig
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((double)tp_arr[mid] / (double)npig < thr) lo = mid + 1; else hi = mid;
      }
      double q = 0.0;
      if (lo < n) {
        const int c = lo / len, ce = min(n, (c + 1) * len);
        q = s_max[c];
        for (int i = lo; i < ce; ++i) q = fmax(q, (double)tp_arr[i] / ((double)fp_arr[i] + (double)tp_arr[i] + eps));
      }
      precision[(((size_t)(t * EV_R + r) * K + k) * EV_A + a) * EV_M + m] = q;
    }
    __syncthreads();
  }
  pdl_launch_dependents();
}

__global__ void eval_init_kernel(double *precision, int np_, double *recall, int nr, int *ints, int ni) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_wait();
  if (i < np_) precision[i] = -1.0;
  if (i < nr) recall[i] = -1.0;
  if (i < ni) ints[i] = 0;
  pdl_launch_dependents();
}

int next_pow2(int64_t v) {
  int64_t p = SORT_TILE;
  while (p < v) p <<= 1;
  return (int)p;
}
}  // namespace

extern "C" int64_t leod_coco_eval_workspace_bytes(int F, int num_classes) {
  if (F <= 0 || num_classes <= 0) return 0;
  const int64_t Npad = next_pow2((int64_t)F * EV_MAXDET);
  const int64_t keys = (int64_t)num_classes * Npad * 8;
  const int64_t masks = 2 * (int64_t)num_classes * F * EV_MAXDET * 8;
  const int64_t scratch = (int64_t)num_classes * EV_A * EV_M * 2 * Npad * 4;
  return keys + masks + scratch + (int64_t)(num_classes * EV_A + 16) * 4;
}

extern "C" int leod_coco_eval(const int64_t *gt_t, const float *gt_xywh, const int32_t *gt_cls, const int32_t *gt_ptr, const int64_t *dt_t,
                              const float *dt_xywh, const int32_t *dt_cls, const float *dt_score, const int32_t *dt_ptr, int F, int num_classes,
                              int64_t skip_ts, int min_box_diag, int min_box_side, int only_class, const double *iou_thrs, const double *rec_thrs,
                              void *ws, double *precision, double *recall, int32_t *counts, void *stream) {
  LEOD_REQUIRE(gt_ptr && dt_ptr && ws && precision && recall && counts && iou_thrs && rec_thrs, "leod_coco_eval: null operand");
  LEOD_REQUIRE(F > 0 && num_classes > 0 && num_classes <= 16, "leod_coco_eval: F %d, classes %d", F, num_classes);
  LEOD_REQUIRE((int64_t)F * EV_MAXDET < (1ll << 30), "leod_coco_eval: too many frames (%d)", F);
  cudaStream_t st = (cudaStream_t)stream;
  const int K = num_classes;
  const int Npad = next_pow2((int64_t)F * EV_MAXDET);
  EvalParams p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < EV_T; ++i) p.iou_thrs[i] = iou_thrs[i];
  for (int i = 0; i < EV_R; ++i) p.rec_thrs[i] = rec_thrs[i];
  const double lo[EV_A] = {0.0, 0.0, 32.0 * 32.0, 96.0 * 96.0}, hi[EV_A] = {1e10, 32.0 * 32.0, 96.0 * 96.0, 1e10};   // cocoeval.py Params.setDetParams
  for (int a = 0; a < EV_A; ++a) { p.area_lo[a] = lo[a]; p.area_hi[a] = hi[a]; }
  p.max_dets[0] = 1; p.max_dets[1] = 10; p.max_dets[2] = 100;
  p.skip_ts = skip_ts;
  p.min_diag_sq = (float)(min_box_diag * min_box_diag);
  p.min_side = (float)min_box_side;
  p.only_class = only_class;
  uint8_t *w = (uint8_t *)ws;
  unsigned long long *keys = (unsigned long long *)w;                      w += (size_t)K * Npad * 8;
  unsigned long long *dtm = (unsigned long long *)w;                       w += (size_t)K * F * EV_MAXDET * 8;
  unsigned long long *dig = (unsigned long long *)w;                       w += (size_t)K * F * EV_MAXDET * 8;
  int *scratch = (int *)w;                                                 w += (size_t)K * EV_A * EV_M * 2 * Npad * 4;
  int *npig = (int *)w;
  const int np_ = EV_T * EV_R * K * EV_A * EV_M, nr = EV_T * K * EV_A * EV_M;
  ProfScope ps(PK_OTHER, 0.0, 0.0, st);
  LEOD_LAUNCH(eval_init_kernel, ceil_div(np_, 256), 256, 0, st, precision, np_, recall, nr, npig, K * EV_A);
  LEOD_LAUNCH_CHECK();
  LEOD_CUDA(cudaMemsetAsync(counts, 0, 3 * sizeof(int32_t), st));
  if ((int64_t)Npad > (int64_t)F * EV_MAXDET) {   // padding keys beyond the last frame
    LEOD_CUDA(cudaMemsetAsync(keys, 0xff, (size_t)K * Npad * 8, st));
  }
  const size_t smem = ((sizeof(MatchSmem) + 15) & ~size_t(15)) + (size_t)EV_MAXDET * EV_MAXG * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    LEOD_CUDA(cudaFuncSetAttribute(eval_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  LEOD_LAUNCH(eval_match_kernel, dim3(F, K), 128, smem, st, (const long long *)gt_t, gt_xywh, gt_cls, gt_ptr, (const long long *)dt_t, dt_xywh, dt_cls,
              dt_score, dt_ptr, F, K, p, keys, Npad, dtm, dig, npig, counts);
  LEOD_LAUNCH_CHECK();
  const dim3 tiles(Npad / SORT_TILE, K);
  LEOD_LAUNCH(eval_sort_local_kernel, tiles, SORT_TILE / 2, 0, st, keys, Npad, 0, 0);
  LEOD_LAUNCH_CHECK();
  for (int kstage = 2 * SORT_TILE; kstage <= Npad; kstage <<= 1) {
    for (int j = kstage >> 1; j >= SORT_TILE; j >>= 1) {
      LEOD_LAUNCH(eval_sort_global_kernel, dim3(ceil_div(Npad / 2, 256), K), 256, 0, st, keys, Npad, kstage, j);
      LEOD_LAUNCH_CHECK();
    }
    LEOD_LAUNCH(eval_sort_local_kernel, tiles, SORT_TILE / 2, 0, st, keys, Npad, kstage, 1);
    LEOD_LAUNCH_CHECK();
  }
  LEOD_LAUNCH(eval_accum_kernel, dim3(K, EV_A, EV_M), ACC_THREADS, 0, st, (const unsigned long long *)keys, Npad, F, K,
              (const unsigned long long *)dtm, (const unsigned long long *)dig, (const int *)npig, p, scratch, precision, recall);
  LEOD_LAUNCH_CHECK();
  return 0;
}
