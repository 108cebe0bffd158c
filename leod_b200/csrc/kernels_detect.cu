// Detection post-processing on the device, one CTA per image, no host round trip:
//   confidence filter -> stable descending-score order -> class-aware greedy NMS  (boxes.py:32-86 +
//   torchvision.ops.batched_nms), and the pseudo-label filters of modules/utils/ssod.py:113-188.
// All box arithmetic uses explicit round-to-nearest intrinsics so that nvcc cannot contract
// multiply-adds: kept index lists must be bit-exact against the fp32 reference.
#include "common.cuh"

namespace {

constexpr int PP_THREADS = 512;

__global__ void __launch_bounds__(PP_THREADS) postprocess_kernel(const float *__restrict__ pred, int A, int ncls, float conf_thre,
                                                                 float nms_thre, int class_agnostic, float *__restrict__ out,
                                                                 int32_t *__restrict__ count, int max_det) {
  pdl_prologue();
  extern __shared__ float sm[];
  // per candidate: score, anchor (unsorted) | sorted: anchor, offset box[4], suppressed flag
  float *c_score = sm;                          // [A]
  int *c_anchor = (int *)(c_score + A);         // [A]
  int *s_anchor = c_anchor + A;                 // [A]
  float *s_box = (float *)(s_anchor + A);       // [4][A] (x1 | y1 | x2 | y2 planes: conflict-free in the NMS loop)
  unsigned char *s_sup = (unsigned char *)(s_box + 4 * (size_t)A);  // [A]
  __shared__ int n_cand, n_keep;
  __shared__ float red[PP_THREADS / 32];
  __shared__ float max_coord_s;

  const int b = blockIdx.x, tid = threadIdx.x;
  const int stride = 5 + ncls;
  const float *p = pred + (size_t)b * A * stride;
  if (tid == 0) { n_cand = 0; n_keep = 0; }
  __syncthreads();

  float local_max = -INFINITY;
  for (int a = tid; a < A; a += PP_THREADS) {
    const float *r = p + (size_t)a * stride;
    float best = r[5];
    for (int c = 1; c < ncls; ++c) best = fmaxf(best, r[5 + c]);  // value only; index recomputed at output
    const float score = __fmul_rn(r[4], best);
    if (score >= conf_thre) {
      const int slot = atomicAdd(&n_cand, 1);
      c_score[slot] = score;
      c_anchor[slot] = a;
      const float hw = __fmul_rn(r[2], 0.5f), hh = __fmul_rn(r[3], 0.5f);
      const float x1 = __fsub_rn(r[0], hw), y1 = __fsub_rn(r[1], hh), x2 = __fadd_rn(r[0], hw), y2 = __fadd_rn(r[1], hh);
      local_max = fmaxf(local_max, fmaxf(fmaxf(x1, y1), fmaxf(x2, y2)));
    }
  }
  local_max = warp_max(local_max);
  if ((tid & 31) == 0) red[tid >> 5] = local_max;
  __syncthreads();
  if (tid == 0) {
    float m = red[0];
    for (int i = 1; i < PP_THREADS / 32; ++i) m = fmaxf(m, red[i]);
    max_coord_s = m;
  }
  __syncthreads();
  const int n = n_cand;
  const float off_unit = class_agnostic ? 0.f : __fadd_rn(max_coord_s, 1.0f);

  // stable descending order by counting: rank = #better candidates
  for (int i = tid; i < n; i += PP_THREADS) {
    const float si = c_score[i];
    const int ai = c_anchor[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const float sj = c_score[j];
      rank += (sj > si) || (sj == si && c_anchor[j] < ai);
    }
    const float *r = p + (size_t)ai * stride;
    int cls = 0;
    float best = r[5];
    for (int c = 1; c < ncls; ++c)
      if (r[5 + c] > best) { best = r[5 + c]; cls = c; }
    const float off = __fmul_rn((float)cls, off_unit);
    const float hw = __fmul_rn(r[2], 0.5f), hh = __fmul_rn(r[3], 0.5f);
    s_anchor[rank] = ai;
    s_box[rank] = __fadd_rn(__fsub_rn(r[0], hw), off);
    s_box[A + rank] = __fadd_rn(__fsub_rn(r[1], hh), off);
    s_box[2 * A + rank] = __fadd_rn(__fadd_rn(r[0], hw), off);
    s_box[3 * A + rank] = __fadd_rn(__fadd_rn(r[1], hh), off);
    s_sup[rank] = 0;
  }
  __syncthreads();

  // greedy NMS: visit in score order; the CTA only synchronises when a box is kept
  float *o = out + (size_t)b * max_det * 7;
  for (int i = 0; i < n; ++i) {
    if (s_sup[i]) continue;  // uniform: flags only change before a barrier
    const int kslot = n_keep;
    if (kslot >= max_det) break;
    const float ix1 = s_box[i], iy1 = s_box[A + i], ix2 = s_box[2 * A + i], iy2 = s_box[3 * A + i];
    const float iarea = __fmul_rn(__fsub_rn(ix2, ix1), __fsub_rn(iy2, iy1));
    for (int j = i + 1 + tid; j < n; j += PP_THREADS) {
      if (s_sup[j]) continue;
      const float jx1 = s_box[j], jy1 = s_box[A + j], jx2 = s_box[2 * A + j], jy2 = s_box[3 * A + j];
      const float w = fmaxf(__fsub_rn(fminf(ix2, jx2), fmaxf(ix1, jx1)), 0.f);
      const float h = fmaxf(__fsub_rn(fminf(iy2, jy2), fmaxf(iy1, jy1)), 0.f);
      const float inter = __fmul_rn(w, h);
      const float jarea = __fmul_rn(__fsub_rn(jx2, jx1), __fsub_rn(jy2, jy1));
      const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(iarea, jarea), inter));
      if (iou > nms_thre) s_sup[j] = 1;
    }
    if (tid == 0) {      // the output rows are written after the loop: no global-memory latency inside the serial chain
      c_anchor[kslot] = i;
      n_keep = kslot + 1;
    }
    __syncthreads();
  }
  __syncthreads();
  const int nk = n_keep;
  for (int k = tid; k < nk; k += PP_THREADS) {
    const float *r = p + (size_t)s_anchor[c_anchor[k]] * stride;
    int cls = 0;
    float best = r[5];
    for (int c = 1; c < ncls; ++c)
      if (r[5 + c] > best) { best = r[5 + c]; cls = c; }
    const float hw = __fmul_rn(r[2], 0.5f), hh = __fmul_rn(r[3], 0.5f);
    float *d = o + (size_t)k * 7;
    d[0] = __fsub_rn(r[0], hw); d[1] = __fsub_rn(r[1], hh); d[2] = __fadd_rn(r[0], hw); d[3] = __fadd_rn(r[1], hh);
    d[4] = r[4]; d[5] = best; d[6] = (float)cls;
  }
  if (tid == 0) count[b] = n_keep;
}

struct P2LArgs {
  float obj_thr[16], cls_thr[16];
};

// one warp per image; order-preserving compaction with ballots
__global__ void pred2label_kernel(const float *__restrict__ dets, const int32_t *__restrict__ count, int max_det, int ncls,
                                  P2LArgs th, int frame_h, int frame_w, float *__restrict__ labels, int32_t *__restrict__ lab_count) {
  pdl_prologue();
  const int b = blockIdx.x, lane = threadIdx.x;
  const int n = count[b];
  const float *d = dets + (size_t)b * max_det * 7;
  float *o = labels + (size_t)b * max_det * 8;
  int base = 0;
  for (int r0 = 0; r0 < n; r0 += 32) {
    const int r = r0 + lane;
    bool sel = false;
    float x1 = 0, y1 = 0, x2 = 0, y2 = 0, obj = 0, cc = 0, cls = 0;
    if (r < n) {
      x1 = d[r * 7 + 0]; y1 = d[r * 7 + 1]; x2 = d[r * 7 + 2]; y2 = d[r * 7 + 3];
      obj = d[r * 7 + 4]; cc = d[r * 7 + 5]; cls = d[r * 7 + 6];
      const int ci = (int)cls;
      sel = ci >= 0 && ci < ncls && obj > th.obj_thr[ci] && cc > th.cls_thr[ci];
      if (frame_h > 0 && frame_w > 0) {
        const float mx = (float)(frame_w - 1), my = (float)(frame_h - 1);
        x1 = fminf(fmaxf(x1, 0.f), mx); x2 = fminf(fmaxf(x2, 0.f), mx);
        y1 = fminf(fmaxf(y1, 0.f), my); y2 = fminf(fmaxf(y2, 0.f), my);
        const float w = __fsub_rn(x2, x1), h = __fsub_rn(y2, y1);
        sel = sel && w > 0.f && h > 0.f && w >= 5.f && h >= 5.f && w <= (float)((9 * frame_w) / 10);
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, sel);
    if (sel) {
      float *q = o + (size_t)(base + __popc(m & ((1u << lane) - 1))) * 8;
      q[0] = 0.f; q[1] = x1; q[2] = y1; q[3] = __fsub_rn(x2, x1); q[4] = __fsub_rn(y2, y1);
      q[5] = cls; q[6] = cc; q[7] = obj;
    }
    base += __popc(m);
  }
  if (lane == 0) lab_count[b] = base;
}


// Second NMS over the concatenated TTA views of one frame (modules/pseudo_labeler.py:37-91): rows are ObjectLabels
// records (t, x, y, w, h, cls, cls_conf, obj); one CTA per frame.  Same order of float operations as the reference:
// x2 = x + w, score = obj * cls_conf, offsets cls * (max coordinate of the kept candidates + 1), IoU > thr drops.
constexpr int TM_THREADS = 128;
__global__ void __launch_bounds__(TM_THREADS) tta_merge_kernel(const float *__restrict__ labels, const int32_t *__restrict__ count,
                                                               int nmax, float conf_thre, float nms_thre, int class_agnostic,
                                                               float *__restrict__ out, int32_t *__restrict__ out_count) {
  pdl_prologue();
  extern __shared__ float sm[];
  float *c_score = sm;                      // [nmax]
  int *c_row = (int *)(c_score + nmax);     // [nmax]
  int *s_row = c_row + nmax;                // [nmax]
  float *s_box = (float *)(s_row + nmax);   // [4 nmax]
  unsigned char *s_sup = (unsigned char *)(s_box + 4 * (size_t)nmax);
  __shared__ int n_cand, n_keep;
  __shared__ float red[TM_THREADS / 32];
  __shared__ float max_coord_s;
  const int f = blockIdx.x, tid = threadIdx.x;
  const int n_in = min(count[f], nmax);
  const float *L = labels + (size_t)f * nmax * 8;
  float *O = out + (size_t)f * nmax * 8;
  // frames that carry ground truth (t > 0 anywhere) pass through untouched (pseudo_labeler.py:50-53)
  __shared__ int has_gt;
  if (tid == 0) { n_cand = 0; n_keep = 0; has_gt = 0; }
  __syncthreads();
  for (int r = tid; r < n_in; r += TM_THREADS)
    if (L[r * 8] > 0.f) has_gt = 1;
  __syncthreads();
  if (has_gt) {
    for (int i = tid; i < n_in * 8; i += TM_THREADS) O[i] = L[i];
    if (tid == 0) out_count[f] = n_in;
    return;
  }
  float local_max = -INFINITY;
  for (int r = tid; r < n_in; r += TM_THREADS) {
    const float *q = L + r * 8;
    const float score = __fmul_rn(q[7], q[6]);
    if (score >= conf_thre) {
      const int slot = atomicAdd(&n_cand, 1);
      c_score[slot] = score;
      c_row[slot] = r;
      const float x2 = __fadd_rn(q[1], q[3]), y2 = __fadd_rn(q[2], q[4]);
      local_max = fmaxf(local_max, fmaxf(fmaxf(q[1], q[2]), fmaxf(x2, y2)));
    }
  }
  local_max = warp_max(local_max);
  if ((tid & 31) == 0) red[tid >> 5] = local_max;
  __syncthreads();
  if (tid == 0) {
    float m = red[0];
    for (int i = 1; i < TM_THREADS / 32; ++i) m = fmaxf(m, red[i]);
    max_coord_s = m;
  }
  __syncthreads();
  const int n = n_cand;
  const float off_unit = class_agnostic ? 0.f : __fadd_rn(max_coord_s, 1.0f);
  for (int i = tid; i < n; i += TM_THREADS) {
    const float si = c_score[i];
    const int ri = c_row[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const float sj = c_score[j];
      rank += (sj > si) || (sj == si && c_row[j] < ri);
    }
    const float *q = L + ri * 8;
    const float off = __fmul_rn(q[5], off_unit);
    s_row[rank] = ri;
    s_box[4 * rank + 0] = __fadd_rn(q[1], off);
    s_box[4 * rank + 1] = __fadd_rn(q[2], off);
    s_box[4 * rank + 2] = __fadd_rn(__fadd_rn(q[1], q[3]), off);
    s_box[4 * rank + 3] = __fadd_rn(__fadd_rn(q[2], q[4]), off);
    s_sup[rank] = 0;
  }
  __syncthreads();
  for (int i = 0; i < n; ++i) {
    if (s_sup[i]) continue;
    const int kslot = n_keep;
    const float ix1 = s_box[4 * i], iy1 = s_box[4 * i + 1], ix2 = s_box[4 * i + 2], iy2 = s_box[4 * i + 3];
    const float iarea = __fmul_rn(__fsub_rn(ix2, ix1), __fsub_rn(iy2, iy1));
    for (int j = i + 1 + tid; j < n; j += TM_THREADS) {
      if (s_sup[j]) continue;
      const float jx1 = s_box[4 * j], jy1 = s_box[4 * j + 1], jx2 = s_box[4 * j + 2], jy2 = s_box[4 * j + 3];
      const float w = fmaxf(__fsub_rn(fminf(ix2, jx2), fmaxf(ix1, jx1)), 0.f);
      const float h = fmaxf(__fsub_rn(fminf(iy2, jy2), fmaxf(iy1, jy1)), 0.f);
      const float inter = __fmul_rn(w, h);
      const float jarea = __fmul_rn(__fsub_rn(jx2, jx1), __fsub_rn(jy2, jy1));
      const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(iarea, jarea), inter));
      if (iou > nms_thre) s_sup[j] = 1;
    }
    if (tid == 0) {
      c_row[kslot] = i;
      n_keep = kslot + 1;
    }
    __syncthreads();
  }
  __syncthreads();
  const int nk = n_keep;
  for (int k = tid; k < nk; k += TM_THREADS) {
    const float *q = L + s_row[c_row[k]] * 8;
    float *d = O + (size_t)k * 8;
    const float x2 = __fadd_rn(q[1], q[3]), y2 = __fadd_rn(q[2], q[4]);
    d[0] = q[0]; d[1] = q[1]; d[2] = q[2]; d[3] = __fsub_rn(x2, q[1]); d[4] = __fsub_rn(y2, q[2]);
    d[5] = q[5]; d[6] = q[6]; d[7] = q[7];
  }
  if (tid == 0) out_count[f] = n_keep;
}

}  // namespace

extern "C" int leod_postprocess(const float *pred, int B, int A, int num_classes, float conf_thre, float nms_thre,
                                int class_agnostic, float *out, int32_t *count, int max_det, void *stream) {
  LEOD_REQUIRE(pred && out && count, "leod_postprocess: null argument");
  LEOD_REQUIRE(B >= 0 && A > 0 && num_classes > 0 && max_det > 0 && max_det <= A, "leod_postprocess: bad sizes B=%d A=%d C=%d max_det=%d",
               B, A, num_classes, max_det);
  if (B == 0) return 0;
  const size_t smem = (size_t)A * (4 + 4 + 4 + 16 + 1) + 16;
  LEOD_REQUIRE(smem <= 220 * 1024, "leod_postprocess: %d anchors per image exceed the shared-memory budget (max ~7700)", A);
  static size_t smem_set = 0;
  if (smem > smem_set) {
    LEOD_CUDA(cudaFuncSetAttribute(postprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  LEOD_LAUNCH((postprocess_kernel), B, PP_THREADS, smem, (cudaStream_t)stream, pred, A, num_classes, conf_thre, nms_thre, class_agnostic, out,
                                                                    count, max_det);
  LEOD_LAUNCH_CHECK();
  return 0;
}

extern "C" int leod_pred2label(const float *dets, const int32_t *count, int B, int max_det, int num_classes,
                               const float *obj_thresh, const float *cls_thresh, int frame_h, int frame_w, float *labels,
                               int32_t *lab_count, void *stream) {
  LEOD_REQUIRE(dets && count && labels && lab_count && obj_thresh && cls_thresh, "leod_pred2label: null argument");
  LEOD_REQUIRE(num_classes > 0 && num_classes <= 16, "leod_pred2label: num_classes %d not in [1,16]", num_classes);
  if (B == 0) return 0;
  P2LArgs th;
  for (int i = 0; i < 16; ++i) {
    th.obj_thr[i] = i < num_classes ? obj_thresh[i] : 2.f;
    th.cls_thr[i] = i < num_classes ? cls_thresh[i] : 2.f;
  }
  LEOD_LAUNCH((pred2label_kernel), B, 32, 0, (cudaStream_t)stream, dets, count, max_det, num_classes, th, frame_h, frame_w, labels, lab_count);
  LEOD_LAUNCH_CHECK();
  return 0;
}

extern "C" int leod_tta_merge(const float *labels, const int32_t *count, int F, int nmax, float conf_thre, float nms_thre,
                              int class_agnostic, float *out, int32_t *out_count, void *stream) {
  LEOD_REQUIRE(labels && count && out && out_count, "leod_tta_merge: null argument");
  LEOD_REQUIRE(F >= 0 && nmax > 0, "leod_tta_merge: bad sizes F=%d nmax=%d", F, nmax);
  if (F == 0) return 0;
  const size_t smem = (size_t)nmax * (4 + 4 + 4 + 16 + 1) + 16;
  LEOD_REQUIRE(smem <= 220 * 1024, "leod_tta_merge: %d boxes per frame exceed the shared-memory budget", nmax);
  static size_t smem_set = 0;
  if (smem > smem_set && smem > 48 * 1024) {
    LEOD_CUDA(cudaFuncSetAttribute(tta_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  LEOD_LAUNCH((tta_merge_kernel), F, TM_THREADS, smem, (cudaStream_t)stream, labels, count, nmax, conf_thre, nms_thre, class_agnostic, out,
                                                                  out_count);
  LEOD_LAUNCH_CHECK();
  return 0;
}
