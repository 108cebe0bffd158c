// Sequence post-processing of the pseudo-label sweep on the device:
//   linear-velocity greedy-IoU tracking of the per-frame boxes, forward and backward in time   modules/tracking/linear.py:10-292,
//                                                                                               modules/tracking/utils.py:7-96
//   "ignore" marking of boxes on short tracks, in-painting of missed detections                 modules/pseudo_labeler.py:201-333
//   BBOX_DTYPE records of the label files                                                       data/genx_utils/labels.py:12-16, 312-325
//
// The reference runs this as Python/numpy loops, one sequence at a time, after the sweep.  Here every (sequence, direction) is one
// CTA: thousands of sequences are independent, the work inside one is a short dependent chain per frame (predict -> IoU -> greedy
// association -> update), so one lane walks the frames with the live tracks in shared memory and the machine is filled by
// running ~1 300 sequences side by side.  Results are BIT-EXACT with the reference: box arithmetic is float32 with explicit
// round-to-nearest operations (no FMA contraction), track confidences are doubles as the reference's Python floats, the powers
// q^age come from a host table computed with the host's pow() (what CPython calls), confidence ties resolve in creation order
// (numpy's argsort is an insertion sort, i.e. stable, up to 16 live tracks; beyond that its order is platform dependent).
#include "common.cuh"

namespace {

constexpr int MAXT = 64;     // live tracks per sequence
constexpr int MAXC = 8;      // cached (uncommitted) misses per track: a track survives at most 6 misses in a row (0.9^6 < 0.55)

struct Track {
  float box[4], last[4], pred[5], v[2];
  float cls;
  int uid, age, hits;
  double conf;
  unsigned char clamp[4];     // top, down, left, right
  unsigned char gt, ncache;
  int cache_f[MAXC];
  float cache_box[MAXC][5];
};

__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float clipf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// tracking/utils.py:72-96 ('xywh'): clamp the corners to the image, report which sides were clamped
__device__ void clamp_center_box(const float *b, int H, int W, float *out, unsigned char *flags) {
  const float x1_ = fsub(b[0], fdiv(b[2], 2.f)), y1_ = fsub(b[1], fdiv(b[3], 2.f));
  const float x2_ = fadd(b[0], fdiv(b[2], 2.f)), y2_ = fadd(b[1], fdiv(b[3], 2.f));
  const float x1 = clipf(x1_, 0.f, (float)(W - 1)), x2 = clipf(x2_, 0.f, (float)(W - 1));
  const float y1 = clipf(y1_, 0.f, (float)(H - 1)), y2 = clipf(y2_, 0.f, (float)(H - 1));
  out[0] = fdiv(fadd(x1, x2), 2.f);
  out[1] = fdiv(fadd(y1, y2), 2.f);
  out[2] = fsub(x2, x1);
  out[3] = fsub(y2, y1);
  flags[0] = y1 != y1_; flags[1] = y2 != y2_; flags[2] = x1 != x1_; flags[3] = x2 != x2_;
}
// tracking/utils.py:21-49: IoU of two centre boxes with a class column, 0 between different classes
__device__ float iou_center(const float *t, const float *d) {
  const float xx1 = fmaxf(fsub(t[0], fdiv(t[2], 2.f)), fsub(d[0], fdiv(d[2], 2.f)));
  const float yy1 = fmaxf(fsub(t[1], fdiv(t[3], 2.f)), fsub(d[1], fdiv(d[3], 2.f)));
  const float xx2 = fminf(fadd(t[0], fdiv(t[2], 2.f)), fadd(d[0], fdiv(d[2], 2.f)));
  const float yy2 = fminf(fadd(t[1], fdiv(t[3], 2.f)), fadd(d[1], fdiv(d[3], 2.f)));
  const float wh = fmul(fmaxf(0.f, fsub(xx2, xx1)), fmaxf(0.f, fsub(yy2, yy1)));
  const float o = fdiv(wh, fsub(fadd(fmul(t[2], t[3]), fmul(d[2], d[3])), wh));
  return d[4] != t[4] ? 0.f : o;
}
// ObjectLabels row (t, x, y, w, h, cls, cls_conf, obj) -> centre box with class (data/genx_utils/labels.py:521-531)
__device__ __forceinline__ void row_to_center(const float *r, float *c) {
  c[0] = fadd(r[1], fmul(0.5f, r[3]));
  c[1] = fadd(r[2], fmul(0.5f, r[4]));
  c[2] = r[3]; c[3] = r[4]; c[4] = r[5];
}

struct TrackArgs {
  const float *rows;          // [total_rows, 8]
  const int *frame_ptr;       // [total_frames + 1] first row of every labelled frame
  const int *frame_idx;       // [total_frames]     index of the frame inside its sequence (ascending)
  const int *seq_ptr;         // [S + 1]            first labelled frame of every sequence
  const int *hw;              // [S, 2]
  const double *qpow;         // [npow] q^age, host computed
  int npow;
  double q, min_conf;
  float iou_thr;
  int min_track_len, inpaint;
  // outputs
  unsigned char *remove;      // [2, total_rows]: per direction, 1 = the box sits on a short finished track
  int *hole_count;            // [S] committed in-painting entries (forward direction)
  int *hole_frame;            // [S, hole_cap]
  int *hole_order;            // [S, hole_cap]: finish order of the owning track * 65536 + commit order (sort key inside a frame)
  float *hole_box;            // [S, hole_cap, 5]
  unsigned char *hole_keep;   // [S, hole_cap]
  int hole_cap;
  int *status;                // [S] 0 ok, 1 too many live tracks, 2 in-painting capacity exceeded, 3 q^age table too short
  // per (direction, sequence) scratch: one int / byte per box
  int *box_owner;             // [2, total_rows] uid (= id of the box that created the track) of the track owning each box
  unsigned char *uid_flags;   // [2, total_rows] per uid: bit0 done, bit1 gt, bit2 short (hits < min_track_len)
  int *uid_finish;            // [2, total_rows] per uid: position in the `finished` list
  int total_rows;
};

// LinearTracker over one sequence in one direction (linear.py:196-292, pseudo_labeler.py:211-225).  Backward = frames in reverse
// order with mirrored frame indices and the rows of every frame reversed (pseudo_labeler.py:286-296).
__global__ void __launch_bounds__(32) track_seq_kernel(TrackArgs a) {
  pdl_prologue();
  __shared__ Track tr[MAXT];
  __shared__ int order[MAXT];
  __shared__ float dets[64][5];
  if (threadIdx.x != 0) return;
  const int s = blockIdx.x, dir = blockIdx.y;
  const int f0 = a.seq_ptr[s], f1 = a.seq_ptr[s + 1];
  const int nfr = f1 - f0;
  const int H = a.hw[2 * s], W = a.hw[2 * s + 1];
  const int row0 = a.frame_ptr[f0], nrow = a.frame_ptr[f1] - row0;
  unsigned char *remove = a.remove + (size_t)dir * a.total_rows + row0;
  int *owner = a.box_owner + (size_t)dir * a.total_rows + row0;
  unsigned char *uflags = a.uid_flags + (size_t)dir * a.total_rows + row0;
  int *ufinish = a.uid_finish + (size_t)dir * a.total_rows + row0;
  if (dir == 0) a.hole_count[s] = 0;
  if (nfr == 0) return;
  const int maxf = a.frame_idx[f1 - 1];
  int nlive = 0, n_boxes = 0, n_finished = 0, n_holes = 0, status = 0;
  int *hole_frame = a.hole_frame + (size_t)s * a.hole_cap;
  int *hole_order = a.hole_order + (size_t)s * a.hole_cap;
  float *hole_box = a.hole_box + (size_t)s * a.hole_cap * 5;

  auto retire = [&](int i, bool done) {
    const Track &t = tr[i];
    uflags[t.uid] = (unsigned char)((done ? 1 : 0) | (t.gt ? 2 : 0) | (t.hits < a.min_track_len ? 4 : 0));
    ufinish[t.uid] = n_finished++;
    for (int k = i; k + 1 < nlive; ++k) tr[k] = tr[k + 1];
    --nlive;
  };

  int k = dir == 0 ? 0 : nfr - 1;     // position in the labelled-frame list
  // the mirrored sequence ends at the mirror image of the FIRST labelled frame (pseudo_labeler.py:288: max(frame_idx) - i)
  const int lastf = dir == 0 ? maxf : maxf - a.frame_idx[f0];
  for (int f = 0; f <= lastf; ++f) {
    // detections of frame f (sequence-local index; mirrored for the backward pass)
    int nd = 0, base = 0;
    bool any_gt = false;
    const bool have = dir == 0 ? (k < nfr && a.frame_idx[f0 + k] == f) : (k >= 0 && maxf - a.frame_idx[f0 + k] == f);
    if (have) {
      const int r0 = a.frame_ptr[f0 + k], r1 = a.frame_ptr[f0 + k + 1];
      nd = min(r1 - r0, 64);
      if (r1 - r0 > 64) status = 1;
      for (int j = 0; j < nd; ++j) {
        const int rj = dir == 0 ? r0 + j : r1 - 1 - j;
        row_to_center(a.rows + (size_t)rj * 8, dets[j]);
        any_gt |= a.rows[(size_t)rj * 8] > 0.f;
      }
      base = dir == 0 ? r0 - row0 : nrow - (r1 - row0);     // id of the frame's first box in this direction's numbering
      k += dir == 0 ? 1 : -1;
    }
    if (nd == 0 && nlive == 0) continue;
    // retire degenerate tracks, predict the others (linear.py:71-83)
    for (int i = nlive - 1; i >= 0; --i)
      if (!(fmul(tr[i].box[2], tr[i].box[3]) > 0.f)) retire(i, true);
    for (int i = 0; i < nlive; ++i) {
      Track &t = tr[i];
      t.age += 1;
      for (int c = 0; c < 4; ++c) t.last[c] = t.box[c];
      t.box[0] = fadd(t.box[0], t.v[0]);
      t.box[1] = fadd(t.box[1], t.v[1]);
      clamp_center_box(t.box, H, W, t.pred, t.clamp);
      t.pred[4] = t.cls;
    }
    // association: tracks by descending confidence (stable), each takes its best remaining detection if IoU >= thr (utils.py:7-18)
    for (int i = 0; i < nlive; ++i) order[i] = i;
    for (int i = 1; i < nlive; ++i) {
      const int oi = order[i];
      int j = i - 1;
      while (j >= 0 && tr[order[j]].conf < tr[oi].conf) { order[j + 1] = order[j]; --j; }
      order[j + 1] = oi;
    }
    unsigned long long taken = 0ull, hit_t = 0ull;
    int match_of[MAXT];
    bool any_pos = false;
    for (int i = 0; i < nlive && !any_pos; ++i)
      for (int j = 0; j < nd; ++j)
        if (iou_center(tr[i].pred, dets[j]) > 0.f) { any_pos = true; break; }
    if (any_pos) {
      for (int oi = 0; oi < nlive; ++oi) {
        const int i = order[oi];
        float best = -1.f;
        int bj = -1;
        for (int j = 0; j < nd; ++j) {
          if (taken >> j & 1ull) continue;
          const float o = iou_center(tr[i].pred, dets[j]);
          if (o > best) { best = o; bj = j; }
        }
        if (bj < 0 || best < a.iou_thr) continue;
        taken |= 1ull << bj;
        hit_t |= 1ull << i;
        match_of[i] = bj;
      }
    }
    // update matched tracks in association order (linear.py:85-124)
    for (int oi = 0; oi < nlive; ++oi) {
      const int i = order[oi];
      if (!(hit_t >> i & 1ull)) continue;
      Track &t = tr[i];
      const float *d = dets[match_of[i]];
      t.hits = t.age + 1;
      float v0 = fsub(d[0], t.last[0]), v1 = fsub(d[1], t.last[1]);
      if (t.clamp[0] | t.clamp[1] | t.clamp[2] | t.clamp[3]) {    // robust velocity: the displacement of the free edge
        const float ox1 = fsub(t.last[0], fdiv(t.last[2], 2.f)), oy1 = fsub(t.last[1], fdiv(t.last[3], 2.f));
        const float ox2 = fadd(t.last[0], fdiv(t.last[2], 2.f)), oy2 = fadd(t.last[1], fdiv(t.last[3], 2.f));
        const float nx1 = fsub(d[0], fdiv(d[2], 2.f)), ny1 = fsub(d[1], fdiv(d[3], 2.f));
        const float nx2 = fadd(d[0], fdiv(d[2], 2.f)), ny2 = fadd(d[1], fdiv(d[3], 2.f));
        if (t.clamp[0]) v1 = fsub(ny2, oy2);
        if (t.clamp[1]) v1 = fsub(ny1, oy1);
        if (t.clamp[2]) v0 = fsub(nx2, ox2);
        if (t.clamp[3]) v0 = fsub(nx1, ox1);
      }
      t.v[0] = v0; t.v[1] = v1;
      for (int c = 0; c < 4; ++c) t.box[c] = d[c];
      const int bid = base + match_of[i];
      owner[bid] = t.uid;
      const int rj = dir == 0 ? row0 + bid : row0 + nrow - 1 - bid;
      t.gt |= a.rows[(size_t)rj * 8] > 0.f;
      if (t.age >= a.npow) { status = 3; }
      const double qa = a.qpow[min(t.age, a.npow - 1)];
      const double w = __ddiv_rn(__dmul_rn(a.q, __dsub_rn(1.0, qa)), __dsub_rn(1.0, a.q));
      t.conf = __ddiv_rn(__dadd_rn(__dmul_rn(w, t.conf), 1.0), __dadd_rn(w, 1.0));
      if (dir == 0 && a.inpaint) {     // commit the cached misses (linear.py:104-105)
        for (int c = 0; c < t.ncache; ++c) {
          if (n_holes >= a.hole_cap) { status = 2; break; }
          hole_frame[n_holes] = t.cache_f[c];
          hole_order[n_holes] = t.uid;     // replaced by the finish order at the end
          for (int e = 0; e < 5; ++e) hole_box[n_holes * 5 + e] = t.cache_box[c][e];
          ++n_holes;
        }
      }
      t.ncache = 0;
    }
    // misses (linear.py:126-134), new tracks (linear.py:254-259), retirement (linear.py:261-267)
    for (int i = 0; i < nlive; ++i) {
      if (hit_t >> i & 1ull) continue;
      Track &t = tr[i];
      t.conf = __dmul_rn(t.conf, a.q);
      if (!any_gt && t.ncache < MAXC) {
        t.cache_f[t.ncache] = f;
        for (int e = 0; e < 5; ++e) t.cache_box[t.ncache][e] = t.pred[e];
        ++t.ncache;
      }
    }
    for (int j = 0; j < nd; ++j) {
      if (taken >> j & 1ull) continue;
      if (nlive >= MAXT) { status = 1; break; }
      Track &t = tr[nlive++];
      for (int c = 0; c < 4; ++c) { t.box[c] = dets[j][c]; t.last[c] = 0.f; }
      t.cls = dets[j][4];
      t.v[0] = t.v[1] = 0.f;
      t.clamp[0] = t.clamp[1] = t.clamp[2] = t.clamp[3] = 0;
      t.uid = base + j;
      const int rj = dir == 0 ? row0 + t.uid : row0 + nrow - 1 - t.uid;
      t.gt = a.rows[(size_t)rj * 8] > 0.f;
      t.conf = a.q; t.age = 0; t.hits = 1; t.ncache = 0;
      owner[t.uid] = t.uid;
    }
    for (int i = nlive - 1; i >= 0; --i)
      if (tr[i].conf < a.min_conf) retire(i, true);
    n_boxes += nd;
  }
  for (int i = nlive - 1; i >= 0; --i) retire(i, false);     // tracker.py:34-39: unfinished tracks are never filtered
  // pseudo_labeler.py:227-238: a box is ignored when its track finished, carries no ground truth and is shorter than min_track_len
  for (int b = 0; b < nrow; ++b) {
    const unsigned char fl = uflags[owner[b]];
    const bool rm = (fl & 1) && !(fl & 2) && (fl & 4);
    remove[dir == 0 ? b : nrow - 1 - b] = rm ? 1 : 0;
  }
  if (dir == 0) {
    for (int h = 0; h < n_holes; ++h) {       // pseudo_labeler.py:243-252: in-paint from the tracks that are kept, in `finished` order
      const int uid = hole_order[h];
      const unsigned char fl = uflags[uid];
      a.hole_keep[(size_t)s * a.hole_cap + h] = ((fl & 1) && !(fl & 2) && (fl & 4)) ? 0 : 1;
      hole_order[h] = ufinish[uid] * 65536 + h;
    }
    a.hole_count[s] = n_holes;
  }
  a.status[s] = max(a.status[s], status);
}

struct AssembleArgs {
  const float *rows;
  const int *frame_ptr, *frame_idx, *seq_ptr;
  const unsigned char *remove;     // [2, total_rows]
  int total_rows, use_backward;
  const int *hole_count, *hole_frame, *hole_order;
  const float *hole_box;
  const unsigned char *hole_keep;
  int hole_cap;
  float ignore_label;
  // outputs, per sequence s: rows at out_rows + out_row_ptr[s] * 8, frames at out_frame_* + out_frame_base[s]
  float *out_rows;
  const int *out_row_base;      // [S] capacity offsets (host: rows of the sequence + hole_cap)
  int *out_frame_idx, *out_frame_start;   // [.., frames of the sequence + hole_cap]
  const int *out_frame_base;
  int *out_counts;              // [S, 2]: rows, frames
};
// EventSeqData._track_filter (pseudo_labeler.py:298-333): ignore labels, in-painted boxes appended to their frame (new frames where the
// detector had nothing), frames in ascending order.
__global__ void __launch_bounds__(32) track_assemble_kernel(AssembleArgs a) {
  pdl_prologue();
  if (threadIdx.x != 0) return;
  const int s = blockIdx.x;
  const int f0 = a.seq_ptr[s], f1 = a.seq_ptr[s + 1];
  float *out = a.out_rows + (size_t)a.out_row_base[s] * 8;
  int *ofi = a.out_frame_idx + a.out_frame_base[s], *ofs = a.out_frame_start + a.out_frame_base[s];
  const int nh = a.hole_count[s];
  const int *hf = a.hole_frame + (size_t)s * a.hole_cap, *ho = a.hole_order + (size_t)s * a.hole_cap;
  const float *hb = a.hole_box + (size_t)s * a.hole_cap * 5;
  const unsigned char *hk = a.hole_keep + (size_t)s * a.hole_cap;
  int nr = 0, nf = 0;
  if (f1 > f0) {
    const int maxf = a.frame_idx[f1 - 1];
    int k = f0;
    for (int f = 0; f <= maxf; ++f) {
      const bool have = k < f1 && a.frame_idx[k] == f;
      int nholes_f = 0;
      for (int h = 0; h < nh; ++h) nholes_f += (hk[h] && hf[h] == f) ? 1 : 0;
      if (!have && nholes_f == 0) continue;
      ofi[nf] = f;
      ofs[nf] = nr;
      ++nf;
      if (have) {
        for (int r = a.frame_ptr[k]; r < a.frame_ptr[k + 1]; ++r) {
          for (int e = 0; e < 8; ++e) out[(size_t)nr * 8 + e] = a.rows[(size_t)r * 8 + e];
          const bool rm = a.remove[r] && (!a.use_backward || a.remove[(size_t)a.total_rows + r]);
          if (rm) out[(size_t)nr * 8 + 5] = a.ignore_label;
          ++nr;
        }
        ++k;
      }
      int last_key = -1;
      for (int c = 0; c < nholes_f; ++c) {       // selection by ascending (finish order, commit order)
        int best = -1;
        for (int h = 0; h < nh; ++h)
          if (hk[h] && hf[h] == f && ho[h] > last_key && (best < 0 || ho[h] < ho[best])) best = h;
        last_key = ho[best];
        const float *b = hb + best * 5;
        float *o = out + (size_t)nr * 8;
        o[0] = 0.f;
        o[1] = fsub(b[0], fdiv(b[2], 2.f));
        o[2] = fsub(b[1], fdiv(b[3], 2.f));
        o[3] = b[2]; o[4] = b[3];
        o[5] = a.ignore_label;
        o[6] = 0.f; o[7] = 0.f;
        ++nr;
      }
    }
  }
  a.out_counts[2 * s] = nr;
  a.out_counts[2 * s + 1] = nf;
}

// ObjectLabels rows -> BBOX_DTYPE records (labels.py:12-16, 312-325).  stride 40: the declared dtype (t i8 @0, x y w h f4 @8..20,
// class_id u4 @24, class_confidence f4 @28, objectness f4 @32, 4 bytes padding); stride 36: the packed form numpy's concatenate
// produces in EventSeqData._summarize (pseudo_labeler.py:179-199).
__global__ void pack_bbox_kernel(const float *__restrict__ rows, int64_t n, unsigned char *__restrict__ out, int stride) {
  pdl_prologue();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float *r = rows + i * 8;
  unsigned char *o = out + i * stride;
  const long long t = (long long)r[0];
  const unsigned cls = (unsigned)r[5];
  const float f[4] = {r[1], r[2], r[3], r[4]};
  memcpy(o, &t, 8);
  memcpy(o + 8, f, 16);
  memcpy(o + 24, &cls, 4);
  memcpy(o + 28, &r[6], 4);
  memcpy(o + 32, &r[7], 4);
  if (stride == 40) memset(o + 36, 0, 4);
}

}  // namespace

// rows [total_rows, 8] ObjectLabels rows of all labelled frames of all S sequences; frame_ptr [F + 1]; frame_idx [F]; seq_ptr [S + 1];
// hw [S, 2]; qpow: HOST array of q^age (age = 0 .. npow-1).  Workspace `ws`: leod_track_workspace_bytes.
extern "C" int64_t leod_track_workspace_bytes(int64_t total_rows, int S, int hole_cap, int npow) {
  const int64_t rows = round_up(total_rows, 16);
  return 2 * rows * (4 + 1 + 4) + (int64_t)S * hole_cap * (4 + 4 + 20 + 1) + (int64_t)S * 8 + round_up((int64_t)npow * 8, 16) + 2 * rows + 1024;
}

extern "C" int leod_track_filter(const float *rows, const int32_t *frame_ptr, const int32_t *frame_idx, const int32_t *seq_ptr, const int32_t *hw,
                                 int64_t total_rows, int total_frames, int S, const double *qpow_host, int npow, double q, double min_conf,
                                 float iou_thr, int min_track_len, int inpaint, int use_backward, float ignore_label, int hole_cap, void *ws,
                                 const int32_t *out_row_base, const int32_t *out_frame_base, float *out_rows, int32_t *out_frame_idx,
                                 int32_t *out_frame_start, int32_t *out_counts, int32_t *status, void *stream) {
  LEOD_REQUIRE(rows && frame_ptr && frame_idx && seq_ptr && hw && ws && out_rows && out_counts && status && qpow_host, "leod_track_filter: null operand");
  LEOD_REQUIRE(S > 0 && hole_cap > 0 && npow > 1 && total_frames >= 0, "leod_track_filter: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t R = round_up(total_rows, 16);
  char *p = (char *)ws;
  auto take = [&](int64_t bytes) { char *r = p; p += round_up(bytes, 16); return r; };
  TrackArgs a;
  a.rows = rows; a.frame_ptr = frame_ptr; a.frame_idx = frame_idx; a.seq_ptr = seq_ptr; a.hw = hw;
  a.npow = npow; a.q = q; a.min_conf = min_conf; a.iou_thr = iou_thr; a.min_track_len = min_track_len; a.inpaint = inpaint;
  a.total_rows = (int)total_rows;
  a.remove = (unsigned char *)take(2 * R);
  a.box_owner = (int *)take(2 * R * 4);
  a.uid_flags = (unsigned char *)take(2 * R);
  a.uid_finish = (int *)take(2 * R * 4);
  a.hole_count = (int *)take((int64_t)S * 4);
  a.hole_frame = (int *)take((int64_t)S * hole_cap * 4);
  a.hole_order = (int *)take((int64_t)S * hole_cap * 4);
  a.hole_box = (float *)take((int64_t)S * hole_cap * 20);
  a.hole_keep = (unsigned char *)take((int64_t)S * hole_cap);
  double *qp = (double *)take((int64_t)npow * 8);
  a.hole_cap = hole_cap;
  a.status = status;
  a.qpow = qp;
  LEOD_CUDA(cudaMemcpyAsync(qp, qpow_host, (size_t)npow * 8, cudaMemcpyHostToDevice, st));
  LEOD_CUDA(cudaMemsetAsync(status, 0, (size_t)S * 4, st));
  LEOD_LAUNCH((track_seq_kernel), dim3(S, use_backward ? 2 : 1), 32, 0, st, a);
  LEOD_LAUNCH_CHECK();
  AssembleArgs b;
  b.rows = rows; b.frame_ptr = frame_ptr; b.frame_idx = frame_idx; b.seq_ptr = seq_ptr; b.remove = a.remove; b.total_rows = (int)total_rows;
  b.use_backward = use_backward;
  b.hole_count = a.hole_count; b.hole_frame = a.hole_frame; b.hole_order = a.hole_order; b.hole_box = a.hole_box; b.hole_keep = a.hole_keep;
  b.hole_cap = hole_cap; b.ignore_label = ignore_label;
  b.out_rows = out_rows; b.out_row_base = out_row_base; b.out_frame_idx = out_frame_idx; b.out_frame_start = out_frame_start;
  b.out_frame_base = out_frame_base; b.out_counts = out_counts;
  LEOD_LAUNCH((track_assemble_kernel), S, 32, 0, st, b);
  LEOD_LAUNCH_CHECK();
  return 0;
}

extern "C" int leod_pack_bbox(const float *rows, int64_t n, void *out, int stride, void *stream) {
  LEOD_REQUIRE(stride == 36 || stride == 40, "leod_pack_bbox: record stride %d (36 = packed, 40 = BBOX_DTYPE)", stride);
  if (n == 0) return 0;
  LEOD_REQUIRE(rows && out, "leod_pack_bbox: null operand");
  LEOD_LAUNCH((pack_bbox_kernel), ceil_div(n, 256), 256, 0, (cudaStream_t)stream, rows, n, (unsigned char *)out, stride);
  LEOD_LAUNCH_CHECK();
  return 0;
}
