// Host-side executor of the YOLOX PAFPN neck + decoupled head + SimOTA loss (forward and backward).
// Reference semantics: models/detection/yolox_extension/models/yolo_pafpn.py:109-140 (data flow),
// models/detection/yolox/models/network_blocks.py:29-54, 79-142 (Conv-BN-SiLU, Bottleneck, CSPLayer),
// models/detection/yolox/models/yolo_head.py:195-287 (head), :403-1148 (loss), train.py:247 (SyncBatchNorm).
//
// Every Conv-BN-SiLU is:  implicit-GEMM convolution on the padded-flat layout (kernels_conv.cu; tcgen05 in bf16, SIMT in
// fp32)  ->  per-channel batch statistics  ->  BN-apply + SiLU written into the consumer's buffer (channel-concatenations
// are column slices of one matrix, never copies).  Convolutions that read the same input are one GEMM (CSP conv1|conv2,
// the first convolutions of the class / regression towers), the three prediction convolutions of a level are one GEMM
// over the concatenated tower outputs with a block-diagonal weight.  Backward mirrors it: BN+SiLU backward as a
// reduction + apply pair, weight gradient as a TN GEMM over shifted rows, input gradient as the transposed tapped GEMM.
// Layers of the three pyramid levels that have no dependency on each other form one PHASE: their batch statistics are
// contiguous in memory, so data-parallel training exchanges them with ONE collective per phase (the callback set with
// leod_detect_set_allreduce) instead of two per layer.
#include <algorithm>
#include <string>
#include <vector>

#include "common.cuh"

namespace {

struct PEntry {
  std::string name;
  int64_t offset;
  int ndim;
  int64_t shape[4];
  int64_t numel() const {
    int64_t n = 1;
    for (int i = 0; i < ndim; ++i) n *= shape[i];
    return n;
  }
};

struct Act {      // one padded-flat activation matrix [R_level, width] and its gradient
  int level, width;
  void *z = nullptr, *dz = nullptr;
};

struct Conv {
  std::string nm[2];
  int nseg = 1;
  int k = 1, stride = 1, cin = 0, cout = 0, cseg = 0;
  int lev_in = 0, lev_out = 0;
  int in_act = -1, in_off = 0, out_act = -1, out_off = 0;
  int rot = 0;             // input-channel rotation of the prepared weight (CSP conv3 reads [x_2 | m_out])
  int cinp = 0, coutp = 0;
  int64_t w[2], bnw[2], bnb[2], rm[2], rv[2];
  int nbt[2];
  bool use_G = false;      // weight gradient goes through a prepared-layout scratch (3x3 / rotated), else straight into grads
  void *Wp = nullptr, *WpT = nullptr;
  float *G = nullptr;
  void *Y = nullptr, *dY = nullptr;   // raw convolution output / its gradient (both kept per layer: the weight-gradient GEMM of a
                                      // layer runs on a side stream while the main stream already works on the next layer)
  cudaEvent_t ev_dy = nullptr;
  double *stats = nullptr, *dloc = nullptr, *dglob = nullptr;
  float *fin = nullptr, *dfin = nullptr;   // mean | rstd  and  mean(dyhat) | mean(dyhat * xhat), see kernels_conv.cu
};

struct Pred {     // the three 1x1 prediction convolutions of one level (yolo_head.py:214-222) as one GEMM
  int64_t cls_w, cls_b, reg_w, reg_b, obj_w, obj_b;
  void *B = nullptr, *BT = nullptr;   // [16, 2*hid], [2*hid, 8]
  float *bias = nullptr, *G = nullptr, *gb = nullptr;   // [8], [8, 2*hid], [8]
  int in_act = -1;
  float *raw = nullptr;    // fp32 [R, 8]
  void *draw = nullptr;    // T    [R, 8]
};

struct Op {
  int kind;                 // 0: phase of convolutions, 1: nearest x2 upsample
  std::vector<int> convs;
  int src_act = -1, src_off = 0, dst_act = -1, dst_off = 0, C = 0;
};

struct PrepDesc {   // one weight tensor of the batched prepare / gradient un-prepare kernels
  int64_t src_off;
  void *dst, *dstT;
  float *G;
  int ldd, lddT, coutp, row0, Cout, Cin, k, cinp, rot, tmode;
};

}  // namespace

struct leod_detect {
  leod_detect_cfg cfg;
  int hid = 0, nb = 1;
  int lh[3], lw[3];
  std::vector<PEntry> pentries, bentries, nentries;   // parameters, fp32 buffers (running stats), int64 buffers (num_batches_tracked)
  int64_t n_params = 0, n_buffers = 0, n_nbt = 0;
  std::vector<Act> acts;
  std::vector<Conv> convs;
  std::vector<Op> ops;
  Pred pred[3];
  int x_act[3], x_off[3];        // where the three backbone features are gathered to (x2, x1, x0)
  int t2_act[3];
  int head_convs[3][4];          // per level: stem, fused first tower convolutions, cls tower 2nd, reg tower 2nd
  size_t n_neck_ops = 0;         // ops[0 .. n_neck_ops) = neck, the rest = head phases (used when statistics are exchanged)
  cudaStream_t lvl_stream[2] = {nullptr, nullptr};   // head levels 1, 2 (level 0 runs on the caller's stream)
  cudaStream_t wg_stream[2] = {nullptr, nullptr};    // weight-gradient GEMMs
  cudaEvent_t ev_fork = nullptr, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};
  int wg_next = 0;
  float *params = nullptr, *grads = nullptr, *buffers = nullptr;
  long long *nbt = nullptr;
  bool layout_only = false;
  std::vector<void *> owned;
  PrepDesc *prep_dev = nullptr;
  int n_prep = 0;
  int max_prep_items = 0;
  // arena (sized for arena_B images)
  int arena_B = 0, cur_B = 0;
  void *arena = nullptr;
  void *col = nullptr, *dcol = nullptr;
  double *stats_all = nullptr, *dloc_all = nullptr, *dglob_all = nullptr;
  int64_t stats_doubles = 0, dstat_doubles = 0;
  float *tout = nullptr, *miou = nullptr, *losses = nullptr;
  uint8_t *flags = nullptr;
  int *match_cnt = nullptr, *match_gt = nullptr, *assign = nullptr;
  double *sums = nullptr;
  leod_allreduce_fn allreduce = nullptr;
  void *allreduce_ctx = nullptr;
  unsigned long long fwd_gen = 0, train_gen = 0;
  bool loss_ready = false, dloss_ready = false;
  size_t esz() const { return cfg.dtype == LEOD_BF16 ? 2 : 4; }
};

namespace {

int64_t add_entry(std::vector<PEntry> &v, int64_t *total, const std::string &name, std::initializer_list<int64_t> shape, int64_t align = 4) {
  PEntry e;
  e.name = name;
  e.ndim = (int)shape.size();
  int i = 0;
  for (auto s : shape) e.shape[i++] = s;
  for (; i < 4; ++i) e.shape[i] = 1;
  e.offset = *total;
  *total += round_up(e.numel(), align);
  v.push_back(e);
  return e.offset;
}

int add_act(leod_detect *h, int level, int width) {
  Act a;
  a.level = level;
  a.width = width;
  h->acts.push_back(a);
  return (int)h->acts.size() - 1;
}

// Conv-BN-SiLU (or a fused twin pair) -> index into h->convs
int add_conv(leod_detect *h, const std::string &n0, const std::string &n1, int k, int stride, int cin, int cout_total, int lev_in, int in_act,
             int in_off, int out_act, int out_off, int rot = 0) {
  Conv c;
  c.nm[0] = n0; c.nm[1] = n1;
  c.nseg = n1.empty() ? 1 : 2;
  c.k = k; c.stride = stride; c.cin = cin; c.cout = cout_total; c.cseg = cout_total / c.nseg;
  c.lev_in = lev_in; c.lev_out = stride == 2 ? lev_in + 1 : lev_in;
  c.in_act = in_act; c.in_off = in_off; c.out_act = out_act; c.out_off = out_off;
  c.rot = rot;
  c.cinp = (int)round_up(cin, 64);
  c.coutp = (int)round_up(cout_total, 64);
  c.use_G = (k == 3) || rot != 0;
  for (int s = 0; s < c.nseg; ++s) c.w[s] = add_entry(h->pentries, &h->n_params, c.nm[s] + ".conv.weight", {c.cseg, cin, k, k});
  for (int s = 0; s < c.nseg; ++s) {
    c.bnw[s] = add_entry(h->pentries, &h->n_params, c.nm[s] + ".bn.weight", {c.cseg});
    c.bnb[s] = add_entry(h->pentries, &h->n_params, c.nm[s] + ".bn.bias", {c.cseg});
    c.rm[s] = add_entry(h->bentries, &h->n_buffers, c.nm[s] + ".bn.running_mean", {c.cseg});
    c.rv[s] = add_entry(h->bentries, &h->n_buffers, c.nm[s] + ".bn.running_var", {c.cseg});
    c.nbt[s] = (int)add_entry(h->nentries, &h->n_nbt, c.nm[s] + ".bn.num_batches_tracked", {1}, 1);
  }
  h->convs.push_back(c);
  return (int)h->convs.size() - 1;
}

void add_phase(leod_detect *h, std::initializer_list<int> convs) {
  Op o;
  o.kind = 0;
  o.convs.assign(convs.begin(), convs.end());
  h->ops.push_back(o);
}
void add_upsample(leod_detect *h, int src_act, int src_off, int dst_act, int dst_off, int C) {
  Op o;
  o.kind = 1;
  o.src_act = src_act; o.src_off = src_off; o.dst_act = dst_act; o.dst_off = dst_off; o.C = C;
  h->ops.push_back(o);
}

// CSPLayer with shortcut=False (network_blocks.py:104-142 as built by yolo_pafpn.py:51-97).
// W = [x_1 | x_2 | m_out]: conv1|conv2 fused write columns [0, 2h), the last bottleneck writes [2h, 3h), conv3 reads [h, 3h).
void add_csp(leod_detect *h, const std::string &p, int level, int in_act, int cin, int out_act, int out_off, int cout) {
  const int hd = cout / 2;
  const int W = add_act(h, level, 3 * hd);
  add_phase(h, {add_conv(h, p + ".conv1", p + ".conv2", 1, 1, cin, 2 * hd, level, in_act, 0, W, 0)});
  int src = W, src_off = 0;
  for (int i = 0; i < h->nb; ++i) {
    const std::string m = p + ".m." + std::to_string(i);
    const int mt = add_act(h, level, hd);
    add_phase(h, {add_conv(h, m + ".conv1", "", 1, 1, hd, hd, level, src, src_off, mt, 0)});
    const bool last = i == h->nb - 1;
    const int mo = last ? W : add_act(h, level, hd);
    add_phase(h, {add_conv(h, m + ".conv2", "", 3, 1, hd, hd, level, mt, 0, mo, last ? 2 * hd : 0)});
    src = mo;
    src_off = 0;
  }
  add_phase(h, {add_conv(h, p + ".conv3", "", 1, 1, 2 * hd, cout, level, W, hd, out_act, out_off, /*rot=*/hd)});
}

int build_graph(leod_detect *h) {
  const leod_detect_cfg &c = h->cfg;
  const int c0 = c.in_channels[0], c1 = c.in_channels[1], c2 = c.in_channels[2];
  LEOD_REQUIRE(c0 % 16 == 0 && c1 % 16 == 0 && c2 % 16 == 0, "detect: stage dims %d/%d/%d must be multiples of 16", c0, c1, c2);
  LEOD_REQUIRE(c.num_classes >= 1 && c.num_classes <= 3, "detect: num_classes %d (1..3 supported)", c.num_classes);
  LEOD_REQUIRE(c.n_bottleneck >= 1 && c.n_bottleneck <= 4, "detect: %d bottlenecks per CSP layer", c.n_bottleneck);
  h->nb = c.n_bottleneck;
  h->hid = (int)(256 * (int64_t)c2 / 1024);   // yolo_head.py:61-66
  LEOD_REQUIRE(h->hid % 16 == 0, "detect: head width %d must be a multiple of 16", h->hid);
  for (int l = 0; l < 3; ++l) {
    LEOD_REQUIRE(c.in_h % c.strides[l] == 0 && c.in_w % c.strides[l] == 0, "detect: input %dx%d not divisible by stride %d", c.in_h, c.in_w, c.strides[l]);
    h->lh[l] = c.in_h / c.strides[l];
    h->lw[l] = c.in_w / c.strides[l];
    if (l > 0) LEOD_REQUIRE(h->lh[l - 1] == 2 * h->lh[l] && h->lw[l - 1] == 2 * h->lw[l], "detect: pyramid levels must halve");
  }
  // --- neck (yolo_pafpn.py:109-140).  Levels: 0 = stride 8 (x2), 1 = stride 16 (x1), 2 = stride 32 (x0)
  const int X0 = add_act(h, 2, c2);
  const int cat_n4 = add_act(h, 2, 2 * c1);   // [bu_conv1 out | fpn_out0]
  const int cat_p4 = add_act(h, 1, 2 * c1);   // [up(fpn_out0) | x1]
  const int f_out0 = add_act(h, 1, c1);
  const int cat_n3 = add_act(h, 1, 2 * c0);   // [bu_conv2 out | fpn_out1]
  const int cat_p3 = add_act(h, 0, 2 * c0);   // [up(fpn_out1) | x2]
  const int pan2 = add_act(h, 0, c0), pan1 = add_act(h, 1, c1), pan0 = add_act(h, 2, c2);
  h->x_act[0] = cat_p3; h->x_off[0] = c0;
  h->x_act[1] = cat_p4; h->x_off[1] = c1;
  h->x_act[2] = X0;     h->x_off[2] = 0;
  add_phase(h, {add_conv(h, "fpn.lateral_conv0", "", 1, 1, c2, c1, 2, X0, 0, cat_n4, c1)});
  add_upsample(h, cat_n4, c1, cat_p4, 0, c1);
  add_csp(h, "fpn.C3_p4", 1, cat_p4, 2 * c1, f_out0, 0, c1);
  add_phase(h, {add_conv(h, "fpn.reduce_conv1", "", 1, 1, c1, c0, 1, f_out0, 0, cat_n3, c0)});
  add_upsample(h, cat_n3, c0, cat_p3, 0, c0);
  add_csp(h, "fpn.C3_p3", 0, cat_p3, 2 * c0, pan2, 0, c0);
  add_phase(h, {add_conv(h, "fpn.bu_conv2", "", 3, 2, c0, c0, 0, pan2, 0, cat_n3, 0)});
  add_csp(h, "fpn.C3_n3", 1, cat_n3, 2 * c0, pan1, 0, c1);
  add_phase(h, {add_conv(h, "fpn.bu_conv1", "", 3, 2, c1, c1, 1, pan1, 0, cat_n4, 0)});
  add_csp(h, "fpn.C3_n4", 2, cat_n4, 2 * c1, pan0, 0, c2);
  // --- head (yolo_head.py:195-222), phase by phase over the three levels
  const int hid = h->hid;
  const int pan[3] = {pan2, pan1, pan0}, cin[3] = {c0, c1, c2};
  int stemz[3], t1[3], s_conv[3], t1_conv[3], cls2[3], reg2[3];
  for (int l = 0; l < 3; ++l) {
    const std::string L = std::to_string(l);
    stemz[l] = add_act(h, l, hid);
    t1[l] = add_act(h, l, 2 * hid);          // [cls tower | reg tower] after their first convolution
    h->t2_act[l] = add_act(h, l, 2 * hid);   // ... after their second convolution: input of the prediction GEMM
    s_conv[l] = add_conv(h, "yolox_head.stems." + L, "", 1, 1, cin[l], hid, l, pan[l], 0, stemz[l], 0);
    t1_conv[l] = add_conv(h, "yolox_head.cls_convs." + L + ".0", "yolox_head.reg_convs." + L + ".0", 3, 1, hid, 2 * hid, l, stemz[l], 0, t1[l], 0);
    cls2[l] = add_conv(h, "yolox_head.cls_convs." + L + ".1", "", 3, 1, hid, hid, l, t1[l], 0, h->t2_act[l], 0);
    reg2[l] = add_conv(h, "yolox_head.reg_convs." + L + ".1", "", 3, 1, hid, hid, l, t1[l], hid, h->t2_act[l], hid);
  }
  for (int l = 0; l < 3; ++l) {
    h->head_convs[l][0] = s_conv[l]; h->head_convs[l][1] = t1_conv[l]; h->head_convs[l][2] = cls2[l]; h->head_convs[l][3] = reg2[l];
  }
  h->n_neck_ops = h->ops.size();
  add_phase(h, {s_conv[0], s_conv[1], s_conv[2]});
  add_phase(h, {t1_conv[0], t1_conv[1], t1_conv[2]});
  add_phase(h, {cls2[0], reg2[0], cls2[1], reg2[1], cls2[2], reg2[2]});
  const int C = c.num_classes;
  for (int l = 0; l < 3; ++l) {
    const std::string L = std::to_string(l);
    Pred &p = h->pred[l];
    p.in_act = h->t2_act[l];
    p.cls_w = add_entry(h->pentries, &h->n_params, "yolox_head.cls_preds." + L + ".weight", {C, hid, 1, 1});
    p.cls_b = add_entry(h->pentries, &h->n_params, "yolox_head.cls_preds." + L + ".bias", {C});
    p.reg_w = add_entry(h->pentries, &h->n_params, "yolox_head.reg_preds." + L + ".weight", {4, hid, 1, 1});
    p.reg_b = add_entry(h->pentries, &h->n_params, "yolox_head.reg_preds." + L + ".bias", {4});
    p.obj_w = add_entry(h->pentries, &h->n_params, "yolox_head.obj_preds." + L + ".weight", {1, hid, 1, 1});
    p.obj_b = add_entry(h->pentries, &h->n_params, "yolox_head.obj_preds." + L + ".bias", {1});
  }
  return 0;
}

int dev_alloc(leod_detect *h, void **p, size_t bytes) {
  LEOD_CUDA(cudaMalloc(p, bytes ? bytes : 16));
  LEOD_CUDA(cudaMemset(*p, 0, bytes ? bytes : 16));
  h->owned.push_back(*p);
  return 0;
}

// ------------------------------------------------------------------ weight preparation (batched over all layers)
// src fp32 [Cout, Cin, k, k] -> dst[(row0 + n), tap*cinp + ci'] with ci = (ci' + rot) % Cin
//   tmode 0: dstT[ci', tap*coutp + row0 + n]   (B operand [Cin, taps*coutp] of the tapped / plain input-gradient GEMM)
//   tmode 1: dstT[tap*cinp + ci', row0 + n]    (B operand [taps*cinp, coutp] of the stride-2 input gradient through the patch matrix)
template <typename T>
__global__ void prep_batched_kernel(const PrepDesc *__restrict__ descs, const float *__restrict__ params) {
  pdl_prologue();
  const PrepDesc d = descs[blockIdx.y];
  const int taps = d.k * d.k;
  const int64_t total = (int64_t)d.Cout * taps * d.Cin;
  const float *src = params + d.src_off;
  T *dst = (T *)d.dst, *dstT = (T *)d.dstT;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cip = (int)(i % d.Cin);
    const int tap = (int)((i / d.Cin) % taps);
    const int n = (int)(i / ((int64_t)d.Cin * taps));
    const int ci = (cip + d.rot) % d.Cin;
    const T v = from_f<T>(src[((int64_t)n * d.Cin + ci) * taps + tap]);
    dst[(int64_t)(d.row0 + n) * d.ldd + tap * d.cinp + cip] = v;
    if (d.tmode == 0)
      dstT[(int64_t)cip * d.lddT + tap * d.coutp + d.row0 + n] = v;
    else
      dstT[(int64_t)(tap * d.cinp + cip) * d.lddT + d.row0 + n] = v;
  }
}
// G fp32 (prepared layout) -> dW [Cout, Cin, k, k] += ; G cleared
__global__ void unprep_batched_kernel(const PrepDesc *__restrict__ descs, float *__restrict__ grads) {
  pdl_prologue();
  const PrepDesc d = descs[blockIdx.y];
  if (!d.G) return;
  const int taps = d.k * d.k;
  const int64_t total = (int64_t)d.Cout * taps * d.Cin;
  float *dW = grads + d.src_off;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cip = (int)(i % d.Cin);
    const int tap = (int)((i / d.Cin) % taps);
    const int n = (int)(i / ((int64_t)d.Cin * taps));
    const int ci = (cip + d.rot) % d.Cin;
    float *gp = d.G + (int64_t)(d.row0 + n) * d.ldd + tap * d.cinp + cip;
    dW[((int64_t)n * d.Cin + ci) * taps + tap] += *gp;
    *gp = 0.f;
  }
}

struct PredPrep {
  int64_t cls_w, cls_b, reg_w, reg_b, obj_w, obj_b;
  void *B, *BT;
  float *bias, *G, *gb;
};
struct PredPrep3 { PredPrep p[3]; };
// Block-diagonal prediction weight over the tower matrix [cls tower (hid) | reg tower (hid)]: output rows
// 0..3 = reg_preds (reg tower), 4 = obj_preds (reg tower), 5.. = cls_preds (cls tower)  (yolo_head.py:214-222, :236)
template <typename T>
__global__ void prep_pred_kernel(PredPrep3 pp, const float *__restrict__ params, int hid, int C) {
  pdl_prologue();
  const PredPrep p = pp.p[blockIdx.x];
  T *B = (T *)p.B, *BT = (T *)p.BT;
  for (int i = threadIdx.x; i < (5 + C) * hid; i += blockDim.x) {
    const int row = i / hid, k = i % hid;
    float v;
    int col;
    if (row < 4) { v = params[p.reg_w + row * hid + k]; col = hid + k; }
    else if (row == 4) { v = params[p.obj_w + k]; col = hid + k; }
    else { v = params[p.cls_w + (row - 5) * hid + k]; col = k; }
    B[row * 2 * hid + col] = from_f<T>(v);
    BT[col * 8 + row] = from_f<T>(v);
  }
  if (threadIdx.x < 5 + C) {
    const int r = threadIdx.x;
    p.bias[r] = r < 4 ? params[p.reg_b + r] : (r == 4 ? params[p.obj_b] : params[p.cls_b + r - 5]);
  }
}
__global__ void unprep_pred_kernel(PredPrep3 pp, float *__restrict__ grads, int hid, int C) {
  pdl_prologue();
  const PredPrep p = pp.p[blockIdx.x];
  for (int i = threadIdx.x; i < (5 + C) * hid; i += blockDim.x) {
    const int row = i / hid, k = i % hid;
    const int col = row < 5 ? hid + k : k;
    float *g = p.G + row * 2 * hid + col;
    if (row < 4) grads[p.reg_w + row * hid + k] += *g;
    else if (row == 4) grads[p.obj_w + k] += *g;
    else grads[p.cls_w + (row - 5) * hid + k] += *g;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 8 * 2 * hid; i += blockDim.x) p.G[i] = 0.f;
  if (threadIdx.x < 5 + C) {
    const int r = threadIdx.x;
    float *dst = r < 4 ? grads + p.reg_b + r : (r == 4 ? grads + p.obj_b : grads + p.cls_b + r - 5);
    *dst += p.gb[r];
  }
  __syncthreads();
  if (threadIdx.x < 8) p.gb[threadIdx.x] = 0.f;
}

PredPrep3 pred_prep_args(const leod_detect *h) {
  PredPrep3 a;
  for (int l = 0; l < 3; ++l) {
    const Pred &p = h->pred[l];
    a.p[l] = PredPrep{p.cls_w, p.cls_b, p.reg_w, p.reg_b, p.obj_w, p.obj_b, p.B, p.BT, p.bias, p.G, p.gb};
  }
  return a;
}

// prepared weights + descriptor table (called once from leod_detect_create)
int alloc_weights(leod_detect *h) {
  const size_t e = h->esz();
  std::vector<PrepDesc> descs;
  for (Conv &c : h->convs) {
    const int taps = c.k * c.k;
    LEOD_TRY(dev_alloc(h, &c.Wp, (size_t)c.cout * taps * c.cinp * e));
    const bool s2 = c.stride == 2;
    // input-gradient operand: tapped/plain [cin, taps*coutp]; stride-2 [taps*cinp, coutp]
    LEOD_TRY(dev_alloc(h, &c.WpT, (s2 ? (size_t)taps * c.cinp * c.coutp : (size_t)c.cin * taps * c.coutp) * e));
    if (c.use_G) LEOD_TRY(dev_alloc(h, (void **)&c.G, (size_t)c.cout * taps * c.cinp * sizeof(float)));
    for (int s = 0; s < c.nseg; ++s) {
      PrepDesc d;
      d.src_off = c.w[s];
      d.dst = c.Wp; d.dstT = c.WpT; d.G = c.G;
      d.ldd = taps * c.cinp;
      d.lddT = s2 ? c.coutp : taps * c.coutp;
      d.coutp = c.coutp; d.row0 = s * c.cseg; d.Cout = c.cseg; d.Cin = c.cin; d.k = c.k; d.cinp = c.cinp; d.rot = c.rot;
      d.tmode = s2 ? 1 : 0;
      descs.push_back(d);
      h->max_prep_items = std::max<int>(h->max_prep_items, c.cseg * taps * c.cin);
    }
  }
  h->n_prep = (int)descs.size();
  LEOD_TRY(dev_alloc(h, (void **)&h->prep_dev, descs.size() * sizeof(PrepDesc)));
  LEOD_CUDA(cudaMemcpy(h->prep_dev, descs.data(), descs.size() * sizeof(PrepDesc), cudaMemcpyHostToDevice));
  for (int l = 0; l < 3; ++l) {
    Pred &p = h->pred[l];
    LEOD_TRY(dev_alloc(h, &p.B, (size_t)16 * 2 * h->hid * e));
    LEOD_TRY(dev_alloc(h, &p.BT, (size_t)2 * h->hid * 8 * e));
    LEOD_TRY(dev_alloc(h, (void **)&p.bias, 8 * sizeof(float)));
    LEOD_TRY(dev_alloc(h, (void **)&p.G, (size_t)8 * 2 * h->hid * sizeof(float)));
    LEOD_TRY(dev_alloc(h, (void **)&p.gb, 8 * sizeof(float)));
  }
  return 0;
}

// ------------------------------------------------------------------ activation arena
int ensure_arena(leod_detect *h, int B) {
  if (B <= h->arena_B) return 0;
  if (h->arena) {
    LEOD_CUDA(cudaFree(h->arena));
    h->arena = nullptr;
    h->arena_B = 0;
  }
  const int64_t e = (int64_t)h->esz();
  int64_t cur = 0;
  auto take = [&](int64_t bytes) {
    const int64_t o = cur;
    cur += round_up(bytes, 256);
    return o;
  };
  PadGeom g[3];
  for (int l = 0; l < 3; ++l) g[l] = make_pad_geom(B, h->lh[l], h->lw[l]);
  std::vector<int64_t> az(h->acts.size()), adz(h->acts.size()), cy(h->convs.size()), cdy(h->convs.size());
  for (size_t i = 0; i < h->acts.size(); ++i) {
    az[i] = take(g[h->acts[i].level].R * h->acts[i].width * e);
    adz[i] = take(g[h->acts[i].level].R * h->acts[i].width * e);
  }
  int64_t col_bytes = 0;
  for (size_t i = 0; i < h->convs.size(); ++i) {
    const Conv &c = h->convs[i];
    cy[i] = take(g[c.lev_out].R * c.cout * e);
    cdy[i] = take(g[c.lev_out].R * c.cout * e);
    if (c.stride == 2) col_bytes = std::max(col_bytes, g[c.lev_out].R * 9 * c.cinp * e);
  }
  const int64_t o_col = take(col_bytes), o_dcol = take(col_bytes);
  // statistics: [2*cout + 4] doubles per convolution in op order (a phase is one contiguous range), forward and backward
  int64_t sd = 0, dd = 0;
  std::vector<int64_t> so(h->convs.size()), dof(h->convs.size());
  for (const Op &o : h->ops)
    if (o.kind == 0)
      for (int ci : o.convs) {
        so[ci] = sd; sd += 2 * h->convs[ci].cout + 4;    // multiples of 4 entries: the float copies are read as float4
        dof[ci] = dd; dd += 2 * h->convs[ci].cout + 4;
      }
  h->stats_doubles = sd;
  h->dstat_doubles = dd;
  const int64_t o_stats = take(sd * 8), o_dloc = take(dd * 8), o_dglob = take(dd * 8), o_fin = take(sd * 4), o_dfin = take(dd * 4);
  int64_t o_raw[3], o_draw[3];
  int A = 0;
  for (int l = 0; l < 3; ++l) {
    o_raw[l] = take(g[l].R * 8 * 4);
    o_draw[l] = take(g[l].R * 8 * e);
    A += h->lh[l] * h->lw[l];
  }
  const int64_t BA = (int64_t)B * A;
  const int64_t o_tout = take(BA * 8 * 4), o_miou = take(BA * 4), o_flags = take(BA), o_cnt = take(BA * 4), o_gt = take(BA * 4),
                o_assign = take(BA * 4), o_sums = take(8 * 8), o_losses = take(8 * 4);
  LEOD_CUDA(cudaMalloc(&h->arena, cur));
  LEOD_CUDA(cudaMemset(h->arena, 0, cur));   // zero borders of raw / draw, zero channel pads
  char *base = (char *)h->arena;
  for (size_t i = 0; i < h->acts.size(); ++i) {
    h->acts[i].z = base + az[i];
    h->acts[i].dz = base + adz[i];
  }
  for (size_t i = 0; i < h->convs.size(); ++i) {
    Conv &c = h->convs[i];
    c.Y = base + cy[i];
    c.dY = base + cdy[i];
    c.stats = (double *)(base + o_stats) + so[i];
    c.dloc = (double *)(base + o_dloc) + dof[i];
    c.dglob = (double *)(base + o_dglob) + dof[i];
    c.fin = (float *)(base + o_fin) + so[i];
    c.dfin = (float *)(base + o_dfin) + dof[i];
  }
  h->col = base + o_col; h->dcol = base + o_dcol;
  h->stats_all = (double *)(base + o_stats); h->dloc_all = (double *)(base + o_dloc); h->dglob_all = (double *)(base + o_dglob);
  for (int l = 0; l < 3; ++l) {
    h->pred[l].raw = (float *)(base + o_raw[l]);
    h->pred[l].draw = base + o_draw[l];
  }
  h->tout = (float *)(base + o_tout); h->miou = (float *)(base + o_miou); h->flags = (uint8_t *)(base + o_flags);
  h->match_cnt = (int *)(base + o_cnt); h->match_gt = (int *)(base + o_gt); h->assign = (int *)(base + o_assign);
  h->sums = (double *)(base + o_sums); h->losses = (float *)(base + o_losses);
  h->arena_B = B;
  return 0;
}

inline void *act_z(const leod_detect *h, int act, int off) { return (char *)h->acts[act].z + (int64_t)off * h->esz(); }
inline void *act_dz(const leod_detect *h, int act, int off) { return (char *)h->acts[act].dz + (int64_t)off * h->esz(); }

int gemm_nt(const leod_detect *h, const GemmNT &g, cudaStream_t st) {
  const double e = (double)h->esz();
  const double ka = g.taps.n > 1 ? (double)g.taps.cin : (double)g.K;   // the tapped A matrix is read once per tap from L2, once from HBM
  ProfScope ps(PK_CONV, 2.0 * g.M * g.N * g.K, e * ((double)g.M * ka + (double)g.N * g.K + (double)g.M * g.N), st, g.M, g.N, g.K);
  if (h->cfg.dtype == LEOD_BF16) return gemm_nt_tc(g, st);
  return gemm_nt_simt(h->cfg.dtype, g, st);
}
int gemm_tn(const leod_detect *h, const void *dY, int ldy, const void *X, int ldx, float *dW, int ldw, float *dbias, int M, int N, int K,
            cudaStream_t st, const ConvTaps *taps = nullptr) {
  const int nt = taps && taps->n > 1 ? taps->n : 1;
  ProfScope ps(PK_CONV, 2.0 * M * N * K * nt, (double)h->esz() * ((double)M * N + (double)M * K) + 8.0 * N * K * nt, st, M, N, K * nt);
  if (h->cfg.dtype == LEOD_BF16) return gemm_tn_tc(dY, ldy, X, ldx, dW, ldw, dbias, M, N, K, st, taps);
  return gemm_tn_simt(h->cfg.dtype, dY, ldy, X, ldx, dW, ldw, dbias, M, N, K, st, taps);
}

GemmNT plain(const void *A, int lda, const void *B, int ldb, void *C, int ldc, int M, int N, int K) {
  GemmNT g;
  g.A = A; g.lda = lda; g.A2 = nullptr; g.lda2 = 0; g.K1 = K;
  g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K;
  g.bias = nullptr; g.epi = EPI_NONE; g.R = nullptr; g.ldr = 0; g.aux = nullptr; g.ldaux = 0;
  return g;
}
ConvTaps taps3x3(int w2, int cin, int cinp, int sign) {
  ConvTaps t;
  t.n = 9; t.cin = cin; t.cinp = cinp;
  for (int k = 0; k < 9; ++k) t.off[k] = sign * ((k / 3 - 1) * w2 + (k % 3 - 1));
  return t;
}

BnLayer bn_layer(const leod_detect *h, const Conv &c, bool with_grads) {
  BnLayer l;
  l.C = c.cout; l.cseg = c.cseg;
  BnSeg *s[2] = {&l.s0, &l.s1};
  for (int i = 0; i < 2; ++i) {
    const int k = i < c.nseg ? i : 0;
    s[i]->gamma = h->params + c.bnw[k];
    s[i]->beta = h->params + c.bnb[k];
    s[i]->rmean = h->buffers ? h->buffers + c.rm[k] : nullptr;
    s[i]->rvar = h->buffers ? h->buffers + c.rv[k] : nullptr;
    s[i]->nbt = h->nbt ? h->nbt + c.nbt[k] : nullptr;
    s[i]->dgamma = with_grads ? h->grads + c.bnw[k] : nullptr;
    s[i]->dbeta = with_grads ? h->grads + c.bnb[k] : nullptr;
  }
  l.stats = c.stats; l.dloc = c.dloc; l.dglob = h->allreduce ? c.dglob : c.dloc;
  l.fin = c.fin; l.dfin = c.dfin;
  l.eps = h->cfg.bn_eps; l.momentum = h->cfg.bn_momentum;
  return l;
}

// convolution (+ batch statistics in training)
int conv_fwd_a(leod_detect *h, Conv &c, const PadGeom g[3], bool training, cudaStream_t st) {
  const int dt = h->cfg.dtype;
  const PadGeom &gi = g[c.lev_in], &go = g[c.lev_out];
  const void *A = act_z(h, c.in_act, c.in_off);
  const int lda = h->acts[c.in_act].width;
  const int taps = c.k * c.k;
  GemmNT gm;
  if (c.stride == 2) {
    LEOD_TRY(im2col_pad_s2(dt, A, lda, gi, h->col, 9 * c.cinp, go, c.cin, c.cinp, st));
    gm = plain(h->col, 9 * c.cinp, c.Wp, taps * c.cinp, c.Y, c.cout, (int)go.R, c.cout, 9 * c.cinp);
  } else if (c.k == 3) {
    gm = plain(A, lda, c.Wp, taps * c.cinp, c.Y, c.cout, (int)go.R, c.cout, 9 * c.cinp);
    gm.taps = taps3x3(go.w2, c.cin, c.cinp, +1);
  } else {
    gm = plain(A, lda, c.Wp, c.cinp, c.Y, c.cout, (int)go.R, c.cout, c.cin);
  }
  LEOD_TRY(gemm_nt(h, gm, st));
  if (training) {
    ProfScope ps(PK_OTHER, 0, (double)go.R * c.cout * h->esz(), st);
    LEOD_TRY(bn_stats(dt, c.Y, c.cout, go, bn_layer(h, c, false), h->allreduce ? 0 : 1, st));
  }
  return 0;
}
// BN-apply + SiLU into the consumer's matrix
int conv_fwd_b(leod_detect *h, Conv &c, const PadGeom g[3], bool training, cudaStream_t st) {
  const PadGeom &go = g[c.lev_out];
  ProfScope ps(PK_OTHER, 0, 2.0 * go.R * c.cout * h->esz(), st);
  return bn_apply_silu(h->cfg.dtype, c.Y, c.cout, go, bn_layer(h, c, false), act_z(h, c.out_act, c.out_off), h->acts[c.out_act].width,
                       training ? 1 : 0, st);
}
int conv_bwd_a(leod_detect *h, Conv &c, const PadGeom g[3], cudaStream_t st) {
  const PadGeom &go = g[c.lev_out];
  ProfScope ps(PK_OTHER, 0, 2.0 * go.R * c.cout * h->esz(), st);
  return bn_bwd_reduce(h->cfg.dtype, act_dz(h, c.out_act, c.out_off), h->acts[c.out_act].width, c.Y, c.cout, go, bn_layer(h, c, true),
                       h->allreduce ? 0 : 1, st);
}
int ensure_streams(leod_detect *h) {
  if (h->ev_fork) return 0;
  for (int i = 0; i < 2; ++i) {
    LEOD_CUDA(cudaStreamCreateWithFlags(&h->lvl_stream[i], cudaStreamNonBlocking));
    LEOD_CUDA(cudaStreamCreateWithFlags(&h->wg_stream[i], cudaStreamNonBlocking));
  }
  for (int i = 0; i < 4; ++i) LEOD_CUDA(cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming));
  for (Conv &c : h->convs) LEOD_CUDA(cudaEventCreateWithFlags(&c.ev_dy, cudaEventDisableTiming));
  LEOD_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  return 0;
}

// BN + SiLU backward (-> dY of this layer), the weight gradient on a side stream, the input gradient on `st`
int conv_bwd_b(leod_detect *h, Conv &c, const PadGeom g[3], cudaStream_t st) {
  const int dt = h->cfg.dtype;
  const PadGeom &gi = g[c.lev_in], &go = g[c.lev_out];
  const int taps = c.k * c.k;
  {
    ProfScope ps(PK_OTHER, 0, 3.0 * go.R * c.cout * h->esz(), st);
    LEOD_TRY(bn_bwd_apply(dt, act_dz(h, c.out_act, c.out_off), h->acts[c.out_act].width, c.Y, c.cout, go, bn_layer(h, c, true), c.dY, c.cout, st));
  }
  const void *X = act_z(h, c.in_act, c.in_off);
  const int ldx = h->acts[c.in_act].width;
  void *dX = act_dz(h, c.in_act, c.in_off);
  float *dW = c.use_G ? c.G : h->grads + c.w[0];
  const int ldw = c.use_G ? taps * c.cinp : c.cin;
  if (c.stride == 2) {   // the patch matrix is a shared scratch: everything stays on the main stream (two layers only)
    LEOD_TRY(im2col_pad_s2(dt, X, ldx, gi, h->col, 9 * c.cinp, go, c.cin, c.cinp, st));
    LEOD_TRY(gemm_tn(h, c.dY, c.cout, h->col, 9 * c.cinp, dW, ldw, nullptr, (int)go.R, c.cout, 9 * c.cinp, st));
    LEOD_TRY(gemm_nt(h, plain(c.dY, c.cout, c.WpT, c.coutp, h->dcol, 9 * c.cinp, (int)go.R, 9 * c.cinp, c.cout), st));
    LEOD_TRY(col2im_pad_s2(dt, h->dcol, 9 * c.cinp, go, dX, ldx, gi, c.cin, c.cinp, /*accumulate=*/1, st));
    return 0;
  }
  // weight gradient: off the critical path (nothing downstream in this backward pass reads it)
  cudaStream_t ws = leod_profiling_on() ? st : h->wg_stream[h->wg_next];
  if (ws != st) {
    h->wg_next ^= 1;
    LEOD_CUDA(cudaEventRecord(c.ev_dy, st));
    LEOD_CUDA(cudaStreamWaitEvent(ws, c.ev_dy, 0));
  }
  if (c.k == 3) {
    const ConvTaps tw = taps3x3(go.w2, c.cin, c.cinp, +1);
    LEOD_TRY(gemm_tn(h, c.dY, c.cout, X, ldx, dW, ldw, nullptr, (int)go.R, c.cout, c.cin, ws, &tw));
    GemmNT gm = plain(c.dY, c.cout, c.WpT, 9 * c.coutp, dX, ldx, (int)gi.R, c.cin, 9 * c.coutp);
    gm.taps = taps3x3(go.w2, c.cout, c.coutp, -1);
    LEOD_TRY(gemm_nt(h, gm, st));
  } else {
    LEOD_TRY(gemm_tn(h, c.dY, c.cout, X, ldx, dW, ldw, nullptr, (int)go.R, c.cout, c.cin, ws));
    LEOD_TRY(gemm_nt(h, plain(c.dY, c.cout, c.WpT, c.coutp, dX, ldx, (int)gi.R, c.cin, c.cout), st));
  }
  return 0;
}

// prediction GEMM of one level (yolo_head.py:214-222) and its backward
int pred_fwd(leod_detect *h, int l, const PadGeom g[3], cudaStream_t st) {
  Pred &p = h->pred[l];
  GemmNT gm = plain(h->acts[p.in_act].z, 2 * h->hid, p.B, 2 * h->hid, p.raw, 8, (int)g[l].R, 5 + h->cfg.num_classes, 2 * h->hid);
  gm.bias = p.bias;
  gm.out_f32 = 1;
  return gemm_nt(h, gm, st);
}
int pred_bwd(leod_detect *h, int l, const PadGeom g[3], cudaStream_t st) {
  Pred &p = h->pred[l];
  LEOD_TRY(gemm_tn(h, p.draw, 8, h->acts[p.in_act].z, 2 * h->hid, p.G, 2 * h->hid, p.gb, (int)g[l].R, 8, 2 * h->hid, st));
  return gemm_nt(h, plain(p.draw, 8, p.BT, 8, h->acts[p.in_act].dz, 2 * h->hid, (int)g[l].R, 2 * h->hid, 8), st);
}
// one phase of convolutions on one stream (+ the statistics exchange in data-parallel mode)
int phase_fwd(leod_detect *h, const Op &o, const PadGeom g[3], bool training, cudaStream_t st) {
  for (int ci : o.convs) LEOD_TRY(conv_fwd_a(h, h->convs[ci], g, training, st));
  if (training && h->allreduce) {
    const Conv &first = h->convs[o.convs.front()], &last = h->convs[o.convs.back()];
    const int64_t n = (last.stats + 2 * last.cout + 2) - first.stats;
    LEOD_REQUIRE(h->allreduce(h->allreduce_ctx, first.stats, n, (void *)st) == 0, "leod_fpn_head_fwd: statistics all-reduce callback failed");
    for (int ci : o.convs) LEOD_TRY(bn_finalize(bn_layer(h, h->convs[ci], false), st));
  }
  for (int ci : o.convs) LEOD_TRY(conv_fwd_b(h, h->convs[ci], g, training, st));
  return 0;
}
int phase_bwd(leod_detect *h, const Op &o, const PadGeom g[3], cudaStream_t st) {
  for (int ci : o.convs) LEOD_TRY(conv_bwd_a(h, h->convs[ci], g, st));
  if (h->allreduce) {
    const Conv &first = h->convs[o.convs.front()], &last = h->convs[o.convs.back()];
    const int64_t n = (last.dloc + 2 * last.cout + 2) - first.dloc;
    LEOD_TRY(device_copy(first.dglob, first.dloc, (size_t)n * 8, st));
    LEOD_REQUIRE(h->allreduce(h->allreduce_ctx, first.dglob, n, (void *)st) == 0, "leod_fpn_head_bwd: statistics all-reduce callback failed");
    for (int ci : o.convs) LEOD_TRY(bn_bwd_finalize(bn_layer(h, h->convs[ci], true), st));
  }
  // the members of a phase write disjoint gradient matrices (the twin towers' second convolutions: disjoint column halves)
  for (int ci : o.convs) LEOD_TRY(conv_bwd_b(h, h->convs[ci], g, st));
  return 0;
}
// The three pyramid levels of the head are independent chains of small kernels: in a single process they run side by side on
// three streams.  With a statistics exchange they run phase by phase on one stream so that every rank issues the same collectives.
bool head_multi_stream(const leod_detect *h) { return !h->allreduce && !leod_profiling_on(); }

void geoms(const leod_detect *h, int B, PadGeom g[3]) {
  for (int l = 0; l < 3; ++l) g[l] = make_pad_geom(B, h->lh[l], h->lw[l]);
}
HeadGeom head_geom(const leod_detect *h, int B) {
  HeadGeom g;
  g.B = B; g.C = h->cfg.num_classes;
  int a = 0;
  for (int l = 0; l < 3; ++l) {
    g.h[l] = h->lh[l]; g.w[l] = h->lw[l]; g.a0[l] = a; g.stride[l] = h->cfg.strides[l];
    g.P[l] = (h->lh[l] + 2) * (h->lw[l] + 2);
    a += h->lh[l] * h->lw[l];
  }
  g.A = a;
  return g;
}
SimotaCfg simota_cfg(const leod_detect *h) {
  SimotaCfg s;
  s.ignore_label = h->cfg.ignore_label;
  s.n_thresh = h->cfg.n_ignore_thresh;
  for (int i = 0; i < 8; ++i) s.thresh[i] = h->cfg.ignore_thresh[i];
  s.reg_w = h->cfg.reg_weight; s.obj_w = h->cfg.obj_weight; s.cls_w = h->cfg.cls_weight;
  return s;
}

int create_impl(const leod_detect_cfg *cfg, leod_detect **out, bool layout_only) {
  LEOD_REQUIRE(cfg && out, "leod_detect_create: null argument");
  LEOD_REQUIRE(cfg->dtype == LEOD_F32 || cfg->dtype == LEOD_BF16, "leod_detect_create: dtype %d", cfg->dtype);
  LEOD_REQUIRE(cfg->n_ignore_thresh >= 0 && cfg->n_ignore_thresh <= 8, "leod_detect_create: %d ignore thresholds", cfg->n_ignore_thresh);
  leod_detect *h = new leod_detect();
  h->cfg = *cfg;
  h->layout_only = layout_only;
  int rc = build_graph(h);
  if (rc == 0 && !layout_only) rc = alloc_weights(h);
  if (rc != 0) {
    leod_detect_destroy(h);
    return rc;
  }
  *out = h;
  return 0;
}

int info_impl(const std::vector<PEntry> &v, int i, char *name, size_t cap, int64_t *offset, int32_t *ndim, int64_t shape[4]) {
  if (i < 0) return (int)v.size();
  LEOD_REQUIRE(i < (int)v.size(), "entry index %d out of range", i);
  const PEntry &e = v[i];
  if (name && cap) snprintf(name, cap, "%s", e.name.c_str());
  if (offset) *offset = e.offset;
  if (ndim) *ndim = e.ndim;
  if (shape)
    for (int k = 0; k < 4; ++k) shape[k] = e.shape[k];
  return 0;
}

}  // namespace

extern "C" int leod_detect_create(const leod_detect_cfg *cfg, leod_detect_t **out) { return create_impl(cfg, out, false); }
extern "C" int leod_detect_layout_only(const leod_detect_cfg *cfg, leod_detect_t **out) { return create_impl(cfg, out, true); }
extern "C" void leod_detect_destroy(leod_detect_t *h) {
  if (!h) return;
  for (void *p : h->owned) cudaFree(p);
  if (h->arena) cudaFree(h->arena);
  for (int i = 0; i < 2; ++i) {
    if (h->lvl_stream[i]) cudaStreamDestroy(h->lvl_stream[i]);
    if (h->wg_stream[i]) cudaStreamDestroy(h->wg_stream[i]);
  }
  for (int i = 0; i < 4; ++i)
    if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  for (auto &c : h->convs)
    if (c.ev_dy) cudaEventDestroy(c.ev_dy);
  delete h;
}
extern "C" int leod_detect_param_info(const leod_detect_t *h, int i, char *name, size_t cap, int64_t *offset, int32_t *ndim, int64_t shape[4]) {
  return info_impl(h->pentries, i, name, cap, offset, ndim, shape);
}
extern "C" int leod_detect_buffer_info(const leod_detect_t *h, int i, char *name, size_t cap, int64_t *offset, int32_t *ndim, int64_t shape[4]) {
  return info_impl(h->bentries, i, name, cap, offset, ndim, shape);
}
extern "C" int leod_detect_counter_info(const leod_detect_t *h, int i, char *name, size_t cap, int64_t *offset) {
  return info_impl(h->nentries, i, name, cap, offset, nullptr, nullptr);
}
extern "C" int64_t leod_detect_param_count(const leod_detect_t *h) { return h->n_params; }
extern "C" int64_t leod_detect_buffer_count(const leod_detect_t *h) { return h->n_buffers; }
extern "C" int64_t leod_detect_counter_count(const leod_detect_t *h) { return h->n_nbt; }
extern "C" int leod_detect_num_anchors(const leod_detect_t *h) {
  int a = 0;
  for (int l = 0; l < 3; ++l) a += h->lh[l] * h->lw[l];
  return a;
}

extern "C" int leod_detect_bind(leod_detect_t *h, float *params, float *grads, float *buffers, int64_t *counters) {
  LEOD_REQUIRE(h && params && buffers, "leod_detect_bind: parameters and running-statistics buffers are required");
  h->params = params; h->grads = grads; h->buffers = buffers; h->nbt = (long long *)counters;
  return 0;
}
extern "C" int leod_detect_set_allreduce(leod_detect_t *h, leod_allreduce_fn fn, void *ctx) {
  h->allreduce = fn;
  h->allreduce_ctx = ctx;
  return 0;
}

extern "C" int leod_detect_prepare(leod_detect_t *h, void *stream) {
  LEOD_REQUIRE(h && !h->layout_only && h->params, "leod_detect_prepare: handle not bound");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(std::max(1, std::min(64, ceil_div(h->max_prep_items, 256))), h->n_prep);
  if (h->cfg.dtype == LEOD_F32) {
    LEOD_LAUNCH((prep_batched_kernel<float>), grid, 256, 0, st, h->prep_dev, h->params);
    LEOD_LAUNCH_CHECK();
    LEOD_LAUNCH((prep_pred_kernel<float>), 3, 256, 0, st, pred_prep_args(h), h->params, h->hid, h->cfg.num_classes);
  } else {
    LEOD_LAUNCH((prep_batched_kernel<bf16>), grid, 256, 0, st, h->prep_dev, h->params);
    LEOD_LAUNCH_CHECK();
    LEOD_LAUNCH((prep_pred_kernel<bf16>), 3, 256, 0, st, pred_prep_args(h), h->params, h->hid, h->cfg.num_classes);
  }
  LEOD_LAUNCH_CHECK();
  return 0;
}

extern "C" int leod_detect_reserve(leod_detect_t *h, int B) {
  LEOD_REQUIRE(h && !h->layout_only && B > 0, "leod_detect_reserve: bad argument");
  return ensure_arena(h, B);
}

// feats[l]: dense channels-last [B, h_l, w_l, c_l] in the handle's dtype (l = 0: stride 8 ... 2: stride 32).
extern "C" int leod_fpn_head_fwd(leod_detect_t *h, const void *const feats[3], int B, int training, float *preds, void *stream) {
  LEOD_REQUIRE(h && !h->layout_only && h->params, "leod_fpn_head_fwd: handle not bound");
  LEOD_REQUIRE(feats && feats[0] && feats[1] && feats[2] && preds && B > 0, "leod_fpn_head_fwd: null operand / empty batch");
  cudaStream_t st = (cudaStream_t)stream;
  LEOD_TRY(ensure_arena(h, B));
  h->cur_B = B;
  ++h->fwd_gen;
  h->loss_ready = h->dloss_ready = false;
  const int dt = h->cfg.dtype;
  PadGeom g[3];
  geoms(h, B, g);
  if (training) LEOD_TRY(device_zero_bytes(h->stats_all, (size_t)h->stats_doubles * 8, st));
  for (int l = 0; l < 3; ++l) {
    ProfScope ps(PK_OTHER, 0, 2.0 * g[l].R * h->cfg.in_channels[l] * h->esz(), st);
    LEOD_TRY(pad_gather(dt, feats[l], act_z(h, h->x_act[l], h->x_off[l]), h->acts[h->x_act[l]].width, g[l], h->cfg.in_channels[l], st));
  }
  LEOD_TRY(ensure_streams(h));
  const bool multi = head_multi_stream(h);
  const size_t n_ops = multi ? h->n_neck_ops : h->ops.size();
  for (size_t oi = 0; oi < n_ops; ++oi) {
    const Op &o = h->ops[oi];
    if (o.kind == 1) {
      const Act &s = h->acts[o.src_act], &d = h->acts[o.dst_act];
      ProfScope ps(PK_OTHER, 0, 1.25 * g[d.level].R * o.C * h->esz(), st);
      LEOD_TRY(upsample2x(dt, act_z(h, o.src_act, o.src_off), s.width, g[s.level], act_z(h, o.dst_act, o.dst_off), d.width, g[d.level], o.C, st));
      continue;
    }
    LEOD_TRY(phase_fwd(h, o, g, training != 0, st));
  }
  HeadPtrs rp;
  for (int l = 0; l < 3; ++l) rp.raw[l] = h->pred[l].raw;
  if (multi) {
    LEOD_CUDA(cudaEventRecord(h->ev_fork, st));
    for (int l = 2; l >= 0; --l) {   // the small levels first: their chains are pure latency
      cudaStream_t sl = l == 0 ? st : h->lvl_stream[l - 1];
      if (l > 0) LEOD_CUDA(cudaStreamWaitEvent(sl, h->ev_fork, 0));
      for (int k = 0; k < 4; ++k) {
        Op o;
        o.kind = 0;
        o.convs.assign(1, h->head_convs[l][k]);
        LEOD_TRY(phase_fwd(h, o, g, training != 0, sl));
      }
      LEOD_TRY(pred_fwd(h, l, g, sl));
      if (l > 0) LEOD_CUDA(cudaEventRecord(h->ev_join[l - 1], sl));
    }
    for (int l = 1; l < 3; ++l) LEOD_CUDA(cudaStreamWaitEvent(st, h->ev_join[l - 1], 0));
  } else {
    for (int l = 0; l < 3; ++l) LEOD_TRY(pred_fwd(h, l, g, st));
  }
  if (training) h->train_gen = h->fwd_gen;
  // eval: decode only; training: the decoded/logit copy and the geometry flags are produced by leod_simota_loss_fwd
  const HeadGeom hg = head_geom(h, B);
  ProfScope ps(PK_OTHER, 0, 0, st);
  return head_decode(rp, hg, simota_cfg(h), preds, nullptr, nullptr, 0, nullptr, nullptr, nullptr, st);
}

// labels: device fp32 [B, nmax, 7] rows (cls, cx, cy, w, h, obj_conf, cls_conf), zero rows = padding.
// losses_out: device fp32 [6] = loss, iou_loss, conf_loss, cls_loss, l1_loss (0), num_fg / num_gts.
extern "C" int leod_simota_loss_fwd(leod_detect_t *h, const float *labels, int nmax, float *losses_out, void *stream) {
  LEOD_REQUIRE(h && labels && losses_out && nmax > 0, "leod_simota_loss_fwd: null operand");
  LEOD_REQUIRE(h->cur_B > 0 && h->train_gen == h->fwd_gen, "leod_simota_loss_fwd: no training-mode leod_fpn_head_fwd precedes this call");
  cudaStream_t st = (cudaStream_t)stream;
  const HeadGeom hg = head_geom(h, h->cur_B);
  const SimotaCfg sc = simota_cfg(h);
  HeadPtrs rp;
  for (int l = 0; l < 3; ++l) rp.raw[l] = h->pred[l].raw;
  // second decode pass: same arithmetic as the eval output, plus logits and flags (preds itself was written by fpn_head_fwd)
  ProfScope ps(PK_OTHER, 0, 0, st);
  LEOD_TRY(head_decode(rp, hg, sc, nullptr, h->tout, labels, nmax, h->flags, h->match_cnt, h->match_gt, st));
  LEOD_TRY(simota_loss_fwd(hg, sc, h->tout, labels, nmax, h->flags, h->match_cnt, h->match_gt, h->assign, h->miou, h->sums, h->losses, st));
  LEOD_TRY(device_copy(losses_out, h->losses, 6 * sizeof(float), st));
  h->loss_ready = true;
  return 0;
}

// Test / diagnostics hook: the SimOTA result of the last leod_simota_loss_fwd.  assign_out: device int32 [B, A] label row per
// anchor (-1 = background); miou_out (may be NULL): device fp32 [B, A] IoU of the matched pair.
extern "C" int leod_simota_assignment(leod_detect_t *h, int32_t *assign_out, float *miou_out, void *stream) {
  LEOD_REQUIRE(h && assign_out && h->loss_ready, "leod_simota_assignment: no leod_simota_loss_fwd result");
  const int64_t n = (int64_t)h->cur_B * leod_detect_num_anchors(h);
  LEOD_TRY(device_copy(assign_out, h->assign, (size_t)n * 4, (cudaStream_t)stream));
  if (miou_out) LEOD_TRY(device_copy(miou_out, h->miou, (size_t)n * 4, (cudaStream_t)stream));
  return 0;
}

// Gradient of gscale[0] * loss (gscale: device fp32 scalar or NULL = 1) w.r.t. the raw head outputs, kept inside the handle
// for leod_fpn_head_bwd.  labels: the tensor given to leod_simota_loss_fwd.
extern "C" int leod_simota_loss_bwd(leod_detect_t *h, const float *labels, int nmax, const float *gscale, void *stream) {
  LEOD_REQUIRE(h && labels && h->loss_ready && h->train_gen == h->fwd_gen, "leod_simota_loss_bwd: no matching leod_simota_loss_fwd");
  cudaStream_t st = (cudaStream_t)stream;
  const HeadGeom hg = head_geom(h, h->cur_B);
  HeadGradPtrs dp;
  for (int l = 0; l < 3; ++l) dp.draw[l] = h->pred[l].draw;
  ProfScope ps(PK_OTHER, 0, 0, st);
  LEOD_TRY(simota_loss_bwd(h->cfg.dtype, hg, simota_cfg(h), h->tout, labels, nmax, h->flags, h->assign, h->miou, h->sums, gscale, dp, st));
  h->dloss_ready = true;
  return 0;
}

// Raw-output gradients as one dense tensor: device fp32 [B, A, 8] (columns: reg x4, obj logit, cls logits, zero padding).
// get: what leod_simota_loss_bwd computed.  set: replaces it, so leod_fpn_head_bwd back-propagates a caller-defined loss on the
// raw head outputs (used by the parity tests to check the backward through a smooth loss).
extern "C" int leod_detect_get_raw_grad(leod_detect_t *h, float *out, void *stream) {
  LEOD_REQUIRE(h && out && h->dloss_ready, "leod_detect_get_raw_grad: no raw-output gradient available");
  HeadGradPtrs dp;
  for (int l = 0; l < 3; ++l) dp.draw[l] = h->pred[l].draw;
  return raw_grad_copy(h->cfg.dtype, head_geom(h, h->cur_B), out, dp, 0, (cudaStream_t)stream);
}
extern "C" int leod_detect_set_raw_grad(leod_detect_t *h, const float *in, void *stream) {
  LEOD_REQUIRE(h && in && h->cur_B > 0 && h->train_gen == h->fwd_gen, "leod_detect_set_raw_grad: no training-mode forward to attach the gradient to");
  HeadGradPtrs dp;
  for (int l = 0; l < 3; ++l) dp.draw[l] = h->pred[l].draw;
  LEOD_TRY(raw_grad_copy(h->cfg.dtype, head_geom(h, h->cur_B), const_cast<float *>(in), dp, 1, (cudaStream_t)stream));
  h->dloss_ready = true;
  return 0;
}
// Raw head outputs of the last forward: device fp32 [B, A, 8] gathered from the per-level matrices (diagnostics).
extern "C" int leod_detect_get_raw(leod_detect_t *h, float *out, void *stream) {
  LEOD_REQUIRE(h && out && h->cur_B > 0, "leod_detect_get_raw: no forward has run");
  HeadGradPtrs dp;
  for (int l = 0; l < 3; ++l) dp.draw[l] = h->pred[l].raw;
  return raw_grad_copy(LEOD_F32, head_geom(h, h->cur_B), out, dp, 0, (cudaStream_t)stream);
}

// Backward of leod_fpn_head_fwd(training) from the raw-output gradients left by leod_simota_loss_bwd.  Parameter gradients
// are ACCUMULATED into the bound gradient buffer; dfeats[l] ([B, h_l, w_l, c_l], handle dtype) are written.
extern "C" int leod_fpn_head_bwd(leod_detect_t *h, void *const dfeats[3], void *stream) {
  LEOD_REQUIRE(h && h->grads, "leod_fpn_head_bwd: no gradient buffer bound");
  LEOD_REQUIRE(h->dloss_ready && h->train_gen == h->fwd_gen,
               "leod_fpn_head_bwd: the activations of the training forward were overwritten by a later forward (or no loss backward ran)");
  cudaStream_t st = (cudaStream_t)stream;
  const int dt = h->cfg.dtype, B = h->cur_B;
  PadGeom g[3];
  geoms(h, B, g);
  LEOD_TRY(device_zero_bytes(h->dloc_all, (size_t)h->dstat_doubles * 8, st));
  LEOD_TRY(ensure_streams(h));
  const bool multi = head_multi_stream(h);
  size_t n_ops = h->ops.size();
  if (multi) {
    n_ops = h->n_neck_ops;
    LEOD_CUDA(cudaEventRecord(h->ev_fork, st));
    for (int l = 2; l >= 0; --l) {
      cudaStream_t sl = l == 0 ? st : h->lvl_stream[l - 1];
      if (l > 0) LEOD_CUDA(cudaStreamWaitEvent(sl, h->ev_fork, 0));
      LEOD_TRY(pred_bwd(h, l, g, sl));
      for (int k = 3; k >= 0; --k) {
        Op o;
        o.kind = 0;
        o.convs.assign(1, h->head_convs[l][k]);
        LEOD_TRY(phase_bwd(h, o, g, sl));
      }
      if (l > 0) LEOD_CUDA(cudaEventRecord(h->ev_join[l - 1], sl));
    }
    for (int l = 1; l < 3; ++l) LEOD_CUDA(cudaStreamWaitEvent(st, h->ev_join[l - 1], 0));
  } else {
    for (int l = 0; l < 3; ++l) LEOD_TRY(pred_bwd(h, l, g, st));
  }
  for (int oi = (int)n_ops - 1; oi >= 0; --oi) {
    const Op &o = h->ops[oi];
    if (o.kind == 1) {
      const Act &s = h->acts[o.src_act], &d = h->acts[o.dst_act];
      ProfScope ps(PK_OTHER, 0, 1.5 * g[d.level].R * o.C * h->esz(), st);
      LEOD_TRY(upsample2x_bwd(dt, act_dz(h, o.dst_act, o.dst_off), d.width, g[d.level], act_dz(h, o.src_act, o.src_off), s.width, g[s.level], o.C,
                              /*accumulate=*/1, st));
      continue;
    }
    LEOD_TRY(phase_bwd(h, o, g, st));
  }
  for (int i = 0; i < 2; ++i) {   // the weight-gradient streams join before the gradients are un-permuted / handed over
    LEOD_CUDA(cudaEventRecord(h->ev_join[2 + i], h->wg_stream[i]));
    LEOD_CUDA(cudaStreamWaitEvent(st, h->ev_join[2 + i], 0));
  }
  {
    dim3 grid(std::max(1, std::min(64, ceil_div(h->max_prep_items, 256))), h->n_prep);
    LEOD_LAUNCH((unprep_batched_kernel), grid, 256, 0, st, h->prep_dev, h->grads);
    LEOD_LAUNCH_CHECK();
    LEOD_LAUNCH((unprep_pred_kernel), 3, 256, 0, st, pred_prep_args(h), h->grads, h->hid, h->cfg.num_classes);
    LEOD_LAUNCH_CHECK();
  }
  if (dfeats)
    for (int l = 0; l < 3; ++l)
      if (dfeats[l])
        LEOD_TRY(pad_scatter(dt, act_dz(h, h->x_act[l], h->x_off[l]), h->acts[h->x_act[l]].width, dfeats[l], g[l], h->cfg.in_channels[l], st));
  return 0;
}
