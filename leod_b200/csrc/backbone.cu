// Host-side executor of the recurrent MaxViT backbone: owns the prepared weights, the workspace and
// the per-step kernel schedule (forward and backward of one timestep over the four stages).
// Reference semantics: models/detection/recurrent_backbone/maxvit_rnn.py:97-115, 182-201.
#include <algorithm>
#include <string>
#include <vector>

#include "common.cuh"

namespace {

struct ParamEntry {
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

struct BlockP {  // offsets into the flat fp32 parameter buffer (-1 = absent)
  int64_t n1w = -1, n1b = -1, qkvw, qkvb, projw, projb, ls1, n2w, n2b, fc1w, fc1b, fc2w, fc2b, ls2;
};
struct StageP {
  int64_t convw, lnw, lnb;
  BlockP blk[2];
  int64_t lstmw, lstmb;
};
struct BlockW {  // prepared copies (operand dtype) + LayerScale scratch (fp32)
  void *Wqkv, *WqkvT, *Wproj, *WprojT, *W1, *W1T, *W2, *W2T;
  float *bproj, *b2, *Gproj, *sproj, *G2, *s2;
};
struct StageW {
  void *Wconv, *WconvT;
  float *Gconv;  // permuted-layout conv weight gradient scratch (stages 1..3)
  BlockW blk[2];
  void *Wl, *WlT;
};
struct StageD {
  int Cin, C, ksz, stride, pad, Hi, Wi, Ho, Wo, K, Kp;
};
struct SaveOff {  // element offsets (operand dtype) inside one step's save buffer, per stage
  int64_t y0, x0;
  struct {
    int64_t xn1, qkv, att, x1, xn2, u, a, x2;
  } blk[2];
  int64_t gates;
};

}  // namespace

struct leod_backbone {
  leod_backbone_cfg cfg;
  StageD d[4];
  StageP p[4];
  StageW w[4];
  std::vector<ParamEntry> entries;
  int64_t n_params = 0;
  float *params = nullptr, *grads = nullptr;
  std::vector<void *> owned;  // cudaMalloc'ed
  // workspace
  int ws_B = 0;
  void *ws_col = nullptr, *ws_big4 = nullptr, *ws_qkv = nullptr, *ws_a = nullptr, *ws_b = nullptr, *ws_c = nullptr,
       *ws_hint = nullptr, *ws_save = nullptr;
  size_t esz() const { return cfg.dtype == LEOD_BF16 ? 2 : 4; }
  int gemm_impl = 0;  // 0 SIMT, 1 tensor core (bf16 only)
};

namespace {

int64_t add_param(leod_backbone *h, const std::string &name, std::initializer_list<int64_t> shape) {
  ParamEntry e;
  e.name = name;
  e.ndim = (int)shape.size();
  int i = 0;
  for (auto s : shape) e.shape[i++] = s;
  for (; i < 4; ++i) e.shape[i] = 1;
  e.offset = h->n_params;
  h->n_params += round_up(e.numel(), 4);  // keep every tensor 16-byte aligned
  h->entries.push_back(e);
  return e.offset;
}

int dev_alloc(leod_backbone *h, void **p, size_t bytes, bool zero = true) {
  LEOD_CUDA(cudaMalloc(p, bytes ? bytes : 16));
  if (zero) LEOD_CUDA(cudaMemset(*p, 0, bytes ? bytes : 16));
  h->owned.push_back(*p);
  return 0;
}

void compute_save_layout(const leod_backbone *h, int B, SaveOff off[4], int64_t *total_elems) {
  int64_t cur = 0;
  auto take = [&](int64_t n) {
    int64_t o = cur;
    cur += round_up(n, 128);
    return o;
  };
  for (int s = 0; s < 4; ++s) {
    const StageD &d = h->d[s];
    const int64_t M = (int64_t)B * d.Ho * d.Wo, C = d.C;
    off[s].y0 = take(M * C);
    off[s].x0 = take(M * C);
    for (int b = 0; b < 2; ++b) {
      off[s].blk[b].xn1 = (b == 1) ? take(M * C) : -1;
      off[s].blk[b].qkv = take(M * 3 * C);
      off[s].blk[b].att = take(M * C);
      off[s].blk[b].x1 = take(M * C);
      off[s].blk[b].xn2 = take(M * C);
      off[s].blk[b].u = take(M * h->cfg.mlp_ratio * C);
      off[s].blk[b].a = take(M * h->cfg.mlp_ratio * C);
      off[s].blk[b].x2 = take(M * C);
    }
    off[s].gates = take(M * 4 * C);
  }
  *total_elems = cur;
}

int ensure_workspace(leod_backbone *h, int B) {
  if (B <= h->ws_B) return 0;
  for (void **p : {&h->ws_col, &h->ws_big4, &h->ws_qkv, &h->ws_a, &h->ws_b, &h->ws_c, &h->ws_hint, &h->ws_save}) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  const size_t e = h->esz();
  int64_t col = 0, mc = 0;
  for (int s = 0; s < 4; ++s) {
    const StageD &d = h->d[s];
    const int64_t M = (int64_t)B * d.Ho * d.Wo;
    col = std::max<int64_t>(col, M * d.Kp);
    mc = std::max<int64_t>(mc, M * d.C);
  }
  SaveOff off[4];
  int64_t save_elems;
  compute_save_layout(h, B, off, &save_elems);
  LEOD_CUDA(cudaMalloc(&h->ws_col, col * e));
  LEOD_CUDA(cudaMalloc(&h->ws_big4, mc * std::max(4, h->cfg.mlp_ratio) * e));
  LEOD_CUDA(cudaMalloc(&h->ws_qkv, mc * 3 * e));
  LEOD_CUDA(cudaMalloc(&h->ws_a, mc * e));
  LEOD_CUDA(cudaMalloc(&h->ws_b, mc * e));
  LEOD_CUDA(cudaMalloc(&h->ws_c, mc * e));
  LEOD_CUDA(cudaMalloc(&h->ws_hint, mc * e));
  LEOD_CUDA(cudaMalloc(&h->ws_save, save_elems * e));
  h->ws_B = B;
  return 0;
}

inline void *eoff(const leod_backbone *h, const void *base, int64_t elems) { return (char *)base + elems * h->esz(); }

int gemm_nt(leod_backbone *h, const GemmNT &g, cudaStream_t st) {
  const double e = (double)h->esz();
  double bytes = e * ((double)g.M * g.K + (double)g.N * g.K + (double)g.M * g.N);
  if (g.epi == EPI_GELU || g.epi == EPI_RESID || g.epi == EPI_GELU_BWD) bytes += e * (double)g.M * g.N;
  ProfScope ps(PK_GEMM_NT, 2.0 * g.M * g.N * g.K, bytes, st, g.M, g.N, g.K);
  if (h->gemm_impl == 1 && h->cfg.dtype == LEOD_BF16) return gemm_nt_tc(g, st);
  return gemm_nt_simt(h->cfg.dtype, g, st);
}
int gemm_tn(leod_backbone *h, const void *dY, int ldy, const void *X, int ldx, float *dW, int ldw, float *dbias, int M, int N,
            int K, cudaStream_t st) {
  ProfScope ps(PK_GEMM_TN, 2.0 * M * N * K, (double)h->esz() * ((double)M * N + (double)M * K) + 8.0 * N * K, st, M, N, K);
  if (h->gemm_impl == 1 && h->cfg.dtype == LEOD_BF16) return gemm_tn_tc(dY, ldy, X, ldx, dW, ldw, dbias, M, N, K, st);
  return gemm_tn_simt(h->cfg.dtype, dY, ldy, X, ldx, dW, ldw, dbias, M, N, K, st);
}

GemmNT mk(const void *A, int lda, const void *B, int ldb, void *C, int ldc, int M, int N, int K, const float *bias = nullptr,
          int epi = EPI_NONE, const void *R = nullptr, int ldr = 0, void *aux = nullptr, int ldaux = 0) {
  GemmNT g;
  g.A = A; g.lda = lda; g.A2 = nullptr; g.lda2 = 0; g.K1 = K;
  g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K;
  g.bias = bias; g.epi = epi; g.R = R; g.ldr = ldr; g.aux = aux; g.ldaux = ldaux;
  return g;
}

__global__ void conv_grad_unpermute_kernel(float *__restrict__ G, float *__restrict__ dW, int N, int Cin, int ksz) {
  const int K = Cin * ksz * ksz;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * K) return;
  const int n = (int)(idx / K), k = (int)(idx % K);
  const int cin = k % Cin, kx = (k / Cin) % ksz, ky = k / (Cin * ksz);
  dW[(((size_t)n * Cin + cin) * ksz + ky) * ksz + kx] += G[idx];
  G[idx] = 0.f;
}

}  // namespace

// ============================================================================================ API
static int backbone_create_impl(const leod_backbone_cfg *cfg, leod_backbone_t **out, bool alloc_device) {
  LEOD_REQUIRE(cfg && out, "leod_backbone_create: null argument");
  LEOD_REQUIRE(cfg->dtype == LEOD_F32 || cfg->dtype == LEOD_BF16, "unsupported dtype %d", cfg->dtype);
  LEOD_REQUIRE(cfg->embed_dim % 8 == 0, "embed_dim must be a multiple of 8 (got %d)", cfg->embed_dim);
  LEOD_REQUIRE(cfg->embed_dim % cfg->dim_head == 0, "embed_dim %d not divisible by dim_head %d", cfg->embed_dim, cfg->dim_head);
  LEOD_REQUIRE(cfg->in_h % 32 == 0 && cfg->in_w % 32 == 0,
               "input resolution %dx%d must be a multiple of 32", cfg->in_h, cfg->in_w);
  LEOD_REQUIRE((cfg->in_h / 32) % cfg->part_h == 0 && (cfg->in_w / 32) % cfg->part_w == 0,
               "stage-4 map %dx%d not divisible by partition %dx%d", cfg->in_h / 32, cfg->in_w / 32, cfg->part_h, cfg->part_w);
  leod_backbone *h = new leod_backbone();
  h->cfg = *cfg;
  h->gemm_impl = (cfg->dtype == LEOD_BF16) ? 1 : 0;
  int Hi = cfg->in_h, Wi = cfg->in_w, Cin = cfg->in_channels;
  for (int s = 0; s < 4; ++s) {
    StageD &d = h->d[s];
    d.Cin = Cin;
    d.C = cfg->embed_dim << s;
    d.stride = s == 0 ? 4 : 2;
    d.ksz = (d.stride - 1) * 2 + 1;
    d.pad = d.ksz / 2;
    d.Hi = Hi; d.Wi = Wi;
    d.Ho = Hi / d.stride; d.Wo = Wi / d.stride;
    d.K = Cin * d.ksz * d.ksz;
    d.Kp = (int)round_up(d.K, 8);
    Hi = d.Ho; Wi = d.Wo; Cin = d.C;
    const int64_t C = d.C, R = cfg->mlp_ratio;
    const std::string sp = "stages." + std::to_string(s);
    StageP &p = h->p[s];
    p.convw = add_param(h, sp + ".downsample_cf2cl.conv.weight", {C, d.Cin, d.ksz, d.ksz});
    p.lnw = add_param(h, sp + ".downsample_cf2cl.norm.weight", {C});
    p.lnb = add_param(h, sp + ".downsample_cf2cl.norm.bias", {C});
    for (int b = 0; b < 2; ++b) {
      const std::string bp = sp + ".att_blocks.0." + (b == 0 ? "att_window" : "att_grid");
      BlockP &q = p.blk[b];
      if (b == 1) {
        q.n1w = add_param(h, bp + ".norm1.weight", {C});
        q.n1b = add_param(h, bp + ".norm1.bias", {C});
      }
      q.qkvw = add_param(h, bp + ".self_attn.qkv.weight", {3 * C, C});
      q.qkvb = add_param(h, bp + ".self_attn.qkv.bias", {3 * C});
      q.projw = add_param(h, bp + ".self_attn.proj.weight", {C, C});
      q.projb = add_param(h, bp + ".self_attn.proj.bias", {C});
      q.ls1 = add_param(h, bp + ".ls1.gamma", {C});
      q.n2w = add_param(h, bp + ".norm2.weight", {C});
      q.n2b = add_param(h, bp + ".norm2.bias", {C});
      q.fc1w = add_param(h, bp + ".mlp.net.0.0.weight", {R * C, C});
      q.fc1b = add_param(h, bp + ".mlp.net.0.0.bias", {R * C});
      q.fc2w = add_param(h, bp + ".mlp.net.2.weight", {C, R * C});
      q.fc2b = add_param(h, bp + ".mlp.net.2.bias", {C});
      q.ls2 = add_param(h, bp + ".ls2.gamma", {C});
    }
    p.lstmw = add_param(h, sp + ".lstm.conv1x1.weight", {4 * C, 2 * C, 1, 1});
    p.lstmb = add_param(h, sp + ".lstm.conv1x1.bias", {4 * C});
  }
  if (!alloc_device) {
    *out = h;
    return 0;
  }
  // prepared weights + scratch
  const size_t e = h->esz();
  int rc = 0;
  for (int s = 0; s < 4 && rc == 0; ++s) {
    const StageD &d = h->d[s];
    const int64_t C = d.C, R = cfg->mlp_ratio;
    StageW &w = h->w[s];
    rc |= dev_alloc(h, &w.Wconv, C * d.Kp * e);
    rc |= dev_alloc(h, &w.WconvT, (int64_t)d.Kp * C * e);
    rc |= dev_alloc(h, (void **)&w.Gconv, C * d.K * sizeof(float));
    for (int b = 0; b < 2; ++b) {
      BlockW &q = w.blk[b];
      rc |= dev_alloc(h, &q.Wqkv, 3 * C * C * e);
      rc |= dev_alloc(h, &q.WqkvT, 3 * C * C * e);
      rc |= dev_alloc(h, &q.Wproj, C * C * e);
      rc |= dev_alloc(h, &q.WprojT, C * C * e);
      rc |= dev_alloc(h, &q.W1, R * C * C * e);
      rc |= dev_alloc(h, &q.W1T, R * C * C * e);
      rc |= dev_alloc(h, &q.W2, R * C * C * e);
      rc |= dev_alloc(h, &q.W2T, R * C * C * e);
      rc |= dev_alloc(h, (void **)&q.bproj, C * sizeof(float));
      rc |= dev_alloc(h, (void **)&q.b2, C * sizeof(float));
      rc |= dev_alloc(h, (void **)&q.Gproj, C * C * sizeof(float));
      rc |= dev_alloc(h, (void **)&q.sproj, C * sizeof(float));
      rc |= dev_alloc(h, (void **)&q.G2, R * C * C * sizeof(float));
      rc |= dev_alloc(h, (void **)&q.s2, C * sizeof(float));
    }
    rc |= dev_alloc(h, &w.Wl, 8 * C * C * e);
    rc |= dev_alloc(h, &w.WlT, 8 * C * C * e);
  }
  if (rc != 0) {
    leod_backbone_destroy(h);
    return -1;
  }
  *out = h;
  return 0;
}

extern "C" int leod_backbone_create(const leod_backbone_cfg *cfg, leod_backbone_t **out) {
  return backbone_create_impl(cfg, out, true);
}
extern "C" int leod_backbone_layout_only(const leod_backbone_cfg *cfg, leod_backbone_t **out) {
  return backbone_create_impl(cfg, out, false);
}

extern "C" void leod_backbone_destroy(leod_backbone_t *h) {
  if (!h) return;
  for (void *p : h->owned) cudaFree(p);
  for (void *p : {h->ws_col, h->ws_big4, h->ws_qkv, h->ws_a, h->ws_b, h->ws_c, h->ws_hint, h->ws_save})
    if (p) cudaFree(p);
  delete h;
}

extern "C" int leod_backbone_param_info(const leod_backbone_t *h, int i, char *name, size_t name_cap, int64_t *offset,
                                        int32_t *ndim, int64_t shape[4]) {
  if (!h) return -1;
  if (i < 0) return (int)h->entries.size();
  if (i >= (int)h->entries.size()) return -1;
  const ParamEntry &e = h->entries[i];
  if (name && name_cap) snprintf(name, name_cap, "%s", e.name.c_str());
  if (offset) *offset = e.offset;
  if (ndim) *ndim = e.ndim;
  if (shape)
    for (int k = 0; k < 4; ++k) shape[k] = e.shape[k];
  return 0;
}

extern "C" int64_t leod_backbone_param_count(const leod_backbone_t *h) { return h ? h->n_params : -1; }

extern "C" int leod_backbone_bind(leod_backbone_t *h, float *params_dev, float *grads_dev) {
  LEOD_REQUIRE(h && params_dev, "leod_backbone_bind: null argument");
  h->params = params_dev;
  h->grads = grads_dev;
  return 0;
}

extern "C" int leod_backbone_set_gemm_impl(leod_backbone_t *h, int impl) {
  LEOD_REQUIRE(h, "null handle");
  LEOD_REQUIRE(impl == 0 || (impl == 1 && h->cfg.dtype == LEOD_BF16), "gemm impl %d unsupported for dtype %d", impl, h->cfg.dtype);
  h->gemm_impl = impl;
  return 0;
}

extern "C" int leod_backbone_prepare(leod_backbone_t *h, void *stream) {
  LEOD_REQUIRE(h && h->params, "leod_backbone_prepare: parameters not bound");
  cudaStream_t st = (cudaStream_t)stream;
  const int dt = h->cfg.dtype;
  const float *P = h->params;
  for (int s = 0; s < 4; ++s) {
    const StageD &d = h->d[s];
    const int C = d.C, R = h->cfg.mlp_ratio;
    const StageP &p = h->p[s];
    StageW &w = h->w[s];
    LEOD_TRY(prep_weight(dt, P + p.convw, nullptr, w.Wconv, d.Kp, w.WconvT, C, C, d.K, s == 0 ? 0 : 1, d.Cin, d.ksz, st));
    for (int b = 0; b < 2; ++b) {
      const BlockP &q = p.blk[b];
      BlockW &bw = w.blk[b];
      LEOD_TRY(prep_weight(dt, P + q.qkvw, nullptr, bw.Wqkv, C, bw.WqkvT, 3 * C, 3 * C, C, 0, 0, 0, st));
      LEOD_TRY(prep_weight(dt, P + q.projw, P + q.ls1, bw.Wproj, C, bw.WprojT, C, C, C, 0, 0, 0, st));
      LEOD_TRY(prep_scaled_bias(P + q.projb, P + q.ls1, bw.bproj, C, st));
      LEOD_TRY(prep_weight(dt, P + q.fc1w, nullptr, bw.W1, C, bw.W1T, R * C, R * C, C, 0, 0, 0, st));
      LEOD_TRY(prep_weight(dt, P + q.fc2w, P + q.ls2, bw.W2, R * C, bw.W2T, C, C, R * C, 0, 0, 0, st));
      LEOD_TRY(prep_scaled_bias(P + q.fc2b, P + q.ls2, bw.b2, C, st));
    }
    LEOD_TRY(prep_weight(dt, P + p.lstmw, nullptr, w.Wl, 2 * C, w.WlT, 4 * C, 4 * C, 2 * C, 0, 0, 0, st));
  }
  return 0;
}

extern "C" int64_t leod_backbone_save_bytes(const leod_backbone_t *h, int B) {
  if (!h || B <= 0) return -1;
  SaveOff off[4];
  int64_t elems;
  compute_save_layout(h, B, off, &elems);
  return elems * (int64_t)h->esz();
}

extern "C" int leod_backbone_reserve(leod_backbone_t *h, int B) {
  LEOD_REQUIRE(h && B > 0, "leod_backbone_reserve: bad argument");
  return ensure_workspace(h, B);
}

extern "C" int leod_backbone_step_fwd(leod_backbone_t *h, const void *x, int x_dtype, int x_h, int x_w, int B,
                                      const void *const h_prev[4], const void *const c_prev[4], void *const h_out[4],
                                      void *const c_out[4], void *save, void *stream) {
  LEOD_REQUIRE(h && x && h_out && c_out && B > 0, "leod_backbone_step_fwd: bad argument");
  LEOD_REQUIRE(h->params, "leod_backbone_step_fwd: parameters not bound");
  LEOD_REQUIRE(x_h <= h->cfg.in_h && x_w <= h->cfg.in_w, "input %dx%d larger than padded resolution %dx%d", x_h, x_w,
               h->cfg.in_h, h->cfg.in_w);
  cudaStream_t st = (cudaStream_t)stream;
  LEOD_TRY(ensure_workspace(h, B));
  const int dt = h->cfg.dtype;
  const float eps = h->cfg.ln_eps;
  const float *P = h->params;
  SaveOff off[4];
  int64_t save_elems;
  compute_save_layout(h, B, off, &save_elems);
  void *sv = save ? save : h->ws_save;
  for (int s = 0; s < 4; ++s) {
    const StageD &d = h->d[s];
    const StageP &p = h->p[s];
    const StageW &w = h->w[s];
    const int M = B * d.Ho * d.Wo, C = d.C, R = h->cfg.mlp_ratio;
    const SaveOff &o = off[s];
    void *y0 = eoff(h, sv, o.y0), *x0 = eoff(h, sv, o.x0);
    if (s == 0)
      LEOD_TRY(im2col_nchw(x_dtype, dt, x, h->ws_col, B, d.Cin, x_h, x_w, d.Hi, d.Wi, d.ksz, d.stride, d.pad, d.Kp, st));
    else
      LEOD_TRY(im2col_nhwc(dt, h_out[s - 1], h->ws_col, B, d.Hi, d.Wi, d.Cin, d.ksz, d.stride, d.pad, d.Kp, st));
    LEOD_TRY(gemm_nt(h, mk(h->ws_col, d.Kp, w.Wconv, d.Kp, y0, C, M, C, d.Kp), st));
    LEOD_TRY(layernorm_fwd(dt, y0, P + p.lnw, P + p.lnb, x0, M, C, 1e-5f, st));
    const void *xin = x0;
    for (int b = 0; b < 2; ++b) {
      const BlockP &q = p.blk[b];
      const BlockW &bw = w.blk[b];
      void *qkv = eoff(h, sv, o.blk[b].qkv), *att = eoff(h, sv, o.blk[b].att), *x1 = eoff(h, sv, o.blk[b].x1);
      void *xn2 = eoff(h, sv, o.blk[b].xn2), *u = eoff(h, sv, o.blk[b].u), *a = eoff(h, sv, o.blk[b].a);
      void *x2 = eoff(h, sv, o.blk[b].x2);
      const void *ain = xin;
      if (b == 1) {
        void *xn1 = eoff(h, sv, o.blk[b].xn1);
        LEOD_TRY(layernorm_fwd(dt, xin, P + q.n1w, P + q.n1b, xn1, M, C, eps, st));
        ain = xn1;
      }
      LEOD_TRY(gemm_nt(h, mk(ain, C, bw.Wqkv, C, qkv, 3 * C, M, 3 * C, C, P + q.qkvb), st));
      LEOD_TRY(attention_fwd(dt, qkv, att, B, d.Ho, d.Wo, C, h->cfg.dim_head, h->cfg.part_h, h->cfg.part_w, b == 0, st));
      LEOD_TRY(gemm_nt(h, mk(att, C, bw.Wproj, C, x1, C, M, C, C, bw.bproj, EPI_RESID, xin, C), st));
      LEOD_TRY(layernorm_fwd(dt, x1, P + q.n2w, P + q.n2b, xn2, M, C, eps, st));
      LEOD_TRY(gemm_nt(h, mk(xn2, C, bw.W1, C, a, R * C, M, R * C, C, P + q.fc1b, EPI_GELU, nullptr, 0, u, R * C), st));
      LEOD_TRY(gemm_nt(h, mk(a, R * C, bw.W2, R * C, x2, C, M, C, R * C, bw.b2, EPI_RESID, x1, C), st));
      xin = x2;
    }
    void *gates = eoff(h, sv, o.gates);
    GemmNT g = mk(xin, C, w.Wl, 2 * C, gates, 4 * C, M, 4 * C, C, P + p.lstmb);
    const void *hp = h_prev ? h_prev[s] : nullptr;
    if (hp) {
      g.A2 = hp; g.lda2 = C; g.K = 2 * C; g.K1 = C;
    }
    LEOD_TRY(gemm_nt(h, g, st));
    LEOD_TRY(lstm_pointwise_fwd(dt, gates, c_prev ? c_prev[s] : nullptr, h_out[s], c_out[s], M, C, st));
  }
  return 0;
}

extern "C" int leod_backbone_step_bwd(leod_backbone_t *h, const void *x, int x_dtype, int x_h, int x_w, int B,
                                      const void *const h_prev[4], const void *const c_prev[4], const void *const h_out[4],
                                      const void *const c_out[4], const void *save, const void *const dh_out[4], const void *const dc_out[4],
                                      void *const dh_prev[4], void *const dc_prev[4], void *stream) {
  LEOD_REQUIRE(h && x && save && h_out && c_out && dc_prev && B > 0, "leod_backbone_step_bwd: bad argument");
  LEOD_REQUIRE(h->params && h->grads, "leod_backbone_step_bwd: parameter/gradient buffers not bound");
  cudaStream_t st = (cudaStream_t)stream;
  LEOD_TRY(ensure_workspace(h, B));
  const int dt = h->cfg.dtype;
  const float eps = h->cfg.ln_eps;
  const float *P = h->params;
  float *G = h->grads;
  SaveOff off[4];
  int64_t save_elems;
  compute_save_layout(h, B, off, &save_elems);
  const void *sv = save;
  bool have_hint = false;  // gradient flowing down from stage s+1's conv into h_out[s]
  for (int s = 3; s >= 0; --s) {
    const StageD &d = h->d[s];
    const StageP &p = h->p[s];
    const StageW &w = h->w[s];
    const int M = B * d.Ho * d.Wo, C = d.C, R = h->cfg.mlp_ratio;
    const SaveOff &o = off[s];
    const void *gates = eoff(h, sv, o.gates);
    const void *x2last = eoff(h, sv, o.blk[1].x2);
    const void *hp = h_prev ? h_prev[s] : nullptr;
    void *dgates = h->ws_big4;
    // ---- LSTM
    LEOD_TRY(lstm_pointwise_bwd(dt, gates, c_prev ? c_prev[s] : nullptr, c_out[s], dh_out ? dh_out[s] : nullptr,
                                have_hint ? h->ws_hint : nullptr, dc_out ? dc_out[s] : nullptr, dgates, dc_prev[s], M, C, st));
    LEOD_TRY(gemm_tn(h, dgates, 4 * C, x2last, C, G + p.lstmw, 2 * C, G + p.lstmb, M, 4 * C, C, st));
    if (hp) LEOD_TRY(gemm_tn(h, dgates, 4 * C, hp, C, G + p.lstmw + C, 2 * C, nullptr, M, 4 * C, C, st));
    void *dy = h->ws_a, *dy1 = h->ws_b, *dxn = h->ws_c;
    LEOD_TRY(gemm_nt(h, mk(dgates, 4 * C, w.WlT, 4 * C, dy, C, M, C, 4 * C), st));
    if (dh_prev && dh_prev[s])
      LEOD_TRY(gemm_nt(h, mk(dgates, 4 * C, eoff(h, w.WlT, (int64_t)C * 4 * C), 4 * C, dh_prev[s], C, M, C, 4 * C), st));
    // ---- attention blocks, grid then window
    for (int b = 1; b >= 0; --b) {
      const BlockP &q = p.blk[b];
      const BlockW &bw = w.blk[b];
      const void *qkv = eoff(h, sv, o.blk[b].qkv), *att = eoff(h, sv, o.blk[b].att), *x1 = eoff(h, sv, o.blk[b].x1);
      const void *xn2 = eoff(h, sv, o.blk[b].xn2), *u = eoff(h, sv, o.blk[b].u), *a = eoff(h, sv, o.blk[b].a);
      const void *xin = (b == 1) ? eoff(h, sv, o.blk[0].x2) : eoff(h, sv, o.x0);
      void *du = h->ws_big4, *dqkv = h->ws_qkv;
      // MLP: x2 = x1 + W2'(gelu(W1 xn2 + b1)) + b2'
      LEOD_TRY(gemm_tn(h, dy, C, a, R * C, bw.G2, R * C, bw.s2, M, C, R * C, st));
      LEOD_TRY(gemm_nt(h, mk(dy, C, bw.W2T, C, du, R * C, M, R * C, C, nullptr, EPI_GELU_BWD, nullptr, 0, (void *)u, R * C), st));
      LEOD_TRY(gemm_tn(h, du, R * C, xn2, C, G + q.fc1w, C, G + q.fc1b, M, R * C, C, st));
      LEOD_TRY(gemm_nt(h, mk(du, R * C, bw.W1T, R * C, dxn, C, M, C, R * C), st));
      LEOD_TRY(layernorm_bwd(dt, x1, P + q.n2w, dxn, dy, dy1, G + q.n2w, G + q.n2b, M, C, eps, st));
      // attention: x1 = xin + Wp'(attn(Wqkv LN1(xin))) + bp'
      LEOD_TRY(gemm_tn(h, dy1, C, att, C, bw.Gproj, C, bw.sproj, M, C, C, st));
      LEOD_TRY(gemm_nt(h, mk(dy1, C, bw.WprojT, C, dxn, C, M, C, C), st));
      LEOD_TRY(attention_bwd(dt, qkv, dxn, dqkv, B, d.Ho, d.Wo, C, h->cfg.dim_head, h->cfg.part_h, h->cfg.part_w, b == 0, st));
      if (b == 1) {
        const void *xn1 = eoff(h, sv, o.blk[b].xn1);
        LEOD_TRY(gemm_tn(h, dqkv, 3 * C, xn1, C, G + q.qkvw, C, G + q.qkvb, M, 3 * C, C, st));
        LEOD_TRY(gemm_nt(h, mk(dqkv, 3 * C, bw.WqkvT, 3 * C, dxn, C, M, C, 3 * C), st));
        LEOD_TRY(layernorm_bwd(dt, xin, P + q.n1w, dxn, dy1, dy, G + q.n1w, G + q.n1b, M, C, eps, st));
      } else {
        LEOD_TRY(gemm_tn(h, dqkv, 3 * C, xin, C, G + q.qkvw, C, G + q.qkvb, M, 3 * C, C, st));
        LEOD_TRY(gemm_nt(h, mk(dqkv, 3 * C, bw.WqkvT, 3 * C, dy, C, M, C, 3 * C, nullptr, EPI_RESID, dy1, C), st));
      }
    }
    // ---- downsample: x0 = LN(conv(in))
    const void *y0 = eoff(h, sv, o.y0);
    void *dy0 = h->ws_b;
    LEOD_TRY(layernorm_bwd(dt, y0, P + p.lnw, dy, nullptr, dy0, G + p.lnw, G + p.lnb, M, C, 1e-5f, st));
    if (s == 0) {
      LEOD_TRY(im2col_nchw(x_dtype, dt, x, h->ws_col, B, d.Cin, x_h, x_w, d.Hi, d.Wi, d.ksz, d.stride, d.pad, d.Kp, st));
      LEOD_TRY(gemm_tn(h, dy0, C, h->ws_col, d.Kp, G + p.convw, d.K, nullptr, M, C, d.K, st));
      have_hint = false;
    } else {
      // weight gradient in the (ky,kx,cin) patch order -> scratch, un-permuted in grads_finalize
      LEOD_TRY(im2col_nhwc(dt, h_out[s - 1], h->ws_col, B, d.Hi, d.Wi, d.Cin, d.ksz, d.stride, d.pad, d.Kp, st));
      LEOD_TRY(gemm_tn(h, dy0, C, h->ws_col, d.Kp, w.Gconv, d.K, nullptr, M, C, d.K, st));
      // input gradient: dcol = dy0 * Wconv, then gather-scatter back to the stage s-1 map
      LEOD_TRY(gemm_nt(h, mk(dy0, C, w.WconvT, C, h->ws_col, d.Kp, M, d.K, C), st));
      LEOD_TRY(col2im_nhwc(dt, h->ws_col, d.Kp, nullptr, h->ws_hint, B, d.Hi, d.Wi, d.Cin, d.ksz, d.stride, d.pad, st));
      have_hint = true;
    }
  }
  return 0;
}

extern "C" int leod_backbone_grads_finalize(leod_backbone_t *h, void *stream) {
  LEOD_REQUIRE(h && h->params && h->grads, "leod_backbone_grads_finalize: buffers not bound");
  cudaStream_t st = (cudaStream_t)stream;
  const float *P = h->params;
  float *G = h->grads;
  for (int s = 0; s < 4; ++s) {
    const StageD &d = h->d[s];
    const StageP &p = h->p[s];
    const StageW &w = h->w[s];
    const int C = d.C, R = h->cfg.mlp_ratio;
    if (s > 0) {
      const int64_t n = (int64_t)C * d.K;
      conv_grad_unpermute_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(w.Gconv, G + p.convw, C, d.Cin, d.ksz);
      LEOD_LAUNCH_CHECK();
    }
    for (int b = 0; b < 2; ++b) {
      const BlockP &q = p.blk[b];
      const BlockW &bw = w.blk[b];
      LEOD_TRY(layerscale_grad_finalize(bw.Gproj, bw.sproj, P + q.projw, P + q.projb, P + q.ls1, G + q.projw, G + q.projb,
                                        G + q.ls1, C, C, st));
      LEOD_TRY(layerscale_grad_finalize(bw.G2, bw.s2, P + q.fc2w, P + q.fc2b, P + q.ls2, G + q.fc2w, G + q.fc2b, G + q.ls2, C,
                                        R * C, st));
    }
  }
  return 0;
}
