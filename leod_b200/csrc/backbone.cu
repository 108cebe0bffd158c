// Host-side executor of the recurrent MaxViT backbone: owns the prepared weights, the workspace and
// the per-step kernel schedule (forward and backward of one timestep over the four stages).
// Reference semantics: models/detection/recurrent_backbone/maxvit_rnn.py:97-115, 182-201.
#include <algorithm>
#include <cstdlib>
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
// Activations kept for the backward pass (SB) and the output-gradient tensors the weight-gradient GEMMs read
// (GB).  Every entry is a matrix [rows, cols] in the operand dtype; in sequence mode rows = L * M_stage with the
// timestep as the slowest index, so one weight-gradient GEMM covers the whole BPTT window.
struct SB {
  void *y0, *x0, *xn1, *qkv[2], *att[2], *x1[2], *xn2[2], *u[2], *a[2], *x2[2], *gates;
};
struct GB {
  void *dgates, *dy2[2], *du[2], *dy1[2], *dqkv[2], *dy0;
};
struct StageLayout {  // element offsets of one stage inside an arena holding `slots` timesteps
  int64_t y0, x0, xn1, qkv[2], att[2], x1[2], xn2[2], u[2], a[2], x2[2], gates, c_all;
  int64_t dgates, dy2[2], du[2], dy1[2], dqkv[2], dy0;
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
  // workspace (sized for ws_rows0 = stage-0 rows of the largest batch of images processed at once)
  int64_t ws_imgs = 0;
  int ws_B = 0;
  void *ws_col = nullptr, *ws_dxn = nullptr, *ws_dx0 = nullptr, *ws_hint = nullptr, *ws_save = nullptr, *ws_stash = nullptr;
  int64_t ws_col_elems = 0;
  void *ws_dhc[4] = {nullptr, nullptr, nullptr, nullptr}, *ws_dcc[4] = {nullptr, nullptr, nullptr, nullptr};
  // sequence arena
  int seq_B = 0, seq_L = 0;
  void *seq_arena = nullptr;
  void *seq_col0 = nullptr;      // stem patch matrix of the whole window [L*B*Ho*Wo, Kp] (kept from forward to backward)
  int64_t seq_col0_imgs = 0;
  bool col0_live = false;
  cudaStream_t side[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};
  unsigned *seq_flags = nullptr;  // per-(tile, timestep) arrival counters of the fused recurrence kernels
  int fused_lstm = 1;
  int fused_bwd_min_tiles = 0;    // LEOD_FUSED_LSTM_BWD_MIN_TILES: stages with fewer 128-token tiles use per-step launches        // set for the duration of a sequence-mode forward/backward pair
  size_t esz() const { return cfg.dtype == LEOD_BF16 ? 2 : 4; }
  int gemm_impl = 0;  // 0 SIMT, 1 tensor core (bf16 only)
};

namespace {

int ensure_side_streams(leod_backbone *h) {
  if (h->ev_fork) return 0;
  for (int i = 0; i < 4; ++i) {
    LEOD_CUDA(cudaStreamCreateWithFlags(&h->side[i], cudaStreamNonBlocking));
    LEOD_CUDA(cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming));
  }
  LEOD_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  return 0;
}

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

// The stem runs as an implicit GEMM (kernels_stem.cu: operand tiles built on chip from the uint8 event tensor, no patch matrix)
// on the product path: bf16 + tensor-core GEMMs + uint8 input.  LEOD_IMPLICIT_STEM=0 restores the explicit patch matrix, which
// also serves the fp32 / SIMT / float-input paths and frame sizes that are not multiples of the 8 x 16 pixel tile.
bool implicit_stem(const leod_backbone *h, int s, const void *x, int x_dtype, int x_h, int x_w) {
  static const int enabled = [] {
    const char *e = getenv("LEOD_IMPLICIT_STEM");
    return e ? atoi(e) : 1;
  }();
  if (!enabled || s != 0 || h->cfg.dtype != LEOD_BF16 || h->gemm_impl != 1 || x_dtype != LEOD_U8) return false;
  const StageD &d = h->d[0];
  return d.ksz == 7 && d.stride == 4 && d.pad == 3 && d.Kp == d.Cin * 56 && stem_implicit_supported(d.Cin, x_h, x_w, d.Ho, d.Wo, d.C, x);
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
  pdl_prologue();
  const int K = Cin * ksz * ksz;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * K) return;
  const int n = (int)(idx / K), k = (int)(idx % K);
  const int cin = k % Cin, kx = (k / Cin) % ksz, ky = k / (Cin * ksz);
  dW[(((size_t)n * Cin + cin) * ksz + ky) * ksz + kx] += G[idx];
  G[idx] = 0.f;
}
// stem layout: k = (cin*ksz + ky)*8 + slot, slots 1..ksz <-> kx (kernels_elem.cu, im2col_nchw_kernel)
__global__ void stem_grad_unpermute_kernel(float *__restrict__ G, float *__restrict__ dW, int N, int Cin, int ksz) {
  pdl_prologue();
  const int K = Cin * ksz * 8;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * K) return;
  const int n = (int)(idx / K), k = (int)(idx % K);
  const int slot = k & 7, r = k >> 3;
  if (slot >= 1 && slot <= ksz) dW[(size_t)n * Cin * ksz * ksz + (size_t)r * ksz + (slot - 1)] += G[idx];
  G[idx] = 0.f;
}


// ------------------------------------------------------------------ arena layout
// slots = number of timesteps the arena holds; with_stash adds the gradient-side tensors; with_c adds c_all.
void compute_layout(const leod_backbone *h, int B, int slots, bool with_stash, bool with_c, StageLayout lay[4], int64_t *total) {
  int64_t cur = 0;
  auto take = [&](int64_t n) {
    const int64_t o = cur;
    cur += round_up(n, 128);
    return o;
  };
  const int64_t R = h->cfg.mlp_ratio;
  for (int s = 0; s < 4; ++s) {
    const StageD &d = h->d[s];
    const int64_t M = (int64_t)B * d.Ho * d.Wo * slots, C = d.C;
    StageLayout &l = lay[s];
    l.y0 = take(M * C);
    l.x0 = take(M * C);
    l.xn1 = take(M * C);
    for (int b = 0; b < 2; ++b) {
      l.qkv[b] = take(M * 3 * C);
      l.att[b] = take(M * C);
      l.x1[b] = take(M * C);
      l.xn2[b] = take(M * C);
      l.u[b] = take(M * R * C);
      l.a[b] = take(M * R * C);
      l.x2[b] = take(M * C);
    }
    l.gates = take(M * 4 * C);
    l.c_all = with_c ? take(M * C) : -1;
    if (with_stash) {
      l.dgates = take(M * 4 * C);
      for (int b = 0; b < 2; ++b) {
        l.dy2[b] = take(M * C);
        l.du[b] = take(M * R * C);
        l.dy1[b] = take(M * C);
        l.dqkv[b] = take(M * 3 * C);
      }
      l.dy0 = take(M * C);
    }
  }
  *total = cur;
}

// saved-activation pointers of stage s starting at timestep slot t
SB sb_at(const leod_backbone *h, const void *base, const StageLayout &l, int s, int B, int t) {
  const StageD &d = h->d[s];
  const int64_t M = (int64_t)B * d.Ho * d.Wo, C = d.C, R = h->cfg.mlp_ratio, e = (int64_t)h->esz();
  auto at = [&](int64_t off, int64_t cols) { return (void *)((char *)base + (off + t * M * cols) * e); };
  SB b;
  b.y0 = at(l.y0, C); b.x0 = at(l.x0, C); b.xn1 = at(l.xn1, C);
  for (int k = 0; k < 2; ++k) {
    b.qkv[k] = at(l.qkv[k], 3 * C); b.att[k] = at(l.att[k], C); b.x1[k] = at(l.x1[k], C); b.xn2[k] = at(l.xn2[k], C);
    b.u[k] = at(l.u[k], R * C); b.a[k] = at(l.a[k], R * C); b.x2[k] = at(l.x2[k], C);
  }
  b.gates = at(l.gates, 4 * C);
  return b;
}
GB gb_at(const leod_backbone *h, const void *base, const StageLayout &l, int s, int B, int t) {
  const StageD &d = h->d[s];
  const int64_t M = (int64_t)B * d.Ho * d.Wo, C = d.C, R = h->cfg.mlp_ratio, e = (int64_t)h->esz();
  auto at = [&](int64_t off, int64_t cols) { return (void *)((char *)base + (off + t * M * cols) * e); };
  GB g;
  g.dgates = at(l.dgates, 4 * C);
  for (int k = 0; k < 2; ++k) {
    g.dy2[k] = at(l.dy2[k], C); g.du[k] = at(l.du[k], R * C); g.dy1[k] = at(l.dy1[k], C); g.dqkv[k] = at(l.dqkv[k], 3 * C);
  }
  g.dy0 = at(l.dy0, C);
  return g;
}

void free_workspace(leod_backbone *h) {
  for (void **p : {&h->ws_col, &h->ws_dxn, &h->ws_dx0, &h->ws_hint, &h->ws_save, &h->ws_stash}) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  for (int s = 0; s < 4; ++s) {
    if (h->ws_dhc[s]) cudaFree(h->ws_dhc[s]);
    if (h->ws_dcc[s]) cudaFree(h->ws_dcc[s]);
    h->ws_dhc[s] = h->ws_dcc[s] = nullptr;
  }
  h->ws_imgs = 0;
  h->ws_B = 0;
}

// Workspace for processing up to `imgs` images of any stage in one go (imgs = B per-step, L*B batched).
int ensure_workspace(leod_backbone *h, int64_t imgs, int B) {
  if (imgs <= h->ws_imgs && B <= h->ws_B) return 0;
  imgs = std::max<int64_t>(imgs, h->ws_imgs);
  B = std::max(B, h->ws_B);
  free_workspace(h);
  const int64_t e = (int64_t)h->esz();
  int64_t mc = 0, col = 0;
  for (int s = 0; s < 4; ++s) {
    const StageD &d = h->d[s];
    mc = std::max<int64_t>(mc, imgs * d.Ho * d.Wo * d.C);
    // patch matrix: whole-batch for a single step, capped (chunked) for long batched runs
    col = std::max<int64_t>(col, std::min<int64_t>(imgs, std::max<int64_t>(B, 32)) * d.Ho * d.Wo * d.Kp);
  }
  h->ws_col_elems = col;
  LEOD_CUDA(cudaMalloc(&h->ws_col, col * e));
  LEOD_CUDA(cudaMalloc(&h->ws_dxn, mc * e));
  LEOD_CUDA(cudaMalloc(&h->ws_dx0, mc * e));
  LEOD_CUDA(cudaMalloc(&h->ws_hint, mc * e));
  StageLayout lay[4];
  int64_t n;
  compute_layout(h, B, 1, true, false, lay, &n);   // one-slot save (inference) + stash (per-step backward)
  LEOD_CUDA(cudaMalloc(&h->ws_save, n * e));
  LEOD_CUDA(cudaMalloc(&h->ws_stash, n * e));
  for (int s = 0; s < 4; ++s) {
    const StageD &d = h->d[s];
    const int64_t m = (int64_t)B * d.Ho * d.Wo * d.C;
    LEOD_CUDA(cudaMalloc(&h->ws_dhc[s], m * e));
    LEOD_CUDA(cudaMalloc(&h->ws_dcc[s], m * e));
  }
  h->ws_imgs = imgs;
  h->ws_B = B;
  return 0;
}

// ------------------------------------------------------------------ stage pieces (shared by step and sequence mode)
// conv downsample + LN + window block + grid block for `nimg` images.  `in`: stage input (x for stage 0, else the
// previous stage's hidden state, channels-last).
int front_fwd(leod_backbone *h, int s, int64_t nimg, const void *in, int x_dtype, int x_h, int x_w, const SB &b, cudaStream_t st) {
  const StageD &d = h->d[s];
  const StageP &p = h->p[s];
  const StageW &w = h->w[s];
  const int dt = h->cfg.dtype, C = d.C, R = h->cfg.mlp_ratio;
  const float eps = h->cfg.ln_eps;
  const float *P = h->params;
  const int64_t rows_per_img = (int64_t)d.Ho * d.Wo, e = (int64_t)h->esz();
  const int64_t M64 = nimg * rows_per_img;
  // row indices are 32-bit in the kernels (element offsets are computed in 64 bits)
  LEOD_REQUIRE(M64 < (1LL << 31) - 1024, "stage %d: %lld rows exceed the 32-bit row indexing of the kernels", s, (long long)M64);
  const int M = (int)M64;
  const bool implicit = implicit_stem(h, s, in, x_dtype, x_h, x_w);
  if (implicit) {
    ProfScope ps(PK_GEMM_NT, 2.0 * M * C * d.Kp, (double)nimg * d.Cin * x_h * x_w + e * ((double)C * d.Kp + (double)M * C), st, M, C, d.Kp);
    LEOD_TRY(stem_fwd_tc((const uint8_t *)in, (int)nimg, d.Cin, x_h, x_w, d.Ho, d.Wo, C, w.WconvT, d.Kp, b.y0, st));
  }
  // the stem's patch matrix of a whole BPTT window is kept for the weight-gradient GEMM (seq_col0, sequence mode)
  const bool keep = s == 0 && h->col0_live && h->seq_col0 && nimg == h->seq_col0_imgs;
  const int64_t chunk_imgs = keep ? nimg : std::max<int64_t>(1, h->ws_col_elems / (rows_per_img * d.Kp));
  for (int64_t i0 = 0; i0 < nimg && !implicit; i0 += chunk_imgs) {
    const int n = (int)std::min<int64_t>(chunk_imgs, nimg - i0);
    void *colp = keep ? h->seq_col0 : h->ws_col;
    if (s == 0) {
      const char *xin = (const char *)in + i0 * d.Cin * x_h * x_w * (int64_t)dtype_size(x_dtype);
      LEOD_TRY(im2col_nchw(x_dtype, dt, xin, colp, n, d.Cin, x_h, x_w, d.Hi, d.Wi, d.ksz, d.stride, d.pad, d.Kp, st));
    } else {
      const char *xin = (const char *)in + i0 * d.Hi * d.Wi * d.Cin * e;
      LEOD_TRY(im2col_nhwc(dt, xin, colp, n, d.Hi, d.Wi, d.Cin, d.ksz, d.stride, d.pad, d.Kp, st));
    }
    LEOD_TRY(gemm_nt(h, mk(colp, d.Kp, w.Wconv, d.Kp, (char *)b.y0 + i0 * rows_per_img * C * e, C, (int)(n * rows_per_img), C, d.Kp), st));
  }
  LEOD_TRY(layernorm_fwd(dt, b.y0, P + p.lnw, P + p.lnb, b.x0, M, C, 1e-5f, st));
  const void *xin = b.x0;
  const int nimg_i = (int)nimg;
  for (int k = 0; k < 2; ++k) {
    const BlockP &q = p.blk[k];
    const BlockW &bw = w.blk[k];
    const void *ain = xin;
    if (k == 1) {
      LEOD_TRY(layernorm_fwd(dt, xin, P + q.n1w, P + q.n1b, b.xn1, M, C, eps, st));
      ain = b.xn1;
    }
    LEOD_TRY(gemm_nt(h, mk(ain, C, bw.Wqkv, C, b.qkv[k], 3 * C, M, 3 * C, C, P + q.qkvb), st));
    LEOD_TRY(attention_fwd(dt, b.qkv[k], b.att[k], nimg_i, d.Ho, d.Wo, C, h->cfg.dim_head, h->cfg.part_h, h->cfg.part_w, k == 0, st));
    LEOD_TRY(gemm_nt(h, mk(b.att[k], C, bw.Wproj, C, b.x1[k], C, M, C, C, bw.bproj, EPI_RESID, xin, C), st));
    LEOD_TRY(layernorm_fwd(dt, b.x1[k], P + q.n2w, P + q.n2b, b.xn2[k], M, C, eps, st));
    LEOD_TRY(gemm_nt(h, mk(b.xn2[k], C, bw.W1, C, b.a[k], R * C, M, R * C, C, P + q.fc1b, EPI_GELU, nullptr, 0, b.u[k], R * C), st));
    LEOD_TRY(gemm_nt(h, mk(b.a[k], R * C, bw.W2, R * C, b.x2[k], C, M, C, R * C, bw.b2, EPI_RESID, b.x1[k], C), st));
    xin = b.x2[k];
  }
  return 0;
}

int lstm_fwd(leod_backbone *h, int s, int M, const void *x2, const void *h_prev, const void *c_prev, void *gates, void *h_out,
             void *c_out, cudaStream_t st) {
  const StageD &d = h->d[s];
  const int C = d.C;
  GemmNT g = mk(x2, C, h->w[s].Wl, 2 * C, gates, 4 * C, M, 4 * C, C, h->params + h->p[s].lstmb);
  if (h_prev) {
    g.A2 = h_prev; g.lda2 = C; g.K = 2 * C; g.K1 = C;
  }
  LEOD_TRY(gemm_nt(h, g, st));
  return lstm_pointwise_fwd(h->cfg.dtype, gates, c_prev, h_out, c_out, M, C, st);
}

// gate math backward + the two input-gradient GEMMs.  dx2 -> g.dy2[1]; dh_prev written when non-null.
int lstm_bwd(leod_backbone *h, int s, int M, const void *gates, const void *c_prev, const void *c_out, const void *dh_a,
             const void *dh_b, const void *dh_c, const void *dc, const GB &g, void *dc_prev, void *dh_prev, cudaStream_t st) {
  const StageD &d = h->d[s];
  const int C = d.C;
  const StageW &w = h->w[s];
  LEOD_TRY(lstm_pointwise_bwd(h->cfg.dtype, gates, c_prev, c_out, dh_a, dh_b, dc, g.dgates, dc_prev, M, C, st, dh_c));
  LEOD_TRY(gemm_nt(h, mk(g.dgates, 4 * C, w.WlT, 4 * C, g.dy2[1], C, M, C, 4 * C), st));
  if (dh_prev) LEOD_TRY(gemm_nt(h, mk(g.dgates, 4 * C, eoff(h, w.WlT, (int64_t)C * 4 * C), 4 * C, dh_prev, C, M, C, 4 * C), st));
  return 0;
}

// input gradients of the two blocks and the downsample; output-gradient tensors land in g for the deferred
// weight-gradient GEMMs.  hint_out (stage > 0): gradient w.r.t. the stage input (previous stage's h), channels-last.
int front_bwd(leod_backbone *h, int s, int64_t nimg, const SB &b, const GB &g, void *hint_out, cudaStream_t st,
              const void *hint_res = nullptr /* added to the hint: the caller's own gradient w.r.t. the previous stage's h */) {
  const StageD &d = h->d[s];
  const StageP &p = h->p[s];
  const StageW &w = h->w[s];
  const int dt = h->cfg.dtype, C = d.C, R = h->cfg.mlp_ratio;
  const float eps = h->cfg.ln_eps;
  const float *P = h->params;
  float *G = h->grads;
  const int64_t rows_per_img = (int64_t)d.Ho * d.Wo, e = (int64_t)h->esz();
  const int M = (int)(nimg * rows_per_img);
  void *dxn = h->ws_dxn;
  for (int k = 1; k >= 0; --k) {
    const BlockP &q = p.blk[k];
    const BlockW &bw = w.blk[k];
    const void *xin = (k == 1) ? b.x2[0] : b.x0;
    // MLP: x2 = x1 + W2'(gelu(W1 LN2(x1) + b1)) + b2'
    LEOD_TRY(gemm_nt(h, mk(g.dy2[k], C, bw.W2T, C, g.du[k], R * C, M, R * C, C, nullptr, EPI_GELU_BWD, nullptr, 0, b.u[k], R * C), st));
    LEOD_TRY(gemm_nt(h, mk(g.du[k], R * C, bw.W1T, R * C, dxn, C, M, C, R * C), st));
    LEOD_TRY(layernorm_bwd(dt, b.x1[k], P + q.n2w, dxn, g.dy2[k], g.dy1[k], G + q.n2w, G + q.n2b, M, C, eps, st));
    // attention: x1 = xin + Wp'(attn(Wqkv LN1(xin))) + bp'
    LEOD_TRY(gemm_nt(h, mk(g.dy1[k], C, bw.WprojT, C, dxn, C, M, C, C), st));
    LEOD_TRY(attention_bwd(dt, b.qkv[k], dxn, g.dqkv[k], (int)nimg, d.Ho, d.Wo, C, h->cfg.dim_head, h->cfg.part_h, h->cfg.part_w, k == 0, st));
    if (k == 1) {
      LEOD_TRY(gemm_nt(h, mk(g.dqkv[k], 3 * C, bw.WqkvT, 3 * C, dxn, C, M, C, 3 * C), st));
      LEOD_TRY(layernorm_bwd(dt, xin, P + q.n1w, dxn, g.dy1[k], g.dy2[0], G + q.n1w, G + q.n1b, M, C, eps, st));
    } else {
      LEOD_TRY(gemm_nt(h, mk(g.dqkv[k], 3 * C, bw.WqkvT, 3 * C, h->ws_dx0, C, M, C, 3 * C, nullptr, EPI_RESID, g.dy1[k], C), st));
    }
  }
  LEOD_TRY(layernorm_bwd(dt, b.y0, P + p.lnw, h->ws_dx0, nullptr, g.dy0, G + p.lnw, G + p.lnb, M, C, 1e-5f, st));
  if (s > 0 && hint_out) {
    const int64_t chunk_imgs = std::max<int64_t>(1, h->ws_col_elems / (rows_per_img * d.Kp));
    for (int64_t i0 = 0; i0 < nimg; i0 += chunk_imgs) {
      const int n = (int)std::min<int64_t>(chunk_imgs, nimg - i0);
      LEOD_TRY(gemm_nt(h, mk((char *)g.dy0 + i0 * rows_per_img * C * e, C, w.WconvT, C, h->ws_col, d.Kp, (int)(n * rows_per_img), d.K, C), st));
      LEOD_TRY(col2im_nhwc(dt, h->ws_col, d.Kp, hint_res ? (const char *)hint_res + i0 * d.Hi * d.Wi * d.Cin * e : nullptr,
                           (char *)hint_out + i0 * d.Hi * d.Wi * d.Cin * e, n, d.Hi, d.Wi, d.Cin, d.ksz, d.stride, d.pad, st));
    }
  }
  return 0;
}

// weight gradients of one stage over `nimg` images worth of rows (all timesteps at once in sequence mode).
// The LSTM's hidden-state half is paired by the caller (it needs h_{t-1}).
int stage_wgrads(leod_backbone *h, int s, int64_t nimg, const void *in, int x_dtype, int x_h, int x_w, const SB &b, const GB &g,
                 cudaStream_t st, int parts = 3 /* bit 0: linear layers, bit 1: downsample conv (uses the shared patch workspace) */) {
  const StageD &d = h->d[s];
  const StageP &p = h->p[s];
  const StageW &w = h->w[s];
  const int dt = h->cfg.dtype, C = d.C, R = h->cfg.mlp_ratio;
  float *G = h->grads;
  const int64_t rows_per_img = (int64_t)d.Ho * d.Wo, e = (int64_t)h->esz();
  const int M = (int)(nimg * rows_per_img);
  if (parts & 1) {
    LEOD_TRY(gemm_tn(h, g.dgates, 4 * C, b.x2[1], C, G + p.lstmw, 2 * C, G + p.lstmb, M, 4 * C, C, st));
    for (int k = 0; k < 2; ++k) {
      const BlockP &q = p.blk[k];
      const BlockW &bw = w.blk[k];
      LEOD_TRY(gemm_tn(h, g.dy2[k], C, b.a[k], R * C, bw.G2, R * C, bw.s2, M, C, R * C, st));
      LEOD_TRY(gemm_tn(h, g.du[k], R * C, b.xn2[k], C, G + q.fc1w, C, G + q.fc1b, M, R * C, C, st));
      LEOD_TRY(gemm_tn(h, g.dy1[k], C, b.att[k], C, bw.Gproj, C, bw.sproj, M, C, C, st));
      LEOD_TRY(gemm_tn(h, g.dqkv[k], 3 * C, k == 1 ? b.xn1 : b.x0, C, G + q.qkvw, C, G + q.qkvb, M, 3 * C, C, st));
    }
  }
  if (!(parts & 2)) return 0;
  if (implicit_stem(h, s, in, x_dtype, x_h, x_w)) {
    ProfScope ps(PK_GEMM_TN, 2.0 * M * C * d.K, 2.0 * (double)nimg * d.Cin * x_h * x_w + 2.0 * e * (double)M * C + 8.0 * C * d.K, st, M, C, d.K);
    return stem_wgrad_tc((const uint8_t *)in, (int)nimg, d.Cin, x_h, x_w, d.Ho, d.Wo, C, g.dy0, w.Gconv, d.K, st);
  }
  const bool keep = s == 0 && h->col0_live && h->seq_col0 && nimg == h->seq_col0_imgs;   // patches still there from the forward pass
  const int64_t chunk_imgs = keep ? nimg : std::max<int64_t>(1, h->ws_col_elems / (rows_per_img * d.Kp));
  for (int64_t i0 = 0; i0 < nimg; i0 += chunk_imgs) {
    const int n = (int)std::min<int64_t>(chunk_imgs, nimg - i0);
    void *colp = keep ? h->seq_col0 : h->ws_col;
    if (keep) {
    } else if (s == 0) {
      const char *xin = (const char *)in + i0 * d.Cin * x_h * x_w * (int64_t)dtype_size(x_dtype);
      LEOD_TRY(im2col_nchw(x_dtype, dt, xin, colp, n, d.Cin, x_h, x_w, d.Hi, d.Wi, d.ksz, d.stride, d.pad, d.Kp, st));
    } else {
      const char *xin = (const char *)in + i0 * d.Hi * d.Wi * d.Cin * e;
      LEOD_TRY(im2col_nhwc(dt, xin, colp, n, d.Hi, d.Wi, d.Cin, d.ksz, d.stride, d.pad, d.Kp, st));
    }
    // patch layouts differ from the parameter's (cin,ky,kx) order -> scratch, unpermuted by leod_backbone_grads_finalize
    LEOD_TRY(gemm_tn(h, (char *)g.dy0 + i0 * rows_per_img * C * e, C, colp, d.Kp, w.Gconv, d.K, nullptr,
                     (int)(n * rows_per_img), C, d.K, st));
  }
  return 0;
}

int check_common(const leod_backbone *h, const void *x, int x_h, int x_w, int B, const char *who) {
  LEOD_REQUIRE(h && x && B > 0, "%s: bad argument", who);
  LEOD_REQUIRE(h->params, "%s: parameters not bound", who);
  LEOD_REQUIRE(x_h <= h->cfg.in_h && x_w <= h->cfg.in_w && x_h > 0 && x_w > 0, "%s: input %dx%d does not fit the padded resolution %dx%d",
               who, x_h, x_w, h->cfg.in_h, h->cfg.in_w);
  return 0;
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
  if (const char *ev = getenv("LEOD_FUSED_LSTM_BWD_MIN_TILES")) h->fused_bwd_min_tiles = atoi(ev);
  if (const char *ev = getenv("LEOD_FUSED_LSTM")) h->fused_lstm = atoi(ev);
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
    d.K = s == 0 ? Cin * d.ksz * 8 : Cin * d.ksz * d.ksz;   // the stem pads every kernel row to 8 taps (im2col_nchw_kernel)
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
  free_workspace(h);
  if (h->seq_arena) cudaFree(h->seq_arena);
  if (h->seq_col0) cudaFree(h->seq_col0);
  if (h->seq_flags) cudaFree(h->seq_flags);
  for (int i = 0; i < 4; ++i) {
    if (h->side[i]) cudaStreamDestroy(h->side[i]);
    if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
  }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
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
  // 57 tensors -> two launches (items travel in the kernel parameters, PREP_BATCH_MAX per launch)
  PrepBatch pb;
  int blocks = 0;
  auto add = [&](const float *src, const float *scale, void *dst, int ldd, void *dstT, int lddT, int N, int K, int perm, int Cin, int ksz,
                 float *bias_out) -> int {
    if (pb.n == PREP_BATCH_MAX) {
      LEOD_TRY(prep_batch_launch(dt, pb, blocks, st));
      pb.n = 0;
      blocks = 0;
    }
    return prep_batch_add(pb, blocks, src, scale, dst, ldd, dstT, lddT, N, K, perm, Cin, ksz, bias_out);
  };
  for (int s = 0; s < 4; ++s) {
    const StageD &d = h->d[s];
    const int C = d.C, R = h->cfg.mlp_ratio;
    const StageP &p = h->p[s];
    StageW &w = h->w[s];
    LEOD_TRY(add(P + p.convw, nullptr, w.Wconv, d.Kp, w.WconvT, C, C, d.K, s == 0 ? 2 : 1, d.Cin, d.ksz, nullptr));
    for (int b = 0; b < 2; ++b) {
      const BlockP &q = p.blk[b];
      BlockW &bw = w.blk[b];
      LEOD_TRY(add(P + q.qkvw, nullptr, bw.Wqkv, C, bw.WqkvT, 3 * C, 3 * C, C, 0, 0, 0, nullptr));
      LEOD_TRY(add(P + q.projw, P + q.ls1, bw.Wproj, C, bw.WprojT, C, C, C, 0, 0, 0, nullptr));
      LEOD_TRY(add(P + q.projb, P + q.ls1, nullptr, 0, nullptr, 0, C, 1, 0, 0, 0, bw.bproj));
      LEOD_TRY(add(P + q.fc1w, nullptr, bw.W1, C, bw.W1T, R * C, R * C, C, 0, 0, 0, nullptr));
      LEOD_TRY(add(P + q.fc2w, P + q.ls2, bw.W2, R * C, bw.W2T, C, C, R * C, 0, 0, 0, nullptr));
      LEOD_TRY(add(P + q.fc2b, P + q.ls2, nullptr, 0, nullptr, 0, C, 1, 0, 0, 0, bw.b2));
    }
    LEOD_TRY(add(P + p.lstmw, nullptr, w.Wl, 2 * C, w.WlT, 4 * C, 4 * C, 2 * C, 0, 0, 0, nullptr));
  }
  LEOD_TRY(prep_batch_launch(dt, pb, blocks, st));
  // the stem has no input gradient, so its transposed copy is unused: it holds the FP16 copy the implicit-GEMM forward reads
  if (dt == LEOD_BF16) LEOD_TRY(stem_weight_to_f16(h->w[0].Wconv, h->w[0].WconvT, (int64_t)h->d[0].C * h->d[0].Kp, st));
  return 0;
}

extern "C" int64_t leod_backbone_save_bytes(const leod_backbone_t *h, int B) {
  if (!h || B <= 0) return -1;
  StageLayout lay[4];
  int64_t elems;
  compute_layout(h, B, 1, false, false, lay, &elems);
  return elems * (int64_t)h->esz();
}

extern "C" int leod_backbone_reserve(leod_backbone_t *h, int B) {
  LEOD_REQUIRE(h && B > 0, "leod_backbone_reserve: bad argument");
  return ensure_workspace(h, B, B);
}

// ------------------------------------------------------------------ per-timestep API
extern "C" int leod_backbone_step_fwd(leod_backbone_t *h, const void *x, int x_dtype, int x_h, int x_w, int B,
                                      const void *const h_prev[4], const void *const c_prev[4], void *const h_out[4],
                                      void *const c_out[4], void *save, void *stream) {
  LEOD_TRY(check_common(h, x, x_h, x_w, B, "leod_backbone_step_fwd"));
  LEOD_REQUIRE(h_out && c_out, "leod_backbone_step_fwd: null output arrays");
  h->col0_live = false;
  cudaStream_t st = (cudaStream_t)stream;
  LEOD_TRY(ensure_workspace(h, B, B));
  StageLayout lay[4];
  int64_t n;
  compute_layout(h, B, 1, false, false, lay, &n);
  void *sv = save ? save : h->ws_save;
  for (int s = 0; s < 4; ++s) {
    const StageD &d = h->d[s];
    const int M = B * d.Ho * d.Wo;
    const SB b = sb_at(h, sv, lay[s], s, B, 0);
    LEOD_TRY(front_fwd(h, s, B, s == 0 ? x : h_out[s - 1], x_dtype, x_h, x_w, b, st));
    LEOD_TRY(lstm_fwd(h, s, M, b.x2[1], h_prev ? h_prev[s] : nullptr, c_prev ? c_prev[s] : nullptr, b.gates, h_out[s], c_out[s], st));
  }
  return 0;
}

extern "C" int leod_backbone_step_bwd(leod_backbone_t *h, const void *x, int x_dtype, int x_h, int x_w, int B,
                                      const void *const h_prev[4], const void *const c_prev[4], const void *const h_out[4],
                                      const void *const c_out[4], const void *save, const void *const dh_out[4],
                                      const void *const dc_out[4], void *const dh_prev[4], void *const dc_prev[4], void *stream) {
  LEOD_TRY(check_common(h, x, x_h, x_w, B, "leod_backbone_step_bwd"));
  LEOD_REQUIRE(save && h_out && c_out && dc_prev, "leod_backbone_step_bwd: null argument");
  LEOD_REQUIRE(h->grads, "leod_backbone_step_bwd: gradient buffer not bound");
  h->col0_live = false;
  cudaStream_t st = (cudaStream_t)stream;
  LEOD_TRY(ensure_workspace(h, B, B));
  StageLayout lay[4], glay[4];
  int64_t n;
  compute_layout(h, B, 1, false, false, lay, &n);
  compute_layout(h, B, 1, true, false, glay, &n);
  bool have_hint = false;
  for (int s = 3; s >= 0; --s) {
    const StageD &d = h->d[s];
    const int M = B * d.Ho * d.Wo, C = d.C;
    const SB b = sb_at(h, save, lay[s], s, B, 0);
    const GB g = gb_at(h, h->ws_stash, glay[s], s, B, 0);
    const void *hp = h_prev ? h_prev[s] : nullptr;
    LEOD_TRY(lstm_bwd(h, s, M, b.gates, c_prev ? c_prev[s] : nullptr, c_out[s], dh_out ? dh_out[s] : nullptr,
                      have_hint ? h->ws_hint : nullptr, nullptr, dc_out ? dc_out[s] : nullptr, g, dc_prev[s],
                      (dh_prev && dh_prev[s]) ? dh_prev[s] : nullptr, st));
    LEOD_TRY(front_bwd(h, s, B, b, g, s > 0 ? h->ws_hint : nullptr, st));
    have_hint = s > 0;
    LEOD_TRY(stage_wgrads(h, s, B, s == 0 ? x : h_out[s - 1], x_dtype, x_h, x_w, b, g, st));
    if (hp) LEOD_TRY(gemm_tn(h, g.dgates, 4 * C, hp, C, h->grads + h->p[s].lstmw + C, 2 * C, nullptr, M, 4 * C, C, st));
  }
  return 0;
}

// ------------------------------------------------------------------ sequence (BPTT window) API
static int ensure_seq(leod_backbone *h, int B, int L) {
  LEOD_TRY(ensure_workspace(h, (int64_t)B * L, B));
  if (h->seq_arena && h->seq_B == B && h->seq_L == L) return 0;
  if (h->seq_arena) cudaFree(h->seq_arena);
  if (h->seq_col0) cudaFree(h->seq_col0);
  if (h->seq_flags) cudaFree(h->seq_flags);
  h->seq_arena = h->seq_col0 = nullptr;
  h->seq_flags = nullptr;
  LEOD_CUDA(cudaMalloc((void **)&h->seq_flags, sizeof(unsigned) * (size_t)(((int64_t)B * h->d[0].Ho * h->d[0].Wo + 127) / 128) * L));
  h->seq_col0_imgs = 0;
  StageLayout lay[4];
  int64_t n;
  compute_layout(h, B, L, true, true, lay, &n);
  LEOD_CUDA(cudaMalloc(&h->seq_arena, n * (int64_t)h->esz()));
  {
    const StageD &d0 = h->d[0];
    const int64_t rows = (int64_t)B * L * d0.Ho * d0.Wo;
    // the product path gathers the stem's operands on chip (kernels_stem.cu): no patch matrix to keep
    const char *env = getenv("LEOD_IMPLICIT_STEM");
    const bool maybe_explicit = !(h->cfg.dtype == LEOD_BF16 && h->gemm_impl == 1) || (env && atoi(env) == 0) || d0.Ho % 8 != 0 || d0.Wo % 16 != 0;
    if (maybe_explicit && rows * d0.Kp < (1LL << 31) * 4 && rows < (1LL << 31) &&
        cudaMalloc(&h->seq_col0, rows * d0.Kp * (int64_t)h->esz()) == cudaSuccess) {
      h->seq_col0_imgs = (int64_t)B * L;
    } else {
      cudaGetLastError();   // not enough memory: fall back to chunked re-gathering
      h->seq_col0 = nullptr;
    }
  }
  h->seq_B = B;
  h->seq_L = L;
  return 0;
}

extern "C" int64_t leod_backbone_seq_arena_bytes(const leod_backbone_t *h, int B, int L) {
  if (!h || B <= 0 || L <= 0) return -1;
  StageLayout lay[4];
  int64_t n;
  compute_layout(h, B, L, true, true, lay, &n);
  return n * (int64_t)h->esz();
}

extern "C" int leod_backbone_seq_fwd(leod_backbone_t *h, const void *x, int x_dtype, int x_h, int x_w, int B, int L,
                                     const void *const h0[4], const void *const c0[4], void *const h_all[4], void *const c_last[4],
                                     void *stream) {
  LEOD_TRY(check_common(h, x, x_h, x_w, B, "leod_backbone_seq_fwd"));
  LEOD_REQUIRE(L > 0 && h_all && c_last, "leod_backbone_seq_fwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  LEOD_TRY(ensure_seq(h, B, L));
  StageLayout lay[4];
  int64_t n;
  compute_layout(h, B, L, true, true, lay, &n);
  const int64_t e = (int64_t)h->esz();
  h->col0_live = true;
  // Stage-major ("layer-wise") schedule.  Stage s at time t needs stage s-1 at time t and its own state at t-1; there
  // is no top-down feedback, so a whole stage can run over all L timesteps before the next one starts.  Everything but
  // the hidden-state half of the ConvLSTM gates is then batched over the L*B frames of the window; only
  // gates_t += W_h h_{t-1} and the gate math remain sequential.
  for (int s = 0; s < 4; ++s) {
    const StageD &d = h->d[s];
    const int64_t M = (int64_t)B * d.Ho * d.Wo, C = d.C;
    const SB b = sb_at(h, h->seq_arena, lay[s], s, B, 0);
    LEOD_TRY(front_fwd(h, s, (int64_t)B * L, s == 0 ? x : h_all[s - 1], x_dtype, x_h, x_w, b, st));
    // input half of the gates for every timestep: gates = x2 W_x^T + bias
    LEOD_TRY(gemm_nt(h, mk(b.x2[1], (int)C, h->w[s].Wl, (int)(2 * C), b.gates, (int)(4 * C), (int)(M * L), (int)(4 * C), (int)C,
                           h->params + h->p[s].lstmb), st));
    char *hs = (char *)h_all[s];
    char *cs = (char *)h->seq_arena + lay[s].c_all * e;
    if (h->cfg.dtype == LEOD_BF16 && h->gemm_impl == 1 && h->fused_lstm && C % 16 == 0) {
      // the whole recurrence of this stage in one launch (kernels_gemm_tc.cu, lstm_seq_fwd_kernel)
      ProfScope ps(PK_LSTM, 2.0 * M * L * 4 * C * C + 30.0 * M * L * C, 11.0 * M * L * C * e, st, (int)M, (int)C, L);
      LEOD_TRY(lstm_seq_fwd_tc(b.gates, eoff(h, h->w[s].Wl, C), (int)(2 * C), h0 ? h0[s] : nullptr, c0 ? c0[s] : nullptr, hs, cs,
                               h->seq_flags, (int)M, (int)C, L, st));
    } else {
      for (int t = 0; t < L; ++t) {
        const void *hp = t == 0 ? (h0 ? h0[s] : nullptr) : hs + (t - 1) * M * C * e;
        const void *cp = t == 0 ? (c0 ? c0[s] : nullptr) : cs + (t - 1) * M * C * e;
        void *gt = (char *)b.gates + t * M * 4 * C * e;
        if (hp)  // gates_t += h_{t-1} W_h^T (in place: every element is read and written by the same thread)
          LEOD_TRY(gemm_nt(h, mk(hp, (int)C, eoff(h, h->w[s].Wl, C), (int)(2 * C), gt, (int)(4 * C), (int)M, (int)(4 * C), (int)C, nullptr,
                                 EPI_RESID, gt, (int)(4 * C)), st));
        LEOD_TRY(lstm_pointwise_fwd(h->cfg.dtype, gt, cp, hs + t * M * C * e, cs + t * M * C * e, (int)M, (int)C, st));
      }
    }
  }
  for (int s = 0; s < 4; ++s) {
    const StageD &d = h->d[s];
    const int64_t M = (int64_t)B * d.Ho * d.Wo, C = d.C;
    LEOD_TRY(device_copy(c_last[s], (char *)h->seq_arena + (lay[s].c_all + (L - 1) * M * C) * e, (size_t)(M * C * e), st));
  }
  return 0;
}

extern "C" int leod_backbone_seq_bwd(leod_backbone_t *h, const void *x, int x_dtype, int x_h, int x_w, int B, int L,
                                     const void *const h0[4], const void *const c0[4], const void *const h_all[4],
                                     const void *const dh_all[4], const void *const dc_last[4], void *const dh0[4], void *const dc0[4],
                                     void *stream) {
  LEOD_TRY(check_common(h, x, x_h, x_w, B, "leod_backbone_seq_bwd"));
  LEOD_REQUIRE(L > 0 && h_all, "leod_backbone_seq_bwd: bad argument");
  LEOD_REQUIRE(h->grads, "leod_backbone_seq_bwd: gradient buffer not bound");
  LEOD_REQUIRE(h->seq_arena && h->seq_B == B && h->seq_L == L, "leod_backbone_seq_bwd: no matching leod_backbone_seq_fwd (B=%d L=%d)", B, L);
  cudaStream_t st = (cudaStream_t)stream;
  StageLayout lay[4];
  int64_t n;
  compute_layout(h, B, L, true, true, lay, &n);
  const int64_t e = (int64_t)h->esz();
  // Stage-major, mirroring the forward: per stage the ConvLSTM recurrence runs backwards over time (gate math +
  // dh_{t-1} = dgates_t W_h), then every other input-gradient kernel is batched over the L*B frames.
  for (int s = 3; s >= 0; --s) {
    const StageD &d = h->d[s];
    const int64_t M = (int64_t)B * d.Ho * d.Wo, C = d.C;
    const StageW &w = h->w[s];
    const SB b = sb_at(h, h->seq_arena, lay[s], s, B, 0);
    const GB g = gb_at(h, h->seq_arena, lay[s], s, B, 0);
    const char *cs = (const char *)h->seq_arena + lay[s].c_all * e;
    // external gradient w.r.t. h_all[s], all timesteps: the caller's (head) gradient, plus - for s < 3 - the gradient
    // through stage s+1's downsample, which front_bwd(s+1) has already added into ws_hint
    const char *dh_ext = s < 3 ? (const char *)h->ws_hint : ((dh_all && dh_all[s]) ? (const char *)dh_all[s] : nullptr);
    // The fused backward recurrence re-streams a 128 x 4C operand tile per step and CTA; stages with few token tiles are
    // split over output channels so that ~120 CTAs share a timestep (lstm_seq_bwd_tc).
    if (h->cfg.dtype == LEOD_BF16 && h->gemm_impl == 1 && h->fused_lstm && C % 16 == 0 && M >= (int64_t)h->fused_bwd_min_tiles * 128) {
      ProfScope ps(PK_LSTM, 2.0 * M * L * 4 * C * C + 30.0 * M * L * C, 13.0 * M * L * C * e, st, (int)M, (int)C, -L);
      const bool want_dh0 = h0 && h0[s] && dh0 && dh0[s];
      LEOD_TRY(lstm_seq_bwd_tc(b.gates, cs, c0 ? c0[s] : nullptr, dh_ext, dc_last ? dc_last[s] : nullptr, g.dgates, h->ws_dcc[s],
                               (dc0 && dc0[s]) ? dc0[s] : nullptr, want_dh0 ? dh0[s] : nullptr, eoff(h, w.WlT, C * 4 * C), (int)(4 * C),
                               h->seq_flags, (int)M, (int)C, L, st));
    } else {
      for (int t = L - 1; t >= 0; --t) {
        const void *cp = t == 0 ? (c0 ? c0[s] : nullptr) : cs + (t - 1) * M * C * e;
        const void *dh_e = dh_ext ? dh_ext + t * M * C * e : nullptr;
        const void *dh_next = t < L - 1 ? h->ws_dhc[s] : nullptr;            // from timestep t+1
        const void *dc_in = t < L - 1 ? h->ws_dcc[s] : (dc_last ? dc_last[s] : nullptr);
        const bool has_hp = t > 0 || (h0 && h0[s]);
        void *dc_out_ptr = (t == 0 && dc0 && dc0[s]) ? dc0[s] : h->ws_dcc[s];
        void *dh_out_ptr = !has_hp ? nullptr : ((t == 0) ? ((dh0 && dh0[s]) ? dh0[s] : nullptr) : h->ws_dhc[s]);
        void *dgt = (char *)g.dgates + t * M * 4 * C * e;
        LEOD_TRY(lstm_pointwise_bwd(h->cfg.dtype, (const char *)b.gates + t * M * 4 * C * e, cp, cs + t * M * C * e, dh_e, nullptr, dc_in,
                                    dgt, dc_out_ptr, (int)M, (int)C, st, dh_next));
        if (dh_out_ptr)
          LEOD_TRY(gemm_nt(h, mk(dgt, (int)(4 * C), eoff(h, w.WlT, C * 4 * C), (int)(4 * C), dh_out_ptr, (int)C, (int)M, (int)C, (int)(4 * C)), st));
      }
    }
    // d loss / d x2 for all timesteps, then the two blocks and the downsample
    LEOD_TRY(gemm_nt(h, mk(g.dgates, (int)(4 * C), w.WlT, (int)(4 * C), g.dy2[1], (int)C, (int)(M * L), (int)C, (int)(4 * C)), st));
    LEOD_TRY(front_bwd(h, s, (int64_t)B * L, b, g, s > 0 ? h->ws_hint : nullptr, st,
                       (s > 0 && dh_all && dh_all[s - 1]) ? dh_all[s - 1] : nullptr));
  }
  // weight gradients: one GEMM per layer over all L timesteps.  The ~50 GEMMs are independent of each other and many
  // are too small to fill 148 SMs, so the linear layers of every stage run on their own side stream; the downsample
  // convolutions share the patch workspace and stay in order on the caller's stream.  (Serial when the per-launch
  // profiler is on: its events assume one stream.)
  const bool fork = !leod_profiling_on();
  if (fork) LEOD_TRY(ensure_side_streams(h));
  if (fork) {
    LEOD_CUDA(cudaEventRecord(h->ev_fork, st));
    for (int s = 0; s < 4; ++s) LEOD_CUDA(cudaStreamWaitEvent(h->side[s], h->ev_fork, 0));
  }
  for (int s = 0; s < 4; ++s) {
    const StageD &d = h->d[s];
    const int64_t M = (int64_t)B * d.Ho * d.Wo, C = d.C;
    const SB b = sb_at(h, h->seq_arena, lay[s], s, B, 0);
    const GB g = gb_at(h, h->seq_arena, lay[s], s, B, 0);
    cudaStream_t ss = fork ? h->side[s] : st;
    LEOD_TRY(stage_wgrads(h, s, (int64_t)B * L, s == 0 ? x : h_all[s - 1], x_dtype, x_h, x_w, b, g, ss, 1));
    float *dWl_h = h->grads + h->p[s].lstmw + C;
    if (L > 1)   // dgates of timesteps 1..L-1 pair with the hidden states of timesteps 0..L-2
      LEOD_TRY(gemm_tn(h, (char *)g.dgates + M * 4 * C * e, 4 * C, h_all[s], C, dWl_h, 2 * C, nullptr, (int)((L - 1) * M), 4 * C, C, ss));
    if (h0 && h0[s]) LEOD_TRY(gemm_tn(h, g.dgates, 4 * C, h0[s], C, dWl_h, 2 * C, nullptr, (int)M, 4 * C, C, ss));
    LEOD_TRY(stage_wgrads(h, s, (int64_t)B * L, s == 0 ? x : h_all[s - 1], x_dtype, x_h, x_w, b, g, st, 2));
  }
  if (fork) {
    for (int s = 0; s < 4; ++s) {
      LEOD_CUDA(cudaEventRecord(h->ev_join[s], h->side[s]));
      LEOD_CUDA(cudaStreamWaitEvent(st, h->ev_join[s], 0));
    }
  }
  h->col0_live = false;
  return 0;
}

extern "C" int leod_backbone_grads_finalize(leod_backbone_t *h, void *stream) {
  LEOD_REQUIRE(h && h->params && h->grads, "leod_backbone_grads_finalize: buffers not bound");
  cudaStream_t st = (cudaStream_t)stream;
  const float *P = h->params;
  float *G = h->grads;
  LsBatch lsb;
  int ls_blocks = 0;
  for (int s = 0; s < 4; ++s) {
    const StageD &d = h->d[s];
    const StageP &p = h->p[s];
    const StageW &w = h->w[s];
    const int C = d.C, R = h->cfg.mlp_ratio;
    {
      const int64_t n = (int64_t)C * d.K;
      if (s > 0)
        LEOD_LAUNCH((conv_grad_unpermute_kernel), (int)((n + 255) / 256), 256, 0, st, w.Gconv, G + p.convw, C, d.Cin, d.ksz);
      else
        LEOD_LAUNCH((stem_grad_unpermute_kernel), (int)((n + 255) / 256), 256, 0, st, w.Gconv, G + p.convw, C, d.Cin, d.ksz);
      LEOD_LAUNCH_CHECK();
    }
    for (int b = 0; b < 2; ++b) {
      const BlockP &q = p.blk[b];
      const BlockW &bw = w.blk[b];
      LEOD_TRY(ls_batch_add(lsb, ls_blocks, bw.Gproj, bw.sproj, P + q.projw, P + q.projb, P + q.ls1, G + q.projw, G + q.projb, G + q.ls1, C, C));
      LEOD_TRY(ls_batch_add(lsb, ls_blocks, bw.G2, bw.s2, P + q.fc2w, P + q.fc2b, P + q.ls2, G + q.fc2w, G + q.fc2b, G + q.ls2, C, R * C));
    }
  }
  LEOD_TRY(ls_batch_launch(lsb, ls_blocks, st));     // the 16 LayerScale gradients of the backbone in one launch
  return 0;
}
