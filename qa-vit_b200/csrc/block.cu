// C-ABI entry points and the quad-block schedule (forward and backward) over the kernels in this directory.
// See include/qavit_b200.h for the contract and DESIGN.md for the data layout of `saved` / `scratch`.
#include <stdarg.h>
#include <string.h>

#include "../../include/qavit_b200.h"
#include "kernels.h"

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";
void qv_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* qavit_last_error(void) { return g_err; }
extern "C" int qavit_abi_version(void) { return QAVIT_ABI_VERSION; }
unsigned long long g_qv_launches = 0;
static int qv_env_flag(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}
int g_qv_pdl = qv_env_flag("QAVIT_PDL", 0);   // measured on the headline step: 46.5 ms without, 47.1-47.4 ms with (any entry order)
extern "C" long long qavit_launch_count(void) { return (long long)g_qv_launches; }

// ------------------------------------------------------------------------------------------------ names
namespace {
struct PName { const char* name; int scope; };
const PName kNames[QP_COUNT] = {
    {"norm1.weight", 0}, {"norm1.bias", 0},
    {"swa.qkv.weight", 0}, {"swa.qkv.bias", 0}, {"swa.linformer.E_k", 0}, {"swa.linformer.E_v", 0},
    {"swa.proj.weight", 0}, {"swa.proj.bias", 0}, {"swa.norm.weight", 0}, {"swa.norm.bias", 0},
    {"msda.qkv.weight", 0}, {"msda.qkv.bias", 0}, {"msda.linformer.E_k", 0}, {"msda.linformer.E_v", 0},
    {"msda.proj.weight", 0}, {"msda.proj.bias", 0}, {"msda.norm.weight", 0}, {"msda.norm.bias", 0},
    {"cga.q_proj.weight", 0}, {"cga.q_proj.bias", 0}, {"cga.k_proj.weight", 0}, {"cga.k_proj.bias", 0},
    {"cga.v_proj.weight", 0}, {"cga.v_proj.bias", 0}, {"cga.bank_k_proj.weight", 0}, {"cga.bank_k_proj.bias", 0},
    {"cga.bank_v_proj.weight", 0}, {"cga.bank_v_proj.bias", 0}, {"cga.proj.weight", 0}, {"cga.proj.bias", 0},
    {"cga.norm.weight", 0}, {"cga.norm.bias", 0},
    {"cross_attn.q_proj.weight", 0}, {"cross_attn.q_proj.bias", 0}, {"cross_attn.k_proj.weight", 0},
    {"cross_attn.k_proj.bias", 0}, {"cross_attn.v_proj.weight", 0}, {"cross_attn.v_proj.bias", 0},
    {"cross_attn.proj.weight", 0}, {"cross_attn.proj.bias", 0},
    {"norm_swa.weight", 0}, {"norm_swa.bias", 0}, {"norm_msda.weight", 0}, {"norm_msda.bias", 0},
    {"norm_cga.weight", 0}, {"norm_cga.bias", 0}, {"norm_cross.weight", 0}, {"norm_cross.bias", 0},
    {"compress_swa.weight", 0}, {"compress_swa.bias", 0}, {"compress_msda.weight", 0}, {"compress_msda.bias", 0},
    {"compress_cga.weight", 0}, {"compress_cga.bias", 0}, {"compress_cross.weight", 0}, {"compress_cross.bias", 0},
    {"fusion.fusion_weights", 0},
    {"bottleneck_mlp.fc1.weight", 0}, {"bottleneck_mlp.fc1.bias", 0}, {"bottleneck_mlp.fc2.weight", 0},
    {"bottleneck_mlp.fc2.bias", 0},
    {"norm2.weight", 0}, {"norm2.bias", 0},
    {"ccf_ffn.gamma", 0}, {"ccf_ffn.fc1.weight", 0}, {"ccf_ffn.fc1.bias", 0}, {"ccf_ffn.dwconv_norm.weight", 0},
    {"ccf_ffn.dwconv_norm.bias", 0}, {"ccf_ffn.dwconv.scale", 0}, {"ccf_ffn.dwconv.dwconv.weight", 0},
    {"ccf_ffn.dwconv.dwconv.bias", 0}, {"ccf_ffn.post_dwconv_norm.weight", 0}, {"ccf_ffn.post_dwconv_norm.bias", 0},
    {"ccf_ffn.fc2.weight", 0}, {"ccf_ffn.fc2.bias", 0},
    {"token_learner.attention.0.weight", 1}, {"token_learner.attention.0.bias", 1},
    {"token_learner.attention.1.weight", 1}, {"token_learner.attention.1.bias", 1},
    {"token_upmix.upsample_attn.weight", 1}, {"token_upmix.upsample_attn.bias", 1},
    {"token_upmix.norm.weight", 1}, {"token_upmix.norm.bias", 1},
    {"global_k", 2}, {"global_v", 2}, {"write_norm.weight", 2}, {"write_norm.bias", 2},
    {"write_compression.weight", 2}, {"write_compression.bias", 2}, {"write_gate.weight", 2}, {"write_gate.bias", 2},
};
}  // namespace

extern "C" const char* qavit_block_param_name(int index, int* scope) {
  if (index < 0 || index >= QP_COUNT) return nullptr;
  if (scope) *scope = kNames[index].scope;
  return kNames[index].name;
}

// ------------------------------------------------------------------------------------------------ GEMM dispatch
int gemm_nt(cudaStream_t s, int dt, const void* A, int lda, int M, const Weight& W, const GemmEpi& e) {
  if (dt == QV_BF16 && W.wb && tc_shape_ok_nt(M, W.N, W.K, lda))
    return tc_gemm_nt(s, (const bf16*)A, lda, M, W.N, W.K, W.wb, e);
  return simt_gemm_nt(s, dt, A, lda, M, W.N, W.K, W.w, e);
}
int gemm_nn(cudaStream_t s, int dt, const void* dY, int ldy, int M, const Weight& W, const GemmEpi& e) {
  // dX[M, K] = dY[M, N] W[N, K] = dY (W^T)^T : the NT kernel on the transposed bf16 copy
  if (dt == QV_BF16 && W.wbt && tc_shape_ok_nt(M, W.K, W.N, ldy))
    return tc_gemm_nt(s, (const bf16*)dY, ldy, M, W.K, W.N, W.wbt, e);
  return simt_gemm_nn(s, dt, dY, ldy, M, W.N, W.K, W.w, e);
}
int gemm_tn(cudaStream_t s, int dt, const void* dY, int ldy, const void* X, int ldx, int M, int N, int K, float* dW,
            float* db, const float* scale) {
  if (dt == QV_BF16 && tc_shape_ok_tn(M, N, K, ldy, ldx)) {
    return tc_gemm_tn(s, (const bf16*)dY, ldy, (const bf16*)X, ldx, M, N, K, dW, scale, 0, 0, db, 1, N);
  }
  if (dt == QV_BF16 && K > 256 && tc_shape_ok_tn(M, K, N, ldx, ldy)) {   // wide-K weight: accumulate X^T dY, store transposed
    return tc_gemm_tn(s, (const bf16*)X, ldx, (const bf16*)dY, ldy, M, K, N, dW, scale, 1, K, db, 2, N);
  }
  return simt_gemm_tn(s, dt, dY, ldy, X, ldx, M, N, K, dW, db, scale);
}

// ------------------------------------------------------------------------------------------------ layout
namespace {

struct Bump {
  size_t off = 0;
  size_t take(size_t bytes) {
    size_t r = off;
    off += (bytes + 255) & ~(size_t)255;
    return r;
  }
};

struct Dims {
  int B, Nt, Nf, No, R, Rf, Ro, d, H, hd, kb, G, cg, cpg, ws, side, klin, NM, Lm, cd, bh, fh, dt;
  size_t ts;
  bool tl, v1;
  bool ftok, fup;   // TokenLearner / TokenUpMix run as the fused split-precision kernels of tokens_fused.cu
  bool fcmp;        // branch LayerNorm + compress + fusion scale of all 4 branches as one kernel per direction (cmp_fused.cu)
  bool fmid;        // CCF-FFN GELU -> LayerNorm -> depthwise 3x3 -> LayerNorm as one kernel per direction (ffn_mid.cu)
  bool xstk;        // the three projections that read norm1's output -- swa.qkv (3d), msda.qkv's q rows (d), cross_attn.q_proj (d) -- run
                    // as ONE [R, d] x [5d, d]^T GEMM into one [R, 5d] buffer (consumers index it with a row pitch of 5d); backward: one
                    // dX GEMM over K = 5d and one dW GEMM whose rows scatter to the three parameters' gradients (tc_gemm_tn_seg)
  int ldp, ldq1;    // row pitch of the qkv_swa buffer / of the q_msda and q_cross buffers (5d when stacked, else 3d / d)
  int tdt;          // storage / GEMM type of the TokenLearner gate path (LN output, logits): fp32 whenever no tensor-core token
                    // kernel covers the shape (e.g. 576 -> 16 tokens of the 96 x 96 recipe) -- bf16 gate logits under a softmax over
                    // hundreds of tokens were the largest single bf16 error source of the model
  size_t tts;
};

// gemm weights that get bf16 copies in bf16 runs
enum { W_SWA_QKV, W_SWA_PROJ, W_MSDA_Q, W_MSDA_KV, W_MSDA_PROJ, W_CGA_PROJ, W_CROSS_Q, W_CROSS_PROJ, W_C0, W_C1, W_C2,
       W_C3, W_B1, W_B2, W_F1, W_F2, W_TL, W_WRITE, W_XSTK, W_COUNT };

struct Saved {  // byte offsets into `saved`
  size_t tl_stats, tl_ln, tl_S, tl_Z, xc, n1_stats, xn, alpha, snap[4], qkv_swa, attn_swa, xp, kv_msda, q_msda, attn_msda,
      attn_cga, kbp, vbp, q_cross, Kc, Vc, attn_cross, branch[4], nb_stats[4], nb[4], fused, h1_pre, h1, x1, n2_stats, y,
      h_pre, h, dn_stats, hn, cs, pd_stats, hn2, o, out_blk, up, up_stats, wb[W_COUNT], wbt[W_COUNT], wstack, bstack, bxstk, rng, rs, total;
};
struct Scratch {  // byte offsets into `scratch`
  size_t tl_logits, tn, cgbuf, partial, attn_ws_f, total_fwd;
  size_t d_up, d_blk, d_o, d_hn2, d_cs, d_hn, d_hpre, d_y, d_x1, d_x1t, d_h1, d_h1pre, d_fused, d_nb, d_branch, d_branch3[3], d_attn,
      d_qkv, d_kv, d_xp, d_q, d_qc, d_xn, dKc, dVc, dkbp, dvbp, draw, d_logits, d_tl_ln, attn_ws_b, total_bwd;
};

int msda_tokens(const qavit_block_cfg& c, int side) {
  int n = 0;
  for (int i = 0; i < c.n_dilations; ++i) {
    const int nd = (side + c.dilations[i] - 1) / c.dilations[i];
    n += nd * nd;
  }
  return (n - c.pool_stride) / c.pool_stride + 1;
}

int make_dims(const qavit_block_cfg& c, Dims* D) {
  D->B = c.batch; D->Nt = c.tokens; D->Nf = c.tokens_full; D->tl = c.token_learner != 0;
  D->R = c.batch * c.tokens; D->Rf = c.batch * c.tokens_full;
  D->No = (c.token_learner && c.tokens_out > 0) ? c.tokens_out : c.tokens_full;      // TokenUpMix output tokens
  D->Ro = c.batch * D->No;
  D->d = c.dim; D->H = c.heads; D->hd = c.dim / c.heads; D->kb = c.bank_size; D->G = c.groups;
  D->cg = c.dim / c.groups; D->cpg = (c.dim / 2) / c.groups; D->ws = c.window; D->klin = c.linformer_k;
  D->cd = c.compress_dim; D->bh = c.bottleneck_hidden; D->fh = c.ffn_hidden; D->dt = c.dtype; D->v1 = c.ffn_v1 != 0;
  D->ts = c.dtype == QV_BF16 ? 2 : 4;
  int side = 1;
  while (side * side < c.tokens) ++side;
  QV_CHECK(side * side == c.tokens, "tokens=%d is not a square grid", c.tokens);
  QV_CHECK(side % c.window == 0, "grid side %d %% window %d != 0 (the reference requires it, H:466)", side, c.window);
  QV_CHECK(c.dim % c.heads == 0 && c.dim % c.groups == 0 && (c.dim / 2) % c.groups == 0, "dim/heads/groups mismatch");
  QV_CHECK(c.dim % 8 == 0 && c.dim <= 256, "dim=%d unsupported (multiple of 8, <= 256)", c.dim);
  QV_CHECK(4 * c.compress_dim == c.dim, "4 * compress_dim must equal dim");
  QV_CHECK(c.dtype == QV_F32 || c.dtype == QV_BF16, "dtype %d", c.dtype);
  QV_CHECK(!c.token_learner || c.tokens_full >= c.tokens, "token learner: tokens_full < tokens");
  QV_CHECK(c.token_learner || c.tokens_full == c.tokens, "tokens_full must equal tokens without token learner");
  D->side = side;
  D->ftok = D->tl && c.dtype == QV_BF16 && tokens_fused_ok(c.tokens, c.tokens_full, c.dim);
  D->fup = D->tl && c.dtype == QV_BF16 && tokens_fused_ok(c.tokens, D->No, c.dim);
  D->fcmp = c.dtype == QV_BF16 && cmp_fused_ok(c.dim, c.compress_dim);
  // (8 x 8 maps: the shared-memory tile flavour of ffn_mid.cu is correct but measured SLOWER than the three-kernel chain -- 266 + 612 us
  // against 177 + 362 us per block at B = 4736, QAViTv2 step 79.0 vs 76.5 ms -- so it stays opt-in: QV_FMID_TILE=1)
  D->fmid = c.dtype == QV_BF16 && !c.ffn_v1 && ffn_mid_ok(side, c.ffn_hidden) && getenv("QV_NO_FMID") == nullptr &&
            (side == 4 || getenv("QV_FMID_TILE") != nullptr);
  D->tdt = (c.dtype == QV_BF16 && D->tl && !D->ftok && !tokens_mma_ok(c.tokens, c.tokens_full, c.dim) &&
            !tokens_mma64_ok(c.tokens, c.tokens_full, c.dim)) ? QV_F32 : c.dtype;
  D->tts = D->tdt == QV_BF16 ? 2 : 4;
  D->xstk = c.dtype == QV_BF16 && tc_shape_ok_nt(D->R, 5 * c.dim, c.dim, c.dim) && tc_shape_ok_nt(D->R, c.dim, 5 * c.dim, 5 * c.dim) &&
            tc_shape_ok_tn(D->R, 5 * c.dim, c.dim, 5 * c.dim, c.dim) && getenv("QV_NO_XSTK") == nullptr;
  D->ldp = D->xstk ? 5 * c.dim : 3 * c.dim;
  D->ldq1 = D->xstk ? 5 * c.dim : c.dim;
  D->NM = msda_tokens(c, side);
  D->Lm = D->NM < c.msda_seq_len ? D->NM : c.msda_seq_len;
  return 0;
}

void weight_shape(const Dims& D, int wi, int* N, int* K) {
  const int d = D.d;
  switch (wi) {
    case W_SWA_QKV: *N = 3 * d; *K = d; break;
    case W_MSDA_KV: *N = 2 * d; *K = d; break;
    case W_CGA_PROJ: *N = d; *K = d / 2; break;
    case W_C0: case W_C1: case W_C2: case W_C3: *N = D.cd; *K = d; break;
    case W_B1: *N = D.bh; *K = d; break;
    case W_B2: *N = d; *K = D.bh; break;
    case W_F1: *N = D.fh; *K = d; break;
    case W_F2: *N = d; *K = D.fh; break;
    case W_TL: *N = D.Nt; *K = d; break;
    case W_WRITE: *N = d + D.kb; *K = d; break;
    case W_XSTK: *N = 5 * d; *K = d; break;
    default: *N = d; *K = d; break;
  }
}

void layout_saved(const Dims& D, Saved* S) {
  Bump b;
  const size_t ts = D.ts, R = D.R, Rf = D.Rf, Ro = D.Ro, d = D.d;
  S->tl_stats = b.take(D.tl && !D.ftok ? Rf * 8 : 0);
  S->tl_ln = b.take(D.tl && !D.ftok ? Rf * d * D.tts : 0);
  S->tl_S = b.take(D.tl ? Rf * D.Nt * 4 : 0);
  S->tl_Z = b.take(D.ftok ? Rf * D.Nt * 4 : 0);
  S->xc = b.take(D.tl ? R * d * 4 : 0);
  S->n1_stats = b.take(R * 8);
  S->xn = b.take(R * d * ts);
  S->alpha = b.take(16);
  for (int i = 0; i < 4; ++i) S->snap[i] = b.take(2 * D.kb * d * 4);
  S->qkv_swa = b.take(R * (D.xstk ? 5 : 3) * d * ts);        // stacked: [qkv_swa (3d) | q_msda (d) | q_cross (d)] per row
  S->attn_swa = b.take(R * d * ts);
  S->xp = b.take((size_t)D.B * D.NM * d * ts);
  S->kv_msda = b.take((size_t)D.B * D.NM * 2 * d * ts);
  S->q_msda = D.xstk ? S->qkv_swa + 3 * d * ts : b.take(R * d * ts);
  S->attn_msda = b.take(R * d * ts);
  S->attn_cga = b.take(R * (d / 2) * ts);
  S->kbp = b.take(D.kb * D.cpg * 4);
  S->vbp = b.take(D.kb * D.cpg * 4);
  S->q_cross = D.xstk ? S->qkv_swa + 4 * d * ts : b.take(R * d * ts);
  S->Kc = b.take(D.kb * d * 4);
  S->Vc = b.take(D.kb * d * 4);
  S->attn_cross = b.take(R * d * ts);
  for (int i = 0; i < 4; ++i) S->branch[i] = b.take(R * d * ts);
  for (int i = 0; i < 4; ++i) S->nb_stats[i] = b.take(R * 8);
  for (int i = 0; i < 4; ++i) S->nb[i] = b.take(D.fcmp ? 0 : R * d * ts);   // the fused kernels never materialise LN(branch)
  S->fused = b.take(R * d * ts);
  S->h1_pre = b.take(R * D.bh * ts);
  S->h1 = b.take(R * D.bh * ts);
  S->x1 = b.take(R * d * 4);
  S->n2_stats = b.take(R * 8);
  S->y = b.take(R * d * ts);
  S->h_pre = b.take(R * D.fh * ts);
  S->h = b.take(D.v1 ? R * D.fh * ts : 0);
  S->dn_stats = b.take(R * 8);
  S->hn = b.take(D.fmid ? 0 : R * D.fh * ts);       // the fused mid-section recomputes both in backward
  S->cs = b.take(D.fmid ? 0 : R * D.fh * ts);
  S->pd_stats = b.take(R * 8);
  S->hn2 = b.take(R * D.fh * ts);
  S->o = b.take(R * d * ts);
  S->out_blk = b.take(D.tl ? R * d * 4 : 0);
  S->up = b.take(D.tl && !D.fup ? Ro * d * 4 : 0);
  S->up_stats = b.take(D.tl ? Ro * 8 : 0);
  for (int i = 0; i < W_COUNT; ++i) {
    int N, K;
    weight_shape(D, i, &N, &K);
    const bool stacked_away = D.xstk && (i == W_SWA_QKV || i == W_MSDA_Q || i == W_CROSS_Q);
    const bool need = D.dt == QV_BF16 && (i != W_TL || (D.tl && !D.ftok)) && !(D.fcmp && i >= W_C0 && i <= W_C3) && !stacked_away &&
                      (i != W_XSTK || D.xstk);
    S->wb[i] = b.take(need ? (size_t)N * K * 2 : 0);
    S->wbt[i] = b.take(need && i != W_WRITE ? (size_t)N * K * 2 : 0);
  }
  S->wstack = b.take((size_t)(d + D.kb) * d * 4);
  S->bstack = b.take((size_t)(d + D.kb) * 4);
  S->bxstk = b.take(D.xstk ? (size_t)5 * d * 4 : 0);
  S->rng = b.take(16);                          // Philox {seed, offset} snapshot of this forward call
  S->rs = b.take((size_t)2 * D.B * 4);          // DropPath keep scales of the two residual branches
  S->total = b.off;
}

void layout_scratch(const Dims& D, Scratch* S) {
  const size_t ts = D.ts, R = D.R, Rf = D.Rf, Ro = D.Ro, d = D.d;
  {
    Bump b;
    S->tl_logits = b.take(D.tl && !D.ftok ? Rf * D.Nt * D.tts : 0);
    S->tn = b.take(R * d * ts);
    S->cgbuf = b.take(R * (d + D.kb) * ts);
    S->partial = b.take((size_t)592 * 2 * D.kb * d * 4);   // bank_write_reduce uses <= 592 CTAs
    S->attn_ws_f = b.take(D.dt == QV_BF16 && D.Nt > 16 ? attn_msda64_scratch_bytes(D.B, D.d) : 0);
    S->total_fwd = b.off;
  }
  {
    Bump b;
    S->d_up = b.take(D.tl && !(D.ftok && D.fup) ? (Rf > Ro ? Rf : Ro) * d * 4 : 0);
    S->d_blk = b.take(D.tl ? R * d * 4 : 0);
    S->d_o = b.take(R * d * ts);
    S->d_hn2 = b.take(R * D.fh * ts);
    S->d_cs = b.take(D.fmid ? 0 : R * D.fh * ts);
    S->d_hn = b.take(D.fmid ? 0 : R * D.fh * ts);
    S->d_hpre = b.take(R * D.fh * ts);
    S->d_y = b.take(R * d * ts);
    S->d_x1 = b.take(R * d * 4);
    S->d_x1t = b.take(R * d * ts);
    S->d_h1 = b.take(R * D.bh * ts);
    S->d_h1pre = b.take(R * D.bh * ts);
    S->d_fused = b.take(R * d * ts);
    S->d_nb = b.take(R * d * ts);
    S->d_branch = b.take(R * d * ts);
    for (int i = 0; i < 3; ++i) S->d_branch3[i] = b.take(D.fcmp ? R * d * ts : 0);   // fused backward emits all four d_branch at once
    S->d_attn = b.take(R * d * ts);
    S->d_qkv = b.take(R * (D.xstk ? 5 : 3) * d * ts);        // stacked: [d_qkv_swa | d_q_msda | d_q_cross] per row
    S->d_kv = b.take((size_t)D.B * D.NM * 2 * d * ts);
    S->d_xp = b.take((size_t)D.B * D.NM * d * ts);
    S->d_q = D.xstk ? S->d_qkv + 3 * d * ts : b.take(R * d * ts);
    S->d_qc = D.xstk ? S->d_qkv + 4 * d * ts : S->d_q;
    // ---- zero-initialised accumulators, contiguous: d_xn | dKc | dVc | dkbp | dvbp | draw
    S->d_xn = b.take(R * d * 4);
    S->dKc = b.take(D.kb * d * 4);
    S->dVc = b.take(D.kb * d * 4);
    S->dkbp = b.take(D.kb * D.cpg * 4);
    S->dvbp = b.take(D.kb * D.cpg * 4);
    S->draw = b.take(16);
    const size_t zero_end = b.off;
    (void)zero_end;
    S->d_logits = b.take(D.tl && !D.ftok ? Rf * D.Nt * D.tts : 0);
    S->d_tl_ln = b.take(D.tl && !D.ftok ? Rf * d * D.tts : 0);
    S->attn_ws_b = b.take(D.dt == QV_BF16 && D.Nt > 16 ? attn_msda64_scratch_bytes(D.B, D.d) : 0);
    S->total_bwd = b.off;
  }
}

struct Ctx {
  Dims D;
  Saved S;
  Scratch X;
  const void* const* P;
  uint8_t* saved;
  uint8_t* scratch;
  cudaStream_t st;
  const float* pf(int i) const { return static_cast<const float*>(P[i]); }
  void* sv(size_t off) const { return saved + off; }
  float* svf(size_t off) const { return reinterpret_cast<float*>(saved + off); }
  void* sc(size_t off) const { return scratch + off; }
  float* scf(size_t off) const { return reinterpret_cast<float*>(scratch + off); }
  Weight W(int wi, const float* w) const {
    Weight r;
    weight_shape(D, wi, &r.N, &r.K);
    r.w = w;
    if (D.dt == QV_BF16) {
      r.wb = reinterpret_cast<const bf16*>(saved + S.wb[wi]);
      r.wbt = wi == W_WRITE ? nullptr : reinterpret_cast<const bf16*>(saved + S.wbt[wi]);
    }
    return r;
  }
};

const float* weight_src(const Ctx& c, int wi) {
  const int d = c.D.d;
  switch (wi) {
    case W_SWA_QKV: return c.pf(QP_SWA_QKV_W);
    case W_SWA_PROJ: return c.pf(QP_SWA_PROJ_W);
    case W_MSDA_Q: return c.pf(QP_MSDA_QKV_W);
    case W_MSDA_KV: return c.pf(QP_MSDA_QKV_W) + (size_t)d * d;
    case W_MSDA_PROJ: return c.pf(QP_MSDA_PROJ_W);
    case W_CGA_PROJ: return c.pf(QP_CGA_PROJ_W);
    case W_CROSS_Q: return c.pf(QP_CROSS_Q_W);
    case W_CROSS_PROJ: return c.pf(QP_CROSS_PROJ_W);
    case W_C0: return c.pf(QP_CSWA_W);
    case W_C1: return c.pf(QP_CMSDA_W);
    case W_C2: return c.pf(QP_CCGA_W);
    case W_C3: return c.pf(QP_CCROSS_W);
    case W_B1: return c.pf(QP_BMLP_FC1_W);
    case W_B2: return c.pf(QP_BMLP_FC2_W);
    case W_F1: return c.pf(QP_FFN_FC1_W);
    case W_F2: return c.pf(QP_FFN_FC2_W);
    case W_TL: return c.pf(QP_TL_FC_W);
    case W_WRITE: return c.svf(c.S.wstack);
  }
  return nullptr;
}

int init_ctx(Ctx* c, const qavit_block_cfg* cfg, const void* const* params, const void* saved, void* scratch, void* stream) {
  QV_CHECK(cfg && params && saved && scratch, "null argument");
  QV_TRY(make_dims(*cfg, &c->D));
  layout_saved(c->D, &c->S);
  layout_scratch(c->D, &c->X);
  c->P = params;
  c->saved = static_cast<uint8_t*>(const_cast<void*>(saved));
  c->scratch = static_cast<uint8_t*>(scratch);
  c->st = static_cast<cudaStream_t>(stream);
  return 0;
}

// GlobalTokenBank.write(branch.norm(out)) -- H:468 / 531 / 594 -> H:296-321
int bank_write(const Ctx& c, const qavit_block_cfg& cfg, const void* branch_out, int nw, int nb, long long* update_count) {
  const Dims& D = c.D;
  void* tn = c.sc(c.X.tn);
  void* cg = c.sc(c.X.cgbuf);
  QV_TRY(ln_fwd(c.st, D.dt, branch_out, D.d, D.R, D.d, c.pf(nw), c.pf(nb), 1e-5f, 0, c.pf(QP_BANK_WN_W), c.pf(QP_BANK_WN_B),
                D.dt, tn, D.d, nullptr));
  GemmEpi e;
  e.bias = c.svf(c.S.bstack); e.C = cg; e.ldc = D.d + D.kb;
  e.c_f32 = D.dt == QV_F32;
  QV_TRY(gemm_nt(c.st, D.dt, tn, D.d, D.R, c.W(W_WRITE, weight_src(c, W_WRITE)), e));
  int np = 0;
  QV_TRY(bank_write_reduce(c.st, D.dt, tn, cg, D.d + D.kb, D.B, D.Nt, D.d, D.kb, c.scf(c.X.partial), &np));
  QV_TRY(bank_write_apply(c.st, c.scf(c.X.partial), np, D.B, D.d, D.kb, const_cast<float*>(c.pf(QP_BANK_K)),
                          const_cast<float*>(c.pf(QP_BANK_V)), update_count, cfg.bank_v1));
  return 0;
}

int snapshot_bank(const Ctx& c, int i) {
  const int n = c.D.kb * c.D.d;
  return copy2_f32(c.st, c.svf(c.S.snap[i]), c.pf(QP_BANK_K), n, c.svf(c.S.snap[i]) + n, c.pf(QP_BANK_V), n);
}
const float* snap_k(const Ctx& c, int i) { return c.svf(c.S.snap[i]); }
const float* snap_v(const Ctx& c, int i) { return c.svf(c.S.snap[i]) + (size_t)c.D.kb * c.D.d; }

GemmEpi epi_t(const Ctx& c, const float* bias, void* C, int ldc) {
  GemmEpi e;
  e.bias = bias; e.C = C; e.ldc = ldc; e.c_f32 = c.D.dt == QV_F32;
  return e;
}

AttnP attn_params(const Ctx& c, int mode) {
  const Dims& D = c.D;
  AttnP p{};
  p.mode = mode; p.B = D.B; p.Nt = D.Nt; p.side = D.side; p.ws = D.ws; p.H = D.H; p.hd = D.hd; p.kb = D.kb;
  p.klin = D.klin; p.NM = D.NM;
  p.L = mode == 0 ? D.ws * D.ws : (mode == 1 ? D.Lm : 0);
  return p;
}

// dropout sites of one block call (each call has its own rng snapshot, so ids only need to differ within a block)
enum { DS_ATT = 1 /* +branch */, DS_PROJ = 5 /* +branch */, DS_B1 = 9, DS_B2 = 10, DS_FFN = 11, DS_PATH = 12 };

struct DropCfg {
  bool drop = false, path = false;
  float p = 0.f, p_path = 0.f;
  const unsigned long long* snap = nullptr;
  const float *rs1 = nullptr, *rs2 = nullptr;
  DropP site(uint32_t id) const {   // an nn.Dropout(config.dropout) / SDPA dropout_p site; inactive when !drop
    DropP d;
    if (drop) { d.p = p; d.rng = snap; d.site = id; }
    return d;
  }
};
int make_dropcfg(const Ctx& c, const qavit_block_cfg& cfg, DropCfg* dc) {
  const bool train = cfg.train != 0;
  QV_CHECK(cfg.dropout >= 0.f && cfg.dropout < 1.f && cfg.drop_path >= 0.f && cfg.drop_path < 1.f, "dropout=%f / drop_path=%f out of [0, 1)",
           cfg.dropout, cfg.drop_path);
  dc->drop = train && cfg.dropout > 0.f;
  dc->path = train && cfg.drop_path > 0.f;
  dc->p = cfg.dropout; dc->p_path = cfg.drop_path;
  dc->snap = static_cast<const unsigned long long*>(c.sv(c.S.rng));
  if (dc->path) { dc->rs1 = c.svf(c.S.rs); dc->rs2 = c.svf(c.S.rs) + c.D.B; }
  return 0;
}
// the tcgen05 GEMM epilogue applies a dropout site itself; the fp32 SIMT flavour does not (a drop_rows pass follows it)
bool fused_nt(const Ctx& c, const Weight& W, int M, int lda) { return c.D.dt == QV_BF16 && W.wb && tc_shape_ok_nt(M, W.N, W.K, lda); }
bool fused_nn(const Ctx& c, const Weight& W, int M, int ldy) { return c.D.dt == QV_BF16 && W.wbt && tc_shape_ok_nt(M, W.K, W.N, ldy); }
int drop_inplace(const Ctx& c, const DropCfg& dc, void* x, int C, uint32_t id, const float* rowscale) {
  return drop_rows(c.st, c.D.dt, x, C, c.D.R, C, dc.site(id), rowscale, c.D.Nt, nullptr, 0, nullptr, nullptr, 0);
}

}  // namespace

extern "C" int qavit_block_workspace(const qavit_block_cfg* cfg, size_t* saved_bytes, size_t* scratch_bytes) {
  QV_CHECK(cfg, "null cfg");
  Dims D;
  QV_TRY(make_dims(*cfg, &D));
  Saved S;
  Scratch X;
  layout_saved(D, &S);
  layout_scratch(D, &X);
  if (saved_bytes) *saved_bytes = S.total + 256;
  if (scratch_bytes) *scratch_bytes = (X.total_fwd > X.total_bwd ? X.total_fwd : X.total_bwd) + 256;
  return 0;
}

// =====================================================================================================
// forward
// =====================================================================================================
extern "C" int qavit_block_forward(const qavit_block_cfg* cfg, const void* const* params, long long* update_count,
                                   unsigned long long* rng, const float* x_in, float* out, void* saved, void* scratch,
                                   void* stream) {
  QV_RANGE("qavit_block_forward");
  Ctx c;
  QV_TRY(init_ctx(&c, cfg, params, saved, scratch, stream));
  const Dims& D = c.D;
  const Saved& S = c.S;
  const int dt = D.dt, d = D.d, R = D.R;
  cudaStream_t st = c.st;
  const bool train = cfg->train != 0;
  QV_CHECK(!train || cfg->bank_v1 || update_count, "train mode needs update_count");
  DropCfg dc;
  QV_TRY(make_dropcfg(c, *cfg, &dc));
  if (dc.drop || dc.path) {
    QV_CHECK(rng, "train-mode dropout / DropPath needs the rng state {seed, offset}");
    QV_TRY(rng_snapshot_advance(st, rng, static_cast<unsigned long long*>(c.sv(S.rng))));
    if (dc.path) {
      DropP d; d.p = dc.p_path; d.rng = dc.snap; d.site = DS_PATH;
      QV_TRY(droppath_scales(st, d, D.B, c.svf(S.rs), c.svf(S.rs) + D.B));
    }
  }

  // ---- stacked write weight [write_compression ; write_gate] and the bf16 weight copies (one batched launch)
  if (train) {
    if (dt == QV_F32) {
      QV_CUDA(cudaMemcpyAsync(c.sv(S.wstack), c.pf(QP_BANK_WC_W), (size_t)d * d * 4, cudaMemcpyDeviceToDevice, st));
      QV_CUDA(cudaMemcpyAsync(c.svf(S.wstack) + (size_t)d * d, c.pf(QP_BANK_WG_W), (size_t)D.kb * d * 4, cudaMemcpyDeviceToDevice, st));
    }
    QV_TRY(copy2_f32(st, c.svf(S.bstack), c.pf(QP_BANK_WC_B), d, c.svf(S.bstack) + d, c.pf(QP_BANK_WG_B), D.kb));
  }
  if (dt == QV_BF16) {
    ConvertJobs jobs{};
    for (int wi = 0; wi < W_COUNT; ++wi) {
      if ((wi == W_TL && (!D.tl || D.ftok)) || (wi == W_WRITE && !train) || (D.fcmp && wi >= W_C0 && wi <= W_C3)) continue;
      if (D.xstk && (wi == W_SWA_QKV || wi == W_MSDA_Q || wi == W_CROSS_Q)) continue;
      if (wi == W_XSTK && !D.xstk) continue;
      int N, K;
      weight_shape(D, wi, &N, &K);
      bf16* wb = reinterpret_cast<bf16*>(c.sv(S.wb[wi]));
      if (wi == W_XSTK) {    // three sources stacked along N (rows), their transposes side by side in the [d, 5d] copy, biases concatenated
        bf16* wbt = reinterpret_cast<bf16*>(c.sv(S.wbt[wi]));
        float* bs = c.svf(S.bxstk);
        jobs.j[jobs.n++] = ConvertJob{c.pf(QP_SWA_QKV_W), 3 * d, d, wb, wbt, 5 * d, c.pf(QP_SWA_QKV_B), bs, 3 * d};
        jobs.j[jobs.n++] = ConvertJob{c.pf(QP_MSDA_QKV_W), d, d, wb + (size_t)3 * d * d, wbt + 3 * d, 5 * d, c.pf(QP_MSDA_QKV_B), bs + 3 * d, d};
        jobs.j[jobs.n++] = ConvertJob{c.pf(QP_CROSS_Q_W), d, d, wb + (size_t)4 * d * d, wbt + 4 * d, 5 * d, c.pf(QP_CROSS_Q_B), bs + 4 * d, d};
      } else if (wi == W_WRITE) {   // two sources stacked along N
        jobs.j[jobs.n++] = ConvertJob{c.pf(QP_BANK_WC_W), d, K, wb, nullptr};
        jobs.j[jobs.n++] = ConvertJob{c.pf(QP_BANK_WG_W), D.kb, K, wb + (size_t)d * K, nullptr};
      } else {
        jobs.j[jobs.n++] = ConvertJob{weight_src(c, wi), N, K, wb, reinterpret_cast<bf16*>(c.sv(S.wbt[wi]))};
      }
    }
    QV_TRY(convert_weights_batched(st, jobs));
  }

  // ---- TokenLearner (H:985-1002)
  const float* x = x_in;
  if (D.ftok) {   // LayerNorm + gate + softmax over tokens + pooling in one split-precision kernel; the logits never leave the SM
    QV_TRY(tlf_fwd(st, x_in, D.B, D.Nf, c.pf(QP_TL_LN_W), c.pf(QP_TL_LN_B), c.pf(QP_TL_FC_W), c.pf(QP_TL_FC_B), 1e-5f, c.svf(S.tl_S),
                   c.svf(S.tl_Z), c.svf(S.xc)));
    x = c.svf(S.xc);
  } else if (D.tl) {
    const int tdt = D.tdt;
    QV_TRY(ln_fwd(st, QV_F32, x_in, d, D.Rf, d, c.pf(QP_TL_LN_W), c.pf(QP_TL_LN_B), 1e-5f, 0, nullptr, nullptr, tdt,
                  c.sv(S.tl_ln), d, c.svf(S.tl_stats)));
    GemmEpi et = epi_t(c, c.pf(QP_TL_FC_B), c.sc(c.X.tl_logits), D.Nt);
    et.c_f32 = tdt == QV_F32;
    QV_TRY(gemm_nt(st, tdt, c.sv(S.tl_ln), d, D.Rf, c.W(W_TL, c.pf(QP_TL_FC_W)), et));
    QV_TRY(token_learner_fwd(st, tdt, x_in, c.sc(c.X.tl_logits), D.B, D.Nf, D.Nt, d, c.svf(S.tl_S), c.svf(S.xc)));
    x = c.svf(S.xc);
  }

  // branch output projection + nn.Dropout (H:464-465): the dropout site rides in the tcgen05 epilogue when there is one
  auto proj_gemm = [&](const void* A, int lda, const Weight& W, const float* bias, int i) -> int {
    GemmEpi e = epi_t(c, bias, c.sv(S.branch[i]), d);
    const bool fused = dc.drop && fused_nt(c, W, R, lda);
    if (fused) e.drop = dc.site(DS_PROJ + i);
    QV_TRY(gemm_nt(st, dt, A, lda, R, W, e));
    if (dc.drop && !fused) QV_TRY(drop_inplace(c, dc, c.sv(S.branch[i]), d, DS_PROJ + i, nullptr));
    return 0;
  };

  // ---- norm1, fusion weights
  void* xn = c.sv(S.xn);
  QV_TRY(ln_fwd(st, QV_F32, x, d, R, d, c.pf(QP_NORM1_W), c.pf(QP_NORM1_B), 1e-5f, 0, nullptr, nullptr, dt, xn, d, c.svf(S.n1_stats)));
  QV_TRY(fusion_softmax(st, c.pf(QP_FUSION_W), 4, c.svf(S.alpha)));

  // ---- SWA (H:441-469)
  nvtxRangePushA("swa");
  QV_TRY(snapshot_bank(c, 0));
  if (D.xstk)   // swa.qkv | msda q | cross q in one launch (the other two consumers read their column slices later)
    QV_TRY(gemm_nt(st, dt, xn, d, R, c.W(W_XSTK, nullptr), epi_t(c, c.svf(S.bxstk), c.sv(S.qkv_swa), 5 * d)));
  else
    QV_TRY(gemm_nt(st, dt, xn, d, R, c.W(W_SWA_QKV, c.pf(QP_SWA_QKV_W)), epi_t(c, c.pf(QP_SWA_QKV_B), c.sv(S.qkv_swa), 3 * d)));
  {
    AttnP p = attn_params(c, 0);
    p.q = c.sv(S.qkv_swa); p.ldq = D.ldp; p.qcol = 0;
    p.kv = c.sv(S.qkv_swa); p.ldkv = D.ldp; p.kcol = d; p.vcol = 2 * d;
    p.Ek = c.pf(QP_SWA_EK); p.Ev = c.pf(QP_SWA_EV);
    p.bank_k = snap_k(c, 0); p.bank_v = snap_v(c, 0);
    p.out = c.sv(S.attn_swa); p.ldo = d;
    p.drop = dc.site(DS_ATT + 0);
    QV_TRY(attn_fwd(st, dt, p));
  }
  QV_TRY(proj_gemm(c.sv(S.attn_swa), d, c.W(W_SWA_PROJ, c.pf(QP_SWA_PROJ_W)), c.pf(QP_SWA_PROJ_B), 0));   // H:465, before the bank write
  if (train) QV_TRY(bank_write(c, *cfg, c.sv(S.branch[0]), QP_SWA_NORM_W, QP_SWA_NORM_B, update_count));

  // ---- MSDA (H:496-532)
  nvtxRangePop(); nvtxRangePushA("msda");
  QV_TRY(snapshot_bank(c, 1));
  QV_TRY(msda_pool_fwd(st, dt, xn, D.B, D.Nt, D.side, d, cfg->dilations, cfg->n_dilations, cfg->pool_stride, D.NM, c.sv(S.xp)));
  QV_TRY(gemm_nt(st, dt, c.sv(S.xp), d, D.B * D.NM, c.W(W_MSDA_KV, weight_src(c, W_MSDA_KV)),
                 epi_t(c, c.pf(QP_MSDA_QKV_B) + d, c.sv(S.kv_msda), 2 * d)));
  if (!D.xstk) QV_TRY(gemm_nt(st, dt, xn, d, R, c.W(W_MSDA_Q, c.pf(QP_MSDA_QKV_W)), epi_t(c, c.pf(QP_MSDA_QKV_B), c.sv(S.q_msda), d)));
  {
    AttnP p = attn_params(c, 1);
    p.q = c.sv(S.q_msda); p.ldq = D.ldq1; p.qcol = 0;
    p.kv = c.sv(S.kv_msda); p.ldkv = 2 * d; p.kcol = 0; p.vcol = d;
    p.Ek = c.pf(QP_MSDA_EK); p.Ev = c.pf(QP_MSDA_EV);
    p.bank_k = snap_k(c, 1); p.bank_v = snap_v(c, 1);
    p.out = c.sv(S.attn_msda); p.ldo = d;
    if (dt == QV_BF16 && D.Nt > 16) p.wsp = c.sc(c.X.attn_ws_f);
    p.drop = dc.site(DS_ATT + 1);
    QV_TRY(attn_fwd(st, dt, p));
  }
  QV_TRY(proj_gemm(c.sv(S.attn_msda), d, c.W(W_MSDA_PROJ, c.pf(QP_MSDA_PROJ_W)), c.pf(QP_MSDA_PROJ_B), 1));   // H:529
  if (train) QV_TRY(bank_write(c, *cfg, c.sv(S.branch[1]), QP_MSDA_NORM_W, QP_MSDA_NORM_B, update_count));

  // ---- CGA (H:559-595)
  nvtxRangePop(); nvtxRangePushA("cga");
  QV_TRY(snapshot_bank(c, 2));
  QV_TRY(small_linear_fwd2(st, D.kb, d, D.cpg, snap_k(c, 2), c.pf(QP_CGA_BK_W), c.pf(QP_CGA_BK_B), c.svf(S.kbp),
                           snap_v(c, 2), c.pf(QP_CGA_BV_W), c.pf(QP_CGA_BV_B), c.svf(S.vbp)));
  {
    CgaP p{};
    p.B = D.B; p.Nt = D.Nt; p.G = D.G; p.H = D.H; p.kb = D.kb; p.cg = D.cg; p.cpg = D.cpg;
    p.xn = xn; p.ldx = d;
    p.Wq = c.pf(QP_CGA_Q_W); p.bq = c.pf(QP_CGA_Q_B); p.Wk = c.pf(QP_CGA_K_W); p.bk = c.pf(QP_CGA_K_B);
    p.Wv = c.pf(QP_CGA_V_W); p.bv = c.pf(QP_CGA_V_B);
    p.kbp = c.svf(S.kbp); p.vbp = c.svf(S.vbp);
    p.out = c.sv(S.attn_cga); p.ldo = d / 2;
    p.drop = dc.site(DS_ATT + 2);
    QV_TRY(cga_fwd(st, dt, p));
  }
  QV_TRY(proj_gemm(c.sv(S.attn_cga), d / 2, c.W(W_CGA_PROJ, c.pf(QP_CGA_PROJ_W)), c.pf(QP_CGA_PROJ_B), 2));   // H:592
  if (train) QV_TRY(bank_write(c, *cfg, c.sv(S.branch[2]), QP_CGA_NORM_W, QP_CGA_NORM_B, update_count));

  // ---- Cross (H:613-626)
  nvtxRangePop(); nvtxRangePushA("cross");
  QV_TRY(snapshot_bank(c, 3));
  QV_TRY(small_linear_fwd2(st, D.kb, d, d, snap_k(c, 3), c.pf(QP_CROSS_K_W), c.pf(QP_CROSS_K_B), c.svf(S.Kc),
                           snap_v(c, 3), c.pf(QP_CROSS_V_W), c.pf(QP_CROSS_V_B), c.svf(S.Vc)));
  if (!D.xstk) QV_TRY(gemm_nt(st, dt, xn, d, R, c.W(W_CROSS_Q, c.pf(QP_CROSS_Q_W)), epi_t(c, c.pf(QP_CROSS_Q_B), c.sv(S.q_cross), d)));
  {
    AttnP p = attn_params(c, 2);
    p.q = c.sv(S.q_cross); p.ldq = D.ldq1; p.qcol = 0;
    p.kc = c.svf(S.Kc); p.vc = c.svf(S.Vc);
    p.out = c.sv(S.attn_cross); p.ldo = d;
    p.drop = dc.site(DS_ATT + 3);
    QV_TRY(attn_fwd(st, dt, p));
  }
  QV_TRY(proj_gemm(c.sv(S.attn_cross), d, c.W(W_CROSS_PROJ, c.pf(QP_CROSS_PROJ_W)), c.pf(QP_CROSS_PROJ_B), 3));   // H:625

  // ---- per-branch LN -> compress -> fusion scale + concat (H:1074-1079)
  nvtxRangePop(); nvtxRangePushA("fusion+mlp");
  static const int kNormW[4] = {QP_NSWA_W, QP_NMSDA_W, QP_NCGA_W, QP_NCROSS_W};
  static const int kCompW[4] = {QP_CSWA_W, QP_CMSDA_W, QP_CCGA_W, QP_CCROSS_W};
  if (D.fcmp) {
    const void* xs[4]; const float *gs[4], *bs[4], *ws[4], *cbs[4]; float* sts[4];
    for (int i = 0; i < 4; ++i) {
      xs[i] = c.sv(S.branch[i]); gs[i] = c.pf(kNormW[i]); bs[i] = c.pf(kNormW[i] + 1); ws[i] = c.pf(kCompW[i]); cbs[i] = c.pf(kCompW[i] + 1);
      sts[i] = c.svf(S.nb_stats[i]);
    }
    QV_TRY(cmpf_fwd(st, R, xs, gs, bs, ws, cbs, c.svf(S.alpha), 1e-5f, c.sv(S.fused), sts));
  }
  for (int i = 0; i < 4 && !D.fcmp; ++i) {
    QV_TRY(ln_fwd(st, dt, c.sv(S.branch[i]), d, R, d, c.pf(kNormW[i]), c.pf(kNormW[i] + 1), 1e-5f, 0, nullptr, nullptr, dt,
                  c.sv(S.nb[i]), d, c.svf(S.nb_stats[i])));
    GemmEpi e = epi_t(c, c.pf(kCompW[i] + 1), static_cast<uint8_t*>(c.sv(S.fused)) + (size_t)i * D.cd * D.ts, d);
    e.scale_pre = c.svf(S.alpha) + i;
    QV_TRY(gemm_nt(st, dt, c.sv(S.nb[i]), d, R, c.W(W_C0 + i, c.pf(kCompW[i])), e));
  }
  // ---- BottleneckMLP + residual (H:651-656, 1082)
  {
    GemmEpi e = epi_t(c, c.pf(QP_BMLP_FC1_B), c.sv(S.h1_pre), D.bh);
    e.gelu = 1; e.C2 = c.sv(S.h1); e.ldc2 = D.bh; e.c2_f32 = dt == QV_F32;
    const Weight wb1 = c.W(W_B1, c.pf(QP_BMLP_FC1_W)), wb2 = c.W(W_B2, c.pf(QP_BMLP_FC2_W));
    const bool f1 = dc.drop && fused_nt(c, wb1, R, d), f2 = (dc.drop || dc.path) && fused_nt(c, wb2, R, D.bh);
    if (f1) e.drop = dc.site(DS_B1);                        // dropout after the activation: C2 only
    QV_TRY(gemm_nt(st, dt, c.sv(S.fused), d, R, wb1, e));
    if (dc.drop && !f1) QV_TRY(drop_inplace(c, dc, c.sv(S.h1), D.bh, DS_B1, nullptr));
    if ((!dc.drop && !dc.path) || f2) {   // x1 = x + drop_path1(dropout(fc2(h1)))   (H:654-656, 1082) in the GEMM epilogue
      GemmEpi e2;
      e2.bias = c.pf(QP_BMLP_FC2_B); e2.resid = x; e2.ldr = d; e2.C2 = c.sv(S.x1); e2.ldc2 = d; e2.c2_f32 = 1;
      if (f2) { e2.drop = dc.site(DS_B2); e2.rowscale = dc.rs1; e2.rows_per_img = D.Nt; }
      QV_TRY(gemm_nt(st, dt, c.sv(S.h1), D.bh, R, wb2, e2));
    } else {
      void* tmp = c.sc(c.X.tn);   // free since the last bank write
      QV_TRY(gemm_nt(st, dt, c.sv(S.h1), D.bh, R, c.W(W_B2, c.pf(QP_BMLP_FC2_W)), epi_t(c, c.pf(QP_BMLP_FC2_B), tmp, d)));
      QV_TRY(drop_rows(st, dt, tmp, d, R, d, dc.site(DS_B2), dc.rs1, D.Nt, x, d, nullptr, c.svf(S.x1), d));
    }
  }
  // ---- norm2 + CCF-FFN + residual (H:700-712, 1083)
  float* blk_out = D.tl ? c.svf(S.out_blk) : out;
  QV_TRY(ln_fwd(st, QV_F32, c.sv(S.x1), d, R, d, c.pf(QP_NORM2_W), c.pf(QP_NORM2_B), 1e-5f, 0, nullptr, nullptr, dt, c.sv(S.y), d, c.svf(S.n2_stats)));
  if (!D.v1) {
    QV_TRY(gemm_nt(st, dt, c.sv(S.y), d, R, c.W(W_F1, c.pf(QP_FFN_FC1_W)), epi_t(c, c.pf(QP_FFN_FC1_B), c.sv(S.h_pre), D.fh)));
    if (D.fmid) {
      QV_TRY(ffn_mid_fwd(st, c.sv(S.h_pre), D.B, D.side, D.fh, c.pf(QP_FFN_DWN_W), c.pf(QP_FFN_DWN_B), c.pf(QP_FFN_DW_W),
                         cfg->dwconv_bias ? c.pf(QP_FFN_DW_B) : nullptr, c.pf(QP_FFN_SCALE), c.pf(QP_FFN_PDN_W), c.pf(QP_FFN_PDN_B), 1e-5f,
                         c.sv(S.hn2), c.svf(S.dn_stats), c.svf(S.pd_stats)));
    } else {
      QV_TRY(ln_fwd(st, dt, c.sv(S.h_pre), D.fh, R, D.fh, c.pf(QP_FFN_DWN_W), c.pf(QP_FFN_DWN_B), 1e-5f, 1, nullptr, nullptr, dt,
                    c.sv(S.hn), D.fh, c.svf(S.dn_stats)));
      QV_TRY(dwconv_fwd(st, dt, c.sv(S.hn), D.B, D.side, D.fh, c.pf(QP_FFN_DW_W), cfg->dwconv_bias ? c.pf(QP_FFN_DW_B) : nullptr,
                        c.pf(QP_FFN_SCALE), c.sv(S.cs)));
      QV_TRY(ln_fwd(st, dt, c.sv(S.cs), D.fh, R, D.fh, c.pf(QP_FFN_PDN_W), c.pf(QP_FFN_PDN_B), 1e-5f, 0, nullptr, nullptr, dt,
                    c.sv(S.hn2), D.fh, c.svf(S.pd_stats)));
    }
    GemmEpi e = epi_t(c, c.pf(QP_FFN_FC2_B), c.sv(S.o), d);
    const Weight wf2 = c.W(W_F2, c.pf(QP_FFN_FC2_W));
    const bool ff = (dc.drop || dc.path) && fused_nt(c, wf2, R, D.fh);
    if ((!dc.drop && !dc.path) || ff) {
      e.resid = c.svf(S.x1); e.ldr = d; e.scale_res = c.pf(QP_FFN_GAMMA); e.C2 = blk_out; e.ldc2 = d; e.c2_f32 = 1;
      if (ff) { e.drop = dc.site(DS_FFN); e.rowscale = dc.rs2; e.rows_per_img = D.Nt; }   // `o` = the dropped, path-scaled value
      QV_TRY(gemm_nt(st, dt, c.sv(S.hn2), D.fh, R, wf2, e));
    } else {   // out = x1 + drop_path2(gamma * dropout(fc2))   (H:710-712, 1083); `o` keeps the dropped, path-scaled value
      QV_TRY(gemm_nt(st, dt, c.sv(S.hn2), D.fh, R, c.W(W_F2, c.pf(QP_FFN_FC2_W)), e));
      QV_TRY(drop_rows(st, dt, c.sv(S.o), d, R, d, dc.site(DS_FFN), dc.rs2, D.Nt, c.svf(S.x1), d, c.pf(QP_FFN_GAMMA), blk_out, d));
    }
  } else {
    GemmEpi e = epi_t(c, c.pf(QP_FFN_FC1_B), c.sv(S.h_pre), D.fh);
    e.gelu = 1; e.C2 = c.sv(S.h); e.ldc2 = D.fh; e.c2_f32 = dt == QV_F32;
    QV_TRY(gemm_nt(st, dt, c.sv(S.y), d, R, c.W(W_F1, c.pf(QP_FFN_FC1_W)), e));
    QV_TRY(dwconv_fwd(st, dt, c.sv(S.h), D.B, D.side, D.fh, c.pf(QP_FFN_DW_W), cfg->dwconv_bias ? c.pf(QP_FFN_DW_B) : nullptr, nullptr, c.sv(S.cs)));
    const Weight wf2 = c.W(W_F2, c.pf(QP_FFN_FC2_W));
    const bool ff = (dc.drop || dc.path) && fused_nt(c, wf2, R, D.fh);
    if ((!dc.drop && !dc.path) || ff) {
      GemmEpi e2;
      e2.bias = c.pf(QP_FFN_FC2_B); e2.resid = c.svf(S.x1); e2.ldr = d; e2.C2 = blk_out; e2.ldc2 = d; e2.c2_f32 = 1;
      if (ff) { e2.drop = dc.site(DS_FFN); e2.rowscale = dc.rs2; e2.rows_per_img = D.Nt; }
      QV_TRY(gemm_nt(st, dt, c.sv(S.cs), D.fh, R, wf2, e2));
    } else {   // QAViT.py:580-582: out = x1 + drop_path2(dropout(fc2))
      QV_TRY(gemm_nt(st, dt, c.sv(S.cs), D.fh, R, c.W(W_F2, c.pf(QP_FFN_FC2_W)), epi_t(c, c.pf(QP_FFN_FC2_B), c.sv(S.o), d)));
      QV_TRY(drop_rows(st, dt, c.sv(S.o), d, R, d, dc.site(DS_FFN), dc.rs2, D.Nt, c.svf(S.x1), d, nullptr, blk_out, d));
    }
  }
  nvtxRangePop();
  // ---- TokenUpMix (H:1016-1031)
  QV_RANGE("token_upmix");
  if (D.fup) {   // up-mix GEMM + LayerNorm in one kernel: the [B No, d] pre-norm tensor is never written (backward recomputes it)
    QV_TRY(upf_fwd(st, blk_out, D.B, D.No, c.pf(QP_UP_FC_W), c.pf(QP_UP_FC_B), c.pf(QP_UP_LN_W), c.pf(QP_UP_LN_B), 1e-5f, out,
                   c.svf(S.up_stats)));
  } else if (D.tl) {
    QV_TRY(token_upmix_fwd(st, dt, blk_out, D.B, D.Nt, D.No, d, c.pf(QP_UP_FC_W), c.pf(QP_UP_FC_B), c.svf(S.up)));
    QV_TRY(ln_fwd(st, QV_F32, c.sv(S.up), d, D.Ro, d, c.pf(QP_UP_LN_W), c.pf(QP_UP_LN_B), 1e-5f, 0, nullptr, nullptr, QV_F32, out, d, c.svf(S.up_stats)));
  }
  return 0;
}

// =====================================================================================================
// backward
// =====================================================================================================
extern "C" int qavit_block_backward(const qavit_block_cfg* cfg, const void* const* params, float* const* grads,
                                    const float* x_in, const float* dout, float* dx, const void* saved, void* scratch,
                                    void* stream) {
  QV_RANGE("qavit_block_backward");
  Ctx c;
  QV_TRY(init_ctx(&c, cfg, params, saved, scratch, stream));
  QV_CHECK(grads && dout && dx, "null argument");
  const Dims& D = c.D;
  const Saved& S = c.S;
  const Scratch& X = c.X;
  const int dt = D.dt, d = D.d, R = D.R;
  cudaStream_t st = c.st;
  auto G = [&](int i) { return grads[i]; };
  DropCfg dc;
  QV_TRY(make_dropcfg(c, *cfg, &dc));
  const float* x = D.tl ? c.svf(S.xc) : x_in;
  const float* blk_out = D.tl ? c.svf(S.out_blk) : nullptr;
  (void)blk_out;

  // zero the accumulators (one contiguous range: d_xn .. draw)
  QV_CUDA(cudaMemsetAsync(c.sc(X.d_xn), 0, X.draw + 16 - X.d_xn, st));

  // ---- TokenUpMix backward
  const float* dblk = dout;
  if (D.fup) {   // (the up-mix bias gradient is exactly zero -- a per-token constant removed by the LayerNorm that follows -- and stays 0)
    QV_TRY(upf_bwd(st, c.svf(S.out_blk), dout, c.svf(S.up_stats), D.B, D.No, c.pf(QP_UP_FC_W), c.pf(QP_UP_FC_B), c.pf(QP_UP_LN_W),
                   c.scf(X.d_blk), G(QP_UP_FC_W), G(QP_UP_LN_W), G(QP_UP_LN_B)));
    dblk = c.scf(X.d_blk);
  } else if (D.tl) {
    QV_TRY(ln_bwd(st, QV_F32, c.sv(S.up), d, QV_F32, dout, d, D.Ro, d, c.pf(QP_UP_LN_W), c.svf(S.up_stats), 0, QV_F32, nullptr,
                  c.scf(X.d_up), nullptr, G(QP_UP_LN_W), G(QP_UP_LN_B)));
    QV_TRY(token_upmix_bwd(st, dt, c.svf(S.out_blk), c.scf(X.d_up), D.B, D.Nt, D.No, d, c.pf(QP_UP_FC_W), c.scf(X.d_blk),
                           G(QP_UP_FC_W), G(QP_UP_FC_B)));
    dblk = c.scf(X.d_blk);
  }

  // ---- CCF-FFN backward
  if (!D.v1) {
    {   // d_o = gamma * dblk through the FFN dropout site / DropPath scale (one pass)
      const DropP site = dc.site(DS_FFN);
      QV_TRY(gamma_bwd(st, dt, dblk, c.sv(S.o), (long)R * d, c.pf(QP_FFN_GAMMA), c.sc(X.d_o), G(QP_FFN_GAMMA), &site, dc.rs2, D.Nt, d));
    }
    QV_TRY(gemm_tn(st, dt, c.sc(X.d_o), d, c.sv(S.hn2), D.fh, R, d, D.fh, G(QP_FFN_FC2_W), G(QP_FFN_FC2_B), nullptr));
    QV_TRY(gemm_nn(st, dt, c.sc(X.d_o), d, R, c.W(W_F2, c.pf(QP_FFN_FC2_W)), epi_t(c, nullptr, c.sc(X.d_hn2), D.fh)));
    if (D.fmid) {
      QV_TRY(ffn_mid_bwd(st, c.sv(S.h_pre), c.sc(X.d_hn2), c.svf(S.dn_stats), c.svf(S.pd_stats), D.B, D.side, D.fh, c.pf(QP_FFN_DWN_W),
                         c.pf(QP_FFN_DWN_B), c.pf(QP_FFN_DW_W), cfg->dwconv_bias ? c.pf(QP_FFN_DW_B) : nullptr, c.pf(QP_FFN_SCALE),
                         c.pf(QP_FFN_PDN_W), c.sc(X.d_hpre), G(QP_FFN_DWN_W), G(QP_FFN_DWN_B), G(QP_FFN_DW_W),
                         cfg->dwconv_bias ? G(QP_FFN_DW_B) : nullptr, G(QP_FFN_SCALE), G(QP_FFN_PDN_W), G(QP_FFN_PDN_B)));
    } else {
      QV_TRY(ln_bwd(st, dt, c.sv(S.cs), D.fh, dt, c.sc(X.d_hn2), D.fh, R, D.fh, c.pf(QP_FFN_PDN_W), c.svf(S.pd_stats), 0, dt,
                    c.sc(X.d_cs), nullptr, nullptr, G(QP_FFN_PDN_W), G(QP_FFN_PDN_B)));
      QV_TRY(dwconv_bwd(st, dt, c.sv(S.hn), c.sc(X.d_cs), D.B, D.side, D.fh, c.pf(QP_FFN_DW_W),
                        cfg->dwconv_bias ? c.pf(QP_FFN_DW_B) : nullptr, c.pf(QP_FFN_SCALE), c.sc(X.d_hn), G(QP_FFN_DW_W),
                        cfg->dwconv_bias ? G(QP_FFN_DW_B) : nullptr, G(QP_FFN_SCALE)));
      QV_TRY(ln_bwd(st, dt, c.sv(S.h_pre), D.fh, dt, c.sc(X.d_hn), D.fh, R, D.fh, c.pf(QP_FFN_DWN_W), c.svf(S.dn_stats), 1, dt,
                    c.sc(X.d_hpre), nullptr, nullptr, G(QP_FFN_DWN_W), G(QP_FFN_DWN_B)));
    }
  } else {
    if (dc.drop || dc.path) {
      const DropP site = dc.site(DS_FFN);
      QV_TRY(gamma_bwd(st, dt, dblk, nullptr, (long)R * d, nullptr, c.sc(X.d_o), nullptr, &site, dc.rs2, D.Nt, d));
    } else {
      QV_TRY(cast_f32_to_t(st, dt, dblk, (long)R * d, c.sc(X.d_o)));
    }
    QV_TRY(gemm_tn(st, dt, c.sc(X.d_o), d, c.sv(S.cs), D.fh, R, d, D.fh, G(QP_FFN_FC2_W), G(QP_FFN_FC2_B), nullptr));
    QV_TRY(gemm_nn(st, dt, c.sc(X.d_o), d, R, c.W(W_F2, c.pf(QP_FFN_FC2_W)), epi_t(c, nullptr, c.sc(X.d_cs), D.fh)));
    QV_TRY(dwconv_bwd(st, dt, c.sv(S.h), c.sc(X.d_cs), D.B, D.side, D.fh, c.pf(QP_FFN_DW_W),
                      cfg->dwconv_bias ? c.pf(QP_FFN_DW_B) : nullptr, nullptr, c.sc(X.d_hn), G(QP_FFN_DW_W),
                      cfg->dwconv_bias ? G(QP_FFN_DW_B) : nullptr, nullptr));
    QV_TRY(gelu_bwd(st, dt, c.sv(S.h_pre), c.sc(X.d_hn), (long)R * D.fh, c.sc(X.d_hpre)));
  }
  QV_TRY(gemm_tn(st, dt, c.sc(X.d_hpre), D.fh, c.sv(S.y), d, R, D.fh, d, G(QP_FFN_FC1_W), G(QP_FFN_FC1_B), nullptr));
  QV_TRY(gemm_nn(st, dt, c.sc(X.d_hpre), D.fh, R, c.W(W_F1, c.pf(QP_FFN_FC1_W)), epi_t(c, nullptr, c.sc(X.d_y), d)));
  // d_x1 = dblk + LN2-backward(d_y)   (fp32 and a T copy for the next GEMMs)
  // (with dropout: the T copy, which only the bottleneck GEMMs read, goes through the B2 site and DropPath scale)
  {
    const DropP site = dc.site(DS_B2);
    QV_TRY(ln_bwd(st, QV_F32, c.sv(S.x1), d, dt, c.sc(X.d_y), d, R, d, c.pf(QP_NORM2_W), c.svf(S.n2_stats), 0, dt, c.sc(X.d_x1t),
                  c.scf(X.d_x1), dblk, G(QP_NORM2_W), G(QP_NORM2_B), &site, dc.rs1, D.Nt));
  }

  // ---- BottleneckMLP backward
  QV_TRY(gemm_tn(st, dt, c.sc(X.d_x1t), d, c.sv(S.h1), D.bh, R, d, D.bh, G(QP_BMLP_FC2_W), G(QP_BMLP_FC2_B), nullptr));
  {
    const Weight wb2 = c.W(W_B2, c.pf(QP_BMLP_FC2_W));
    GemmEpi e = epi_t(c, nullptr, c.sc(X.d_h1), D.bh);
    const bool fused = dc.drop && fused_nn(c, wb2, R, d);
    if (fused) e.drop = dc.site(DS_B1);
    QV_TRY(gemm_nn(st, dt, c.sc(X.d_x1t), d, R, wb2, e));
    if (dc.drop && !fused) QV_TRY(drop_inplace(c, dc, c.sc(X.d_h1), D.bh, DS_B1, nullptr));
  }
  QV_TRY(gelu_bwd(st, dt, c.sv(S.h1_pre), c.sc(X.d_h1), (long)R * D.bh, c.sc(X.d_h1pre)));
  QV_TRY(gemm_tn(st, dt, c.sc(X.d_h1pre), D.bh, c.sv(S.fused), d, R, D.bh, d, G(QP_BMLP_FC1_W), G(QP_BMLP_FC1_B), nullptr));
  QV_TRY(gemm_nn(st, dt, c.sc(X.d_h1pre), D.bh, R, c.W(W_B1, c.pf(QP_BMLP_FC1_W)), epi_t(c, nullptr, c.sc(X.d_fused), d)));
  // ---- fusion weights
  QV_TRY(fusion_bwd(st, dt, c.sc(X.d_fused), c.sv(S.fused), R, 4, D.cd, c.svf(S.alpha), c.scf(X.draw)));
  QV_TRY(fusion_bwd_final(st, c.svf(S.alpha), c.scf(X.draw), 4, G(QP_FUSION_W)));

  static const int kNormW[4] = {QP_NSWA_W, QP_NMSDA_W, QP_NCGA_W, QP_NCROSS_W};
  static const int kCompW[4] = {QP_CSWA_W, QP_CMSDA_W, QP_CCGA_W, QP_CCROSS_W};
  float* d_xn = c.scf(X.d_xn);
  GemmEpi acc_xn;
  acc_xn.C = d_xn; acc_xn.ldc = d; acc_xn.c_f32 = 1; acc_xn.c_accum = 1;

  void* d_branch_buf[4] = {c.sc(X.d_branch), c.sc(X.d_branch), c.sc(X.d_branch), c.sc(X.d_branch)};
  if (D.fcmp) {   // compress_i and norm_i backward of all four branches in one launch (dropout site of each branch output applied)
    const void* xs[4]; const float *sts[4], *gs[4], *bs[4], *ws[4]; void* dxs[4]; float *dws[4], *dbs[4], *dgs[4], *dbt[4];
    DropP sites[4];
    for (int i = 0; i < 4; ++i) {
      if (i > 0) d_branch_buf[i] = c.sc(X.d_branch3[i - 1]);
      xs[i] = c.sv(S.branch[i]); sts[i] = c.svf(S.nb_stats[i]); gs[i] = c.pf(kNormW[i]); bs[i] = c.pf(kNormW[i] + 1); ws[i] = c.pf(kCompW[i]);
      dxs[i] = d_branch_buf[i]; dws[i] = G(kCompW[i]); dbs[i] = G(kCompW[i] + 1); dgs[i] = G(kNormW[i]); dbt[i] = G(kNormW[i] + 1);
      sites[i] = dc.site(DS_PROJ + i);
    }
    QV_TRY(cmpf_bwd(st, R, xs, sts, gs, bs, ws, c.svf(S.alpha), c.sc(X.d_fused), dxs, dws, dbs, dgs, dbt, sites));
  }
  for (int i = 3; i >= 0; --i) {
    if (!D.fcmp) {
      // compress_i and norm_i backward -> d_branch
      const uint8_t* dfs = static_cast<const uint8_t*>(c.sc(X.d_fused)) + (size_t)i * D.cd * D.ts;
      const float* alpha_i = c.svf(S.alpha) + i;
      QV_TRY(gemm_tn(st, dt, dfs, d, c.sv(S.nb[i]), d, R, D.cd, d, G(kCompW[i]), G(kCompW[i] + 1), alpha_i));
      GemmEpi e = epi_t(c, nullptr, c.sc(X.d_nb), d);
      e.scale_pre = alpha_i;
      QV_TRY(gemm_nn(st, dt, dfs, d, R, c.W(W_C0 + i, c.pf(kCompW[i])), e));
      // d_branch goes through the branch's output-dropout site inside the LayerNorm backward
      const DropP site = dc.site(DS_PROJ + i);
      QV_TRY(ln_bwd(st, dt, c.sv(S.branch[i]), d, dt, c.sc(X.d_nb), d, R, d, c.pf(kNormW[i]), c.svf(S.nb_stats[i]), 0, dt,
                    c.sc(X.d_branch), nullptr, nullptr, G(kNormW[i]), G(kNormW[i] + 1), &site));
    }
    const void* d_branch = d_branch_buf[i];
    if (i == 3) {  // ---- cross
      QV_TRY(gemm_tn(st, dt, d_branch, d, c.sv(S.attn_cross), d, R, d, d, G(QP_CROSS_PROJ_W), G(QP_CROSS_PROJ_B), nullptr));
      QV_TRY(gemm_nn(st, dt, d_branch, d, R, c.W(W_CROSS_PROJ, c.pf(QP_CROSS_PROJ_W)), epi_t(c, nullptr, c.sc(X.d_attn), d)));
      AttnP p = attn_params(c, 2);
      p.q = c.sv(S.q_cross); p.ldq = D.ldq1; p.qcol = 0;
      p.kc = c.svf(S.Kc); p.vc = c.svf(S.Vc);
      p.dout = c.sc(X.d_attn); p.lddo = d;
      p.dq = c.sc(X.d_qc); p.lddq = D.ldq1; p.dqcol = 0;
      p.dbank_k = c.scf(X.dKc); p.dbank_v = c.scf(X.dVc);
      p.drop = dc.site(DS_ATT + 3);
      QV_TRY(attn_bwd(st, dt, p));
      if (!D.xstk) {
        QV_TRY(gemm_tn(st, dt, c.sc(X.d_qc), d, c.sv(S.xn), d, R, d, d, G(QP_CROSS_Q_W), G(QP_CROSS_Q_B), nullptr));
        QV_TRY(gemm_nn(st, dt, c.sc(X.d_qc), d, R, c.W(W_CROSS_Q, c.pf(QP_CROSS_Q_W)), acc_xn));
      }
      QV_TRY(small_linear_bwd2(st, D.kb, d, d, snap_k(c, 3), c.pf(QP_CROSS_K_W), c.scf(X.dKc), G(QP_CROSS_K_W), G(QP_CROSS_K_B), G(QP_BANK_K),
                               snap_v(c, 3), c.pf(QP_CROSS_V_W), c.scf(X.dVc), G(QP_CROSS_V_W), G(QP_CROSS_V_B), G(QP_BANK_V)));
    } else if (i == 2) {  // ---- CGA
      QV_TRY(gemm_tn(st, dt, d_branch, d, c.sv(S.attn_cga), d / 2, R, d, d / 2, G(QP_CGA_PROJ_W), G(QP_CGA_PROJ_B), nullptr));
      QV_TRY(gemm_nn(st, dt, d_branch, d, R, c.W(W_CGA_PROJ, c.pf(QP_CGA_PROJ_W)), epi_t(c, nullptr, c.sc(X.d_attn), d / 2)));
      CgaP p{};
      p.B = D.B; p.Nt = D.Nt; p.G = D.G; p.H = D.H; p.kb = D.kb; p.cg = D.cg; p.cpg = D.cpg;
      p.xn = c.sv(S.xn); p.ldx = d;
      p.Wq = c.pf(QP_CGA_Q_W); p.bq = c.pf(QP_CGA_Q_B); p.Wk = c.pf(QP_CGA_K_W); p.bk = c.pf(QP_CGA_K_B);
      p.Wv = c.pf(QP_CGA_V_W); p.bv = c.pf(QP_CGA_V_B);
      p.kbp = c.svf(S.kbp); p.vbp = c.svf(S.vbp);
      p.dout = c.sc(X.d_attn); p.lddo = d / 2;
      p.dxn = d_xn; p.lddx = d;
      p.dWq = G(QP_CGA_Q_W); p.dbq = G(QP_CGA_Q_B); p.dWk = G(QP_CGA_K_W); p.dbk = G(QP_CGA_K_B);
      p.dWv = G(QP_CGA_V_W); p.dbv = G(QP_CGA_V_B);
      p.dkbp = c.scf(X.dkbp); p.dvbp = c.scf(X.dvbp);
      p.drop = dc.site(DS_ATT + 2);
      QV_TRY(cga_bwd(st, dt, p));
      QV_TRY(small_linear_bwd2(st, D.kb, d, D.cpg, snap_k(c, 2), c.pf(QP_CGA_BK_W), c.scf(X.dkbp), G(QP_CGA_BK_W), G(QP_CGA_BK_B), G(QP_BANK_K),
                               snap_v(c, 2), c.pf(QP_CGA_BV_W), c.scf(X.dvbp), G(QP_CGA_BV_W), G(QP_CGA_BV_B), G(QP_BANK_V)));
    } else if (i == 1) {  // ---- MSDA
      QV_TRY(gemm_tn(st, dt, d_branch, d, c.sv(S.attn_msda), d, R, d, d, G(QP_MSDA_PROJ_W), G(QP_MSDA_PROJ_B), nullptr));
      QV_TRY(gemm_nn(st, dt, d_branch, d, R, c.W(W_MSDA_PROJ, c.pf(QP_MSDA_PROJ_W)), epi_t(c, nullptr, c.sc(X.d_attn), d)));
      AttnP p = attn_params(c, 1);
      p.q = c.sv(S.q_msda); p.ldq = D.ldq1; p.qcol = 0;
      p.kv = c.sv(S.kv_msda); p.ldkv = 2 * d; p.kcol = 0; p.vcol = d;
      p.Ek = c.pf(QP_MSDA_EK); p.Ev = c.pf(QP_MSDA_EV);
      p.bank_k = snap_k(c, 1); p.bank_v = snap_v(c, 1);
      p.dout = c.sc(X.d_attn); p.lddo = d;
      p.dq = c.sc(X.d_q); p.lddq = D.ldq1; p.dqcol = 0;
      p.dkv = c.sc(X.d_kv); p.lddkv = 2 * d; p.dkcol = 0; p.dvcol = d;
      p.dEk = G(QP_MSDA_EK); p.dEv = G(QP_MSDA_EV); p.dbank_k = G(QP_BANK_K); p.dbank_v = G(QP_BANK_V);
      if (dt == QV_BF16 && D.Nt > 16) p.wsp = c.sc(X.attn_ws_b);
      p.drop = dc.site(DS_ATT + 1);
      QV_TRY(attn_bwd(st, dt, p));
      if (!D.xstk) {
        QV_TRY(gemm_tn(st, dt, c.sc(X.d_q), d, c.sv(S.xn), d, R, d, d, G(QP_MSDA_QKV_W), G(QP_MSDA_QKV_B), nullptr));
        QV_TRY(gemm_nn(st, dt, c.sc(X.d_q), d, R, c.W(W_MSDA_Q, c.pf(QP_MSDA_QKV_W)), acc_xn));
      }
      QV_TRY(gemm_tn(st, dt, c.sc(X.d_kv), 2 * d, c.sv(S.xp), d, D.B * D.NM, 2 * d, d, G(QP_MSDA_QKV_W) + (size_t)d * d,
                     G(QP_MSDA_QKV_B) + d, nullptr));
      QV_TRY(gemm_nn(st, dt, c.sc(X.d_kv), 2 * d, D.B * D.NM, c.W(W_MSDA_KV, weight_src(c, W_MSDA_KV)), epi_t(c, nullptr, c.sc(X.d_xp), d)));
      QV_TRY(msda_pool_bwd(st, dt, c.sc(X.d_xp), D.B, D.Nt, D.side, d, cfg->dilations, cfg->n_dilations, cfg->pool_stride, D.NM, d_xn));
    } else {  // ---- SWA
      QV_TRY(gemm_tn(st, dt, d_branch, d, c.sv(S.attn_swa), d, R, d, d, G(QP_SWA_PROJ_W), G(QP_SWA_PROJ_B), nullptr));
      QV_TRY(gemm_nn(st, dt, d_branch, d, R, c.W(W_SWA_PROJ, c.pf(QP_SWA_PROJ_W)), epi_t(c, nullptr, c.sc(X.d_attn), d)));
      AttnP p = attn_params(c, 0);
      p.q = c.sv(S.qkv_swa); p.ldq = D.ldp; p.qcol = 0;
      p.kv = c.sv(S.qkv_swa); p.ldkv = D.ldp; p.kcol = d; p.vcol = 2 * d;
      p.Ek = c.pf(QP_SWA_EK); p.Ev = c.pf(QP_SWA_EV);
      p.bank_k = snap_k(c, 0); p.bank_v = snap_v(c, 0);
      p.dout = c.sc(X.d_attn); p.lddo = d;
      p.dq = c.sc(X.d_qkv); p.lddq = D.ldp; p.dqcol = 0;
      p.dkv = c.sc(X.d_qkv); p.lddkv = D.ldp; p.dkcol = d; p.dvcol = 2 * d;
      p.dEk = G(QP_SWA_EK); p.dEv = G(QP_SWA_EV); p.dbank_k = G(QP_BANK_K); p.dbank_v = G(QP_BANK_V);
      p.drop = dc.site(DS_ATT + 0);
      QV_TRY(attn_bwd(st, dt, p));
      if (!D.xstk) {
        QV_TRY(gemm_tn(st, dt, c.sc(X.d_qkv), 3 * d, c.sv(S.xn), d, R, 3 * d, d, G(QP_SWA_QKV_W), G(QP_SWA_QKV_B), nullptr));
        QV_TRY(gemm_nn(st, dt, c.sc(X.d_qkv), 3 * d, R, c.W(W_SWA_QKV, c.pf(QP_SWA_QKV_W)), acc_xn));
      }
    }
  }
  if (D.xstk) {   // [d_qkv_swa | d_q_msda | d_q_cross] against norm1's output: one dW GEMM scattered to the three parameters, one dX GEMM
    TnSegs sg;
    sg.n = 3;
    sg.end[0] = 3 * d; sg.end[1] = 4 * d; sg.end[2] = 5 * d;
    sg.dW[0] = G(QP_SWA_QKV_W); sg.dW[1] = G(QP_MSDA_QKV_W); sg.dW[2] = G(QP_CROSS_Q_W);
    sg.db[0] = G(QP_SWA_QKV_B); sg.db[1] = G(QP_MSDA_QKV_B); sg.db[2] = G(QP_CROSS_Q_B);
    QV_TRY(tc_gemm_tn_seg(st, static_cast<const bf16*>(c.sc(X.d_qkv)), 5 * d, static_cast<const bf16*>(c.sv(S.xn)), d, R, 5 * d, d, sg));
    QV_TRY(gemm_nn(st, dt, c.sc(X.d_qkv), 5 * d, R, c.W(W_XSTK, nullptr), acc_xn));
  }

  // ---- norm1 backward: d_x = d_x1 + LN1-backward(d_xn)
  float* dxc = D.tl ? c.scf(X.d_blk) : dx;   // d_blk is free again (consumed by the FFN backward)
  QV_TRY(ln_bwd(st, QV_F32, x, d, QV_F32, d_xn, d, R, d, c.pf(QP_NORM1_W), c.svf(S.n1_stats), 0, QV_F32, nullptr, dxc,
                c.scf(X.d_x1), G(QP_NORM1_W), G(QP_NORM1_B)));

  // ---- TokenLearner backward
  if (D.ftok) {
    QV_TRY(tlf_bwd(st, x_in, c.svf(S.tl_S), c.svf(S.tl_Z), dxc, D.B, D.Nf, c.pf(QP_TL_LN_W), c.pf(QP_TL_LN_B), c.pf(QP_TL_FC_W), 1e-5f, dx,
                   G(QP_TL_FC_W), G(QP_TL_FC_B), G(QP_TL_LN_W), G(QP_TL_LN_B)));
  } else if (D.tl) {
    const int tdt = D.tdt;
    QV_TRY(token_learner_bwd(st, tdt, x_in, c.svf(S.tl_S), dxc, D.B, D.Nf, D.Nt, d, c.sc(X.d_logits), c.scf(X.d_up)));   // d_up is free again
    QV_TRY(gemm_tn(st, tdt, c.sc(X.d_logits), D.Nt, c.sv(S.tl_ln), d, D.Rf, D.Nt, d, G(QP_TL_FC_W), G(QP_TL_FC_B), nullptr));
    GemmEpi et = epi_t(c, nullptr, c.sc(X.d_tl_ln), d);
    et.c_f32 = tdt == QV_F32;
    QV_TRY(gemm_nn(st, tdt, c.sc(X.d_logits), D.Nt, D.Rf, c.W(W_TL, c.pf(QP_TL_FC_W)), et));
    QV_TRY(ln_bwd(st, QV_F32, x_in, d, tdt, c.sc(X.d_tl_ln), d, D.Rf, d, c.pf(QP_TL_LN_W), c.svf(S.tl_stats), 0, QV_F32, nullptr, dx,
                  c.scf(X.d_up), G(QP_TL_LN_W), G(QP_TL_LN_B)));
  }
  return 0;
}

// =====================================================================================================
// thin wrappers
// =====================================================================================================
extern "C" size_t qavit_patch_embed_scratch_bytes(int B, int Cin, int S, int p, int d) {
  return patch_embed_scratch_bytes(B, Cin, S, p, d) + 256;
}
extern "C" int qavit_patch_embed_forward(const float* img, int B, int Cin, int S, int p, int d, const float* W,
                                         const float* bias, const float* ln_w, const float* ln_b, const float* pos,
                                         float* pre, float* stats, float* out, int dtype, void* scratch, void* stream) {
  return patch_embed_fwd((cudaStream_t)stream, dtype, img, B, Cin, S, p, d, W, bias, ln_w, ln_b, pos, pre, stats, out, scratch);
}
extern "C" int qavit_patch_embed_backward(const float* img, const float* dout, int B, int Cin, int S, int p, int d,
                                          const float* pre, const float* stats, const float* ln_w, float* dpre_scratch,
                                          float* dW, float* dbias, float* dln_w, float* dln_b, float* dpos, int dtype,
                                          void* scratch, void* stream) {
  return patch_embed_bwd((cudaStream_t)stream, dtype, img, dout, B, Cin, S, p, d, pre, stats, ln_w, dpre_scratch, dW, dbias, dln_w, dln_b,
                         dpos, scratch);
}
extern "C" int qavit_head_forward(const float* x, int B, int N, int d, const float* ln_w, const float* ln_b, const float* W,
                                  const float* bias, int classes, float* stats, float* pooled, float* logits, void* stream) {
  return head_fwd((cudaStream_t)stream, x, B, N, d, ln_w, ln_b, W, bias, classes, stats, pooled, logits);
}
extern "C" int qavit_head_backward(const float* x, const float* dlogits, int B, int N, int d, const float* ln_w,
                                   const float* stats, const float* pooled, const float* W, int classes,
                                   float* dpooled_scratch, float* dx, float* dln_w, float* dln_b, float* dW, float* dbias,
                                   void* stream) {
  return head_bwd((cudaStream_t)stream, x, dlogits, B, N, d, ln_w, stats, pooled, W, classes, dpooled_scratch, dx, dln_w, dln_b, dW, dbias);
}
// nn.Dropout on a contiguous fp32 tensor (pos_drop, H:1155 / 1251).  Forward snapshots {seed, offset} into snap and advances the
// offset on the device; backward regenerates the mask from snap.
extern "C" int qavit_dropout_forward(const float* x, float* y, long long n, float p, unsigned long long* rng,
                                     unsigned long long* snap, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  QV_CHECK(x && y && rng && snap && n % 8 == 0 && p >= 0.f && p < 1.f, "dropout_forward: bad argument (n %% 8 == 0, 0 <= p < 1)");
  QV_TRY(rng_snapshot_advance(st, rng, snap));
  if (y != x) QV_CUDA(cudaMemcpyAsync(y, x, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
  DropP d; d.p = p; d.rng = snap; d.site = 0x9051u;
  return drop_rows(st, QV_F32, y, 8, n / 8, 8, d, nullptr, 1, nullptr, 0, nullptr, nullptr, 0);
}
extern "C" int qavit_dropout_backward(const float* dy, float* dx, long long n, float p, const unsigned long long* snap, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  QV_CHECK(dy && dx && snap && n % 8 == 0, "dropout_backward: bad argument");
  if (dx != dy) QV_CUDA(cudaMemcpyAsync(dx, dy, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
  DropP d; d.p = p; d.rng = snap; d.site = 0x9051u;
  return drop_rows(st, QV_F32, dx, 8, n / 8, 8, d, nullptr, 1, nullptr, 0, nullptr, nullptr, 0);
}
extern "C" int qavit_cross_entropy(const float* logits, const long long* ya, const long long* yb, float lam, const float* lam_dev,
                                   int B, int classes, float label_smoothing, float* loss, float* dlogits, float* row_loss,
                                   int* err_flag, void* stream) {
  QV_CHECK(logits && ya && loss, "cross_entropy: null argument");
  return ce_loss_fwd_bwd((cudaStream_t)stream, logits, ya, yb, lam, lam_dev, B, classes, label_smoothing, loss, dlogits, row_loss, err_flag);
}
// y = x * (*scalar_dev): the chain-rule factor of a scalar loss (dlogits * dloss) without an ATen kernel
extern "C" int qavit_scale_by_scalar(const float* x, const float* scalar_dev, long long n, float* y, void* stream) {
  QV_CHECK(x && scalar_dev && y, "scale_by_scalar: null argument");
  return scale_by_scalar((cudaStream_t)stream, x, scalar_dev, (long)n, y);
}
extern "C" int qavit_memset_zero(void* p, size_t bytes, void* stream) {
  QV_CHECK(p || bytes == 0, "memset_zero: null argument");
  if (bytes) QV_CUDA(cudaMemsetAsync(p, 0, bytes, (cudaStream_t)stream));
  return 0;
}
extern "C" int qavit_test_gemm_nt(int use_tc, const void* A, int lda, int M, int N, int K, const float* W, const void* Wb,
                                  const float* bias, void* C, int c_f32, void* stream) {
  GemmEpi e;
  e.bias = bias; e.C = C; e.ldc = N; e.c_f32 = c_f32;
  if (use_tc) return tc_gemm_nt((cudaStream_t)stream, (const bf16*)A, lda, M, N, K, (const bf16*)Wb, e);
  return simt_gemm_nt((cudaStream_t)stream, QV_F32, A, lda, M, N, K, W, e);
}
// mode 0: C = A W^T + b | 1: C = pre, C2 = gelu(pre) | 2: C2 = aux + pre (aux bf16 residual) | 3: C = pre * gelu'(aux)
//      4: C = gelu'(pre), C2 = gelu(pre) | 5: C = pre * aux
extern "C" int qavit_test_gemm_epi(const void* A, int lda, int M, int N, int K, const void* Wb, const float* bias, void* C,
                                   void* C2, int mode, const void* aux, void* stream) {
  GemmEpi e;
  e.bias = bias;
  if (mode != 2) { e.C = C; e.ldc = N; }
  if (mode == 1 || mode == 4) { e.gelu = 1; e.C2 = C2; e.ldc2 = N; e.gelu_dgrad = mode == 4; }
  if (mode == 2) { e.resid = aux; e.ldr = N; e.r_bf16 = 1; e.C2 = C2; e.ldc2 = N; }
  if (mode == 3 || mode == 5) { e.gmul = aux; e.ldg = N; e.g_bf16 = 1; e.gmul_raw = mode == 5; }
  return tc_gemm_nt((cudaStream_t)stream, (const bf16*)A, lda, M, N, K, (const bf16*)Wb, e);
}
extern "C" int qavit_test_gemm_tn(int use_tc, const void* dY, int ldy, const void* X, int ldx, int M, int N, int K,
                                  float* dW, float* db, void* stream) {
  if (use_tc) return gemm_tn((cudaStream_t)stream, QV_BF16, dY, ldy, X, ldx, M, N, K, dW, db, nullptr);   // incl. the wide-K transposed flavour
  return simt_gemm_tn((cudaStream_t)stream, QV_F32, dY, ldy, X, ldx, M, N, K, dW, db, nullptr);
}
// The fused TokenLearner / TokenUpMix kernels of bf16 runs (16 learned tokens, <= 64 stream tokens, 192 channels) on their own:
// op 0 = TokenLearner forward, 1 = TokenLearner backward, 2 = TokenUpMix (+ LayerNorm) forward, 3 = its backward.  All fp32, device:
//   0: in  {x, ln_w, ln_b, W[16, C], b[16]}            out {S[B, N, 16], xc[B, 16, C], Z[B, N, 16]}
//   1: in  {x, S, dxc, ln_w, ln_b, W, Z}                out {dx[B, N, C], dW, db, dln_w, dln_b}   (parameter gradients accumulated)
//   2: in  {xc, W[N, 16], b[N], ln_w, ln_b}             out {out[B, N, C], stats[B N, 2]}
//   3: in  {xc, dout, stats, W, b, ln_w}                out {dxc[B, 16, C], dW, dln_w, dln_b}     (parameter gradients accumulated)
extern "C" int qavit_test_tokens_fused(int op, int B, int N, int C, const float* const* in, float* const* out, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  QV_CHECK(in && out, "tokens_fused: null argument");
  QV_CHECK(tokens_fused_ok(16, N, C), "tokens_fused: N=%d C=%d not covered (N multiple of 16 <= 64, C = 192)", N, C);
  switch (op) {
    case 0: return tlf_fwd(s, in[0], B, N, in[1], in[2], in[3], in[4], 1e-5f, out[0], out[2], out[1]);
    case 1: return tlf_bwd(s, in[0], in[1], in[6], in[2], B, N, in[3], in[4], in[5], 1e-5f, out[0], out[1], out[2], out[3], out[4]);
    case 2: return upf_fwd(s, in[0], B, N, in[1], in[2], in[3], in[4], 1e-5f, out[0], out[1]);
    case 3: return upf_bwd(s, in[0], in[1], in[2], B, N, in[3], in[4], in[5], out[0], out[1], out[2], out[3]);
  }
  qv_set_error("tokens_fused: op %d", op);
  return 1;
}
// The fused branch-LayerNorm + compress kernels (cmp_fused.cu; d = 192, compress_dim = 48) on their own.  op 0 = forward, 1 = backward.
//   0: in  {x0..x3 (bf16 [R, 192]), gamma0..3, beta0..3, W0..3 ([48, 192]), b0..3, alpha[4]}   out {fused (bf16 [R, 192]), stats0..3 ([R, 2])}
//   1: in  {x0..x3, stats0..3, gamma0..3, beta0..3, W0..3, alpha[4], dfused (bf16 [R, 192])}   out {dx0..3 (bf16), dW0..3, db0..3, dgamma0..3, dbeta0..3}
// op 0: forward (in: h_pre, g1, b1, w, bias, scale, g2, b2; out: hn2, stats1, stats2); op 1: backward (in: + d_hn2, stats1, stats2;
// out: d_hpre, dg1, db1, dw, dbias, dscale, dg2, db2)
extern "C" int qavit_test_ffn_mid(int op, int B, int side, int C, const void* const* in, void* const* out, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto f = [&](int i) { return static_cast<const float*>(in[i]); };
  auto o = [&](int i) { return static_cast<float*>(out[i]); };
  QV_CHECK(ffn_mid_ok(side, C), "ffn_mid: side=%d C=%d not covered", side, C);
  if (op == 0) return ffn_mid_fwd(s, in[0], B, side, C, f(1), f(2), f(3), f(4), f(5), f(6), f(7), 1e-5f, out[0], o(1), o(2));
  return ffn_mid_bwd(s, in[0], in[8], f(9), f(10), B, side, C, f(1), f(2), f(3), f(4), f(5), f(6), out[0], o(1), o(2), o(3), o(4), o(5), o(6), o(7));
}

extern "C" int qavit_test_cmp_fused(int op, long long R, const void* const* in, void* const* out, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  QV_CHECK(in && out, "cmp_fused: null argument");
  if (op == 0) {
    return cmpf_fwd(s, (long)R, in, (const float* const*)(in + 4), (const float* const*)(in + 8), (const float* const*)(in + 12),
                    (const float* const*)(in + 16), (const float*)in[20], 1e-5f, out[0], (float* const*)(out + 1));
  }
  if (op == 1) {
    return cmpf_bwd(s, (long)R, in, (const float* const*)(in + 4), (const float* const*)(in + 8), (const float* const*)(in + 12),
                    (const float* const*)(in + 16), (const float*)in[20], in[21], out, (float* const*)(out + 4), (float* const*)(out + 8),
                    (float* const*)(out + 12), (float* const*)(out + 16), nullptr);
  }
  qv_set_error("cmp_fused: op %d", op);
  return 1;
}
extern "C" int qavit_convert_weight(const float* w, int N, int K, void* wb, void* wbt, void* stream) {
  return convert_weight((cudaStream_t)stream, w, N, K, (bf16*)wb, (bf16*)wbt);
}

// nn.LayerNorm forward / backward for the modules around the blocks (SplitFusion, LMFAdapter, RRCV, ConvNeXt; C <= 256).
// x: fp32 (x_bf16 = 0) or bf16; y and dy are fp32 (autocast keeps LayerNorm outputs in fp32);
// dx has x's dtype; dgamma / dbeta are accumulated.
extern "C" int qavit_layer_norm_forward(const void* x, int x_bf16, long long rows, int C, const float* w, const float* b,
                                        float eps, float* y, float* stats, void* stream) {
  return ln_fwd((cudaStream_t)stream, x_bf16 ? QV_BF16 : QV_F32, x, C, (int)rows, C, w, b, eps, 0, nullptr, nullptr, QV_F32, y, C, stats);
}
extern "C" int qavit_layer_norm_backward(const void* x, int x_bf16, const float* dy, long long rows, int C, const float* w,
                                         const float* stats, void* dx, float* dgamma, float* dbeta, void* stream) {
  return ln_bwd((cudaStream_t)stream, x_bf16 ? QV_BF16 : QV_F32, x, C, QV_F32, dy, C, (int)rows, C, w, stats, 0,
                x_bf16 ? QV_BF16 : QV_F32, x_bf16 ? dx : nullptr, x_bf16 ? nullptr : (float*)dx, nullptr, dgamma, dbeta);
}

// nn.Linear / 1x1 convolution on row-major activations (the lateral path's pointwise layers: ConvNeXt pwconv1/2 H:724-726,
// LMFAdapter.proj H:815, RRCV reverse/reembed_proj H:866-874, SplitFusion gate_fc / cat_mlp.0 H:923-927, cnn_stem 1x1s).
// y[M, N] = x[M, K] W^T + b through the same GEMM flavours as the block (tcgen05 when is_bf16, fp32 SIMT otherwise).
// wb_scratch: N*K bf16 (forward) / wbt_scratch: N*K bf16 (backward) for the converted weight; unused in fp32 runs.
extern "C" int qavit_linear_forward(const void* x, int is_bf16, long long M, int K, const float* W, const float* bias, int N,
                                    void* y, void* wb_scratch, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const int dt = is_bf16 ? QV_BF16 : QV_F32;
  Weight w;
  w.w = W; w.N = N; w.K = K;
  if (is_bf16 && tc_shape_ok_nt((int)M, N, K, K)) {
    QV_CHECK(wb_scratch, "linear_forward: bf16 run needs wb_scratch");
    QV_TRY(convert_weight(s, W, N, K, (bf16*)wb_scratch, nullptr));
    w.wb = (const bf16*)wb_scratch;
  }
  GemmEpi e;
  e.bias = bias; e.C = y; e.ldc = N; e.c_f32 = !is_bf16;
  return gemm_nt(s, dt, x, K, (int)M, w, e);
}
extern "C" int qavit_linear_backward(const void* x, const void* dy, int is_bf16, long long M, int K, int N, const float* W,
                                     void* dx, float* dW, float* db, void* wbt_scratch, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const int dt = is_bf16 ? QV_BF16 : QV_F32;
  QV_TRY(gemm_tn(s, dt, dy, N, x, K, (int)M, N, K, dW, db, nullptr));
  if (dx) {
    Weight w;
    w.w = W; w.N = N; w.K = K;
    if (is_bf16 && tc_shape_ok_nt((int)M, K, N, N)) {
      QV_CHECK(wbt_scratch, "linear_backward: bf16 run needs wbt_scratch");
      QV_TRY(convert_weight(s, W, N, K, nullptr, (bf16*)wbt_scratch));
      w.wbt = (const bf16*)wbt_scratch;
    }
    GemmEpi e;
    e.C = dx; e.ldc = K; e.c_f32 = !is_bf16;
    QV_TRY(gemm_nn(s, dt, dy, N, (int)M, w, e));
  }
  return 0;
}
