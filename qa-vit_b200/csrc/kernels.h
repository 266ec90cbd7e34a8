// Internal launcher interface shared by the .cu files of libqavit_b200.so.
// dt codes: 0 = float32, 1 = bfloat16 ("T" = the activation storage type of the run).
#pragma once
#include "common.cuh"

enum { QV_F32 = 0, QV_BF16 = 1 };

// Epilogue of every GEMM flavour:  val = acc (+ bias[j]);  if (scale_pre) val *= *scale_pre;
//   C  [i, j] (T or fp32; += if c_accum)  = val            (pre-activation when gelu != 0)
//   C2 [i, j] (T or fp32)                 = gelu ? gelu(val) : resid[i, j] + (scale_res ? *scale_res : 1) * val
// gmul (dX GEMMs that feed a GELU backward): val *= gelu'(gmul[i, j]) before C is written; not combined with resid.
// drop / rowscale (tcgen05 flavour only; the callers in block.cu fall back to a drop_rows pass otherwise): a dropout site
// on the output with the element ids of drop_rows on a contiguous [M, N] matrix, and a per-image DropPath scale:
//   gelu == 0: val *= keep(i, j) * rowscale[i / rows_per_img] before C and C2 are formed
//   gelu != 0: only the activation C2 is dropped (C keeps the pre-activation)
struct GemmEpi {
  DropP drop;
  const float* rowscale = nullptr;
  int rows_per_img = 1;
  const float* bias = nullptr;
  const float* scale_pre = nullptr;
  const float* scale_res = nullptr;
  int gelu = 0;
  int gelu_dgrad = 0;            // with gelu: C receives gelu'(pre) instead of pre (the factor the backward multiplies by)
  const void* resid = nullptr;   // fp32, or bf16 when r_bf16
  int ldr = 0;
  int r_bf16 = 0;
  const void* gmul = nullptr;    // fp32, or bf16 when g_bf16: the stored pre-activation
  int ldg = 0;
  int g_bf16 = 0;
  int gmul_raw = 0;              // gmul holds the factor itself (written by a gelu_dgrad forward), not the pre-activation
  void* C = nullptr;
  int ldc = 0;
  int c_f32 = 0;
  int c_accum = 0;
  void* C2 = nullptr;
  int ldc2 = 0;
  int c2_f32 = 0;
};

// A weight as the GEMMs see it: fp32 master [N, K] row-major, plus (bf16 runs) bf16 copies of W and W^T.
struct Weight {
  const float* w = nullptr;   // [N, K]
  const bf16* wb = nullptr;   // [N, K]   bf16 copy      (tcgen05 forward GEMM B operand)
  const bf16* wbt = nullptr;  // [K, N]   bf16 transpose (tcgen05 dX GEMM B operand)
  int N = 0, K = 0;
};

// Y[M, N] = A[M, K] W^T             (A: dt, row stride lda)
int gemm_nt(cudaStream_t s, int dt, const void* A, int lda, int M, const Weight& W, const GemmEpi& e);
// dX[M, K] = dY[M, N] W             (dY: dt, row stride ldy)
int gemm_nn(cudaStream_t s, int dt, const void* dY, int ldy, int M, const Weight& W, const GemmEpi& e);
// dW[N, K] += scale * dY[M, N]^T X[M, K] ; db[N] += scale * colsum(dY)     (fp32 outputs, atomically accumulated)
int gemm_tn(cudaStream_t s, int dt, const void* dY, int ldy, const void* X, int ldx, int M, int N, int K, float* dW,
            float* db, const float* scale);

// fp32 SIMT flavours (always available; the parity-mode GEMM and the small-shape GEMM of bf16 runs)
int simt_gemm_nt(cudaStream_t s, int dtA, const void* A, int lda, int M, int N, int K, const float* W, const GemmEpi& e);
int simt_gemm_nn(cudaStream_t s, int dtA, const void* dY, int ldy, int M, int N, int K, const float* W, const GemmEpi& e);
int simt_gemm_tn(cudaStream_t s, int dt, const void* dY, int ldy, const void* X, int ldx, int M, int N, int K,
                 float* dW, float* db, const float* scale);

// tcgen05 / TMEM / TMA flavours (bf16 operands, fp32 accumulate)
int tc_gemm_nt(cudaStream_t s, const bf16* A, int lda, int M, int N, int K, const bf16* Wb, const GemmEpi& e);
// transposed != 0: the product is written transposed, dW[k * ldw + n] (lets a weight gradient with K > 256 be
// computed as X^T dY, whose accumulator width is N)
// db (optional): column sums of the first operand (db_mode 1) or of the second operand (db_mode 2), db_cols wide,
// accumulated inside the same kernel (the bias gradient; no separate pass over dY).
int tc_gemm_tn(cudaStream_t s, const bf16* dY, int ldy, const bf16* X, int ldx, int M, int N, int K, float* dW,
               const float* scale, int transposed = 0, int ldw = 0, float* db = nullptr, int db_mode = 0, int db_cols = 0);
// dW / db of a STACKED projection: row n of the product belongs to segment i (end[i - 1] <= n < end[i]) and lands in that segment's
// own gradient tensors (the stacked weights are separate parameters)
struct TnSegs { int n = 0; int end[4] = {0, 0, 0, 0}; float* dW[4] = {nullptr, nullptr, nullptr, nullptr}; float* db[4] = {nullptr, nullptr, nullptr, nullptr}; };
int tc_gemm_tn_seg(cudaStream_t s, const bf16* dY, int ldy, const bf16* X, int ldx, int M, int N, int K, const TnSegs& segs);
bool tc_shape_ok_nt(int M, int N, int K, int lda);
bool tc_shape_ok_tn(int M, int N, int K, int ldy, int ldx);
int colsum_accum(cudaStream_t s, int dt, const void* dY, int ldy, int M, int N, float* db, const float* scale);
int convert_weight(cudaStream_t s, const float* w, int N, int K, bf16* wb, bf16* wbt);
// wbt (optional): the transpose, row pitch ldt (0 = N) -- a slice of a wider stacked [K, sum N] matrix when ldt > N.
// bsrc / bdst (optional): bn floats copied alongside (the bias slice of a stacked projection)
struct ConvertJob { const float* w; int N, K; bf16* wb; bf16* wbt; int ldt = 0; const float* bsrc = nullptr; float* bdst = nullptr; int bn = 0; };
struct ConvertJobs { ConvertJob j[24]; int n; };
int convert_weights_batched(cudaStream_t s, const ConvertJobs& jobs);   // all fp32 -> bf16 (+ transposed) copies, one launch

// ---- depthwise k x k stencils on channels-last maps (dwconv_nhwc.cu).  Row pitches in elements; resid / resid2 are
// added to the output, copy receives the input tile (LMFAdapter's identity branch); all optional.
struct DwP {
  const void* x; int ldx;
  int B, H, W, C, K;
  const float* w; const float* bias;
  const float* scale;                     // optional per-channel output scale: y = scale * (conv(x) + bias)
  void* y; int ldy;
  const void* resid; int ldr;
  const void* resid2; int ldr2;
  void* copy; int ldcp;
};
int dw2d_fwd(cudaStream_t s, int dt, const DwP& p, bool flip);
// bf16, 8 x 8 maps, C % 16 == 0, no output scale: the same operation on mma.sync (dwconv_mma.cu); dw2d_fwd dispatches to it
bool dwt_ok(const DwP& p);
int dwt_fwd(cudaStream_t s, const DwP& p, bool flip);
// optional scaled form y = scale * (conv(x) + bias): dw / dbias are scaled, dscale[c] += sum dy * (conv(x) + bias)
struct DwScale { const float* w = nullptr; const float* bias = nullptr; const float* scale = nullptr; float* dscale = nullptr; };
int dw2d_wgrad(cudaStream_t s, int dt, int K, const void* x, int ldx, const void* dy, int lddy, int B, int H, int W, int C,
               float* dw, float* dbias, const DwScale& sc = DwScale());

// ---- lateral path / SplitFusion kernels (lateral_kernels.cu)
// BatchNorm over rows of [rows, C] (+ GELU): mr[2C] = per-channel mean | rstd (kept for backward), sums_scratch[4C] (2C 64-bit fixed-point sums in forward).
int bn_fwd(cudaStream_t s, int dt, const void* x, long rows, int C, const float* gamma, const float* beta, float eps,
           float momentum, int train, float* running_mean, float* running_var, long long* num_batches, int gelu,
           float* sums_scratch, float* mr, void* y);
int bn_bwd(cudaStream_t s, int dt, const void* x, const void* dy, long rows, int C, const float* gamma, const float* beta,
           const float* mr, int train, int gelu, float* sums_scratch, void* dx, float* dgamma, float* dbeta);
// 3x3 stride-2 pad-1 convolution as GEMM: column order k = (ky * 3 + kx) * Cin + cin
int im2col_img(cudaStream_t s, int dt, const float* img, int B, int Cin, int S, int Kp, void* col);
int im2col_nhwc(cudaStream_t s, int dt, const void* x, int B, int Hi, int Cin, void* col);
int col2im_nhwc(cudaStream_t s, int dt, const void* dcol, int B, int Hi, int Cin, void* dx);
int conv_w_pack(cudaStream_t s, const float* W, int N, int Cin, int Kp, float* Wp);
int conv_w_unpack_add(cudaStream_t s, const float* dWp, int N, int Cin, int Kp, float* dW);
// y = [resid +] [*scale *] f(LN(x)), f = GELU when gelu_out
// (x / dx: dt; y / resid: dt_y; y32: optional extra fp32 copy of y; dy: dt_dy)
int rowln_fwd(cudaStream_t s, int dt, const void* x, long rows, int C, const float* gamma, const float* beta, float eps,
              int gelu_out, int dt_y, const void* resid, const float* scale, void* y, float* y32, float* stats);
int rowln_bwd(cudaStream_t s, int dt, const void* x, int dt_dy, const void* dy, long rows, int C, const float* gamma, const float* beta,
              const float* stats, int gelu_out, const float* scale, float* dscale, void* dx, float* dgamma, float* dbeta);
struct SfArgs {
  const float* Tin; const float* R; const void* glin; const void* cpre;
  const float *cat_g, *cat_b, *fin_g, *fin_b, *fw;
  float drop_p; const unsigned long long* rng; uint32_t site;
  long rows; int C;
};
int sf_pre_fwd(cudaStream_t s, int dt, const float* Tin, const float* R, long rows, int C, const float* gamma, const float* beta,
               void* g_in, void* cat, float* stats);
int sf_post_fwd(cudaStream_t s, int dt, const SfArgs& a, float* out, float* cat_stats, float* fin_stats);
int sf_post_bwd(cudaStream_t s, int dt, const SfArgs& a, const float* dout, const float* cat_stats, const float* fin_stats, float* dT,
                float* dR, void* dglin, void* dcpre, float* d_fin_g, float* d_fin_b, float* d_cat_g, float* d_cat_b, float* draw);
int sf_pre_bwd(cudaStream_t s, int dt, const float* Tin, const float* R, const void* dg_in, const void* dcat, long rows, int C,
               const float* gamma, const float* stats, float* dT, float* dR, float* dgamma, float* dbeta);
// F.interpolate(mode = 'bilinear', align_corners = False) of a channels-last [B, Hi, Wi, C] map to [B, Ho, Wo, C] and its backward
// (din is overwritten) -- LMFAdapter's resize, H:839-843
int resize_bilinear_fwd(cudaStream_t s, int dt, const void* in, int B, int Hi, int Wi, int Ho, int Wo, int C, void* out);
int resize_bilinear_bwd(cudaStream_t s, int dt, const void* dout, int B, int Hi, int Wi, int Ho, int Wo, int C, void* din);
int rng_snapshot_advance(cudaStream_t s, unsigned long long* rng, unsigned long long* snap);

// ---- HQAViTv2 stem kernels (stem_v2.cu)
// patch rows [B * (S/p)^2, Cin * p * p] in (c, ky, kx) order -- the conv weight's own order -- for a stride == kernel convolution
int patch_rows(cudaStream_t s, int dt, const float* img, int B, int Cin, int S, int p, void* col);
// nn.LayerNorm([C, H, W]) on a channels-last map [B, HW, C]; gamma / beta keep the [C, H, W] layout; stats[B, 2] = mean | rstd
bool sln_ok(int HW, int C);
int sln_fwd(cudaStream_t s, int dt, const void* x, int B, int HW, int C, const float* gamma, const float* beta, float eps, void* y,
            float* stats);
// dx = [resid +] LN-backward(dy); dx may alias dy; dgamma / dbeta accumulated
int sln_bwd(cudaStream_t s, int dt, const void* x, const void* dy, int B, int HW, int C, const float* gamma, const float* stats,
            const void* resid, void* dx, float* dgamma, float* dbeta);
// LayerScale folded into pwconv2: Ws = diag(gamma) W, bs = gamma * b; finish turns the gradients wrt (Ws, bs) into those of W, b, gamma
int layerscale_prepare(cudaStream_t s, const float* W, const float* b, const float* gamma, int N, int K, float* Ws, float* bs);
int layerscale_finish(cudaStream_t s, const float* G, const float* gb, const float* W, const float* b, const float* gamma, int N, int K,
                      float* dW, float* db, float* dgamma);
// y = [resid +] rowscale[row / rows_per_img] * x   (y may alias x)
int scale_rows(cudaStream_t s, int dt, const void* x, long rows, int C, const float* rowscale, int rows_per_img, const void* resid,
               void* y);
int stem_droppath_scales(cudaStream_t s, const unsigned long long* rng, uint32_t site0, int n_sites, int B, const float* rates, float* rs);

// ---- dropout / DropPath of the quad block (drop.cu).  Masks come from (rng snapshot, site, element index): backward
// calls the same function on the gradient.
//   x[i, j] (T, in place) *= keep(i, j) * (rowscale ? rowscale[i / rows_per_img] : 1)
//   out[i, j] (fp32, optional) = resid[i, j] + (scale ? *scale : 1) * x_new[i, j]
int drop_rows(cudaStream_t s, int dt, void* x, int ldx, long rows, int C, const DropP& d, const float* rowscale, int rows_per_img,
              const float* resid, int ldr, const float* scale, float* out, int ldo);
// DropPath keep scales (H:256-264): rs[b] = floor(1 - p + u_b) / (1 - p) for the two residual branches of a block
int droppath_scales(cudaStream_t s, const DropP& d, int B, float* rs1, float* rs2);

// ---- norms
int ln_fwd(cudaStream_t s, int dt_in, const void* x, int ldx, int rows, int C, const float* gamma, const float* beta,
           float eps, int gelu_in, const float* gamma2, const float* beta2, int dt_out, void* y, int ldy, float* stats);
// drop / rowscale (optional, C % 8 == 0 and C > 128 only): the T output dx_t additionally goes through a dropout site
// (ids of drop_rows on [rows, C]) and a per-image scale -- the gradient of a dropped activation; dx_f32 stays unmasked.
int ln_bwd(cudaStream_t s, int dt_x, const void* x, int ldx, int dt_dy, const void* dy, int lddy, int rows, int C,
           const float* gamma, const float* stats, int gelu_in, int dt_out, void* dx_t, float* dx_f32,
           const float* resid, float* dgamma, float* dbeta, const DropP* drop = nullptr, const float* rowscale = nullptr,
           int rows_per_img = 1);

// ---- attention family
struct AttnP {
  int mode;          // 0 = SWA (windowed, linformer), 1 = MSDA (pooled kv, linformer), 2 = cross (direct bank kv)
  int B, Nt, side, ws, H, hd, kb, klin, L;   // L: rows of E actually contracted (ws^2 | min(NM, 128)); NMrows below
  int NM;            // pooled tokens per image (MSDA kv rows per image)
  const void* q; int ldq; int qcol;          // T
  const void* kv; int ldkv; int kcol, vcol;  // T (mode 0/1); mode 2: fp32 Kc/Vc [kb, H*hd] in kc/vc
  const float* kc; const float* vc;
  const float* Ek; const float* Ev;          // [Lfull, klin]
  const float* bank_k; const float* bank_v;  // [kb, H*hd] snapshot
  void* out; int ldo;                        // T  [rows, H*hd]
  // backward only
  const void* dout; int lddo;                // T
  void* dq; int lddq; int dqcol;             // T
  void* dkv; int lddkv; int dkcol, dvcol;    // T
  float* dEk; float* dEv; float* dbank_k; float* dbank_v;   // fp32 accumulators (mode 2: dKc/dVc in dbank_k/v)
  void* wsp;                                 // optional workspace (attn_msda64_scratch_bytes) for the hoisted Linformer path
  DropP drop;                                // attention-probability dropout (SDPA dropout_p, H:461/525/621); p == 0: off
};
int attn_fwd(cudaStream_t s, int dt, const AttnP& p);
bool attn_mma_ok(const AttnP& p);   // bf16 tensor-core (mma.sync) flavour for the 16-query shapes
int attn_mma_fwd(cudaStream_t s, const AttnP& p);
int attn_mma_bwd(cudaStream_t s, const AttnP& p);
int attn_bwd(cudaStream_t s, int dt, const AttnP& p);
// MSDA with more than 16 query tokens per image on mma.sync, Linformer contraction hoisted (attn_msda64.cu)
bool attn_msda64_ok(const AttnP& p);
size_t attn_msda64_scratch_bytes(int B, int D);
int attn_msda64_fwd(cudaStream_t s, const AttnP& p, void* scratch);
int attn_msda64_bwd(cudaStream_t s, const AttnP& p, void* scratch);

struct CgaP {
  int B, Nt, G, H, kb, cg, cpg;              // cg = 32 channels/group, cpg = 16 compressed/group
  const void* xn; int ldx;                   // T [B*Nt, G*cg]
  const float *Wq, *bq, *Wk, *bk, *Wv, *bv;  // [cpg, cg], [cpg]
  const float *kbp, *vbp;                    // projected bank [kb, cpg] fp32
  void* out; int ldo;                        // T [B*Nt, G*cpg]
  const void* dout; int lddo;                // T
  float* dxn; int lddx;                      // fp32 accumulate [B*Nt, G*cg]
  float *dWq, *dbq, *dWk, *dbk, *dWv, *dbv, *dkbp, *dvbp;
  DropP drop;                                // attention-probability dropout (H:587); p == 0: off
};
int cga_fwd(cudaStream_t s, int dt, const CgaP& p);
int cga_bwd(cudaStream_t s, int dt, const CgaP& p);
bool cga_mma_ok(const CgaP& p);
int cga_mma_fwd(cudaStream_t s, const CgaP& p);
int cga_mma_bwd(cudaStream_t s, const CgaP& p);
bool cga_mma64_ok(const CgaP& p);   // 64-token blocks: CTA per image, warp per 16-query tile (cga_mma64.cu)
int cga_mma64_fwd(cudaStream_t s, const CgaP& p);
int cga_mma64_bwd(cudaStream_t s, const CgaP& p);

// small dense [rows<=64] projections of the bank: Y = X W^T + b and its backward (single CTA, fp32)
int small_linear_fwd(cudaStream_t s, const float* X, int rows, int K, const float* W, const float* b, int N, float* Y);
int small_linear_bwd(cudaStream_t s, const float* X, int rows, int K, const float* W, int N, const float* dY,
                     float* dW, float* db, float* dX_accum);
// two same-shape problems in one launch (the K and V projections of one bank snapshot); outputs must be distinct
int small_linear_fwd2(cudaStream_t s, int rows, int K, int N, const float* X0, const float* W0, const float* b0, float* Y0,
                      const float* X1, const float* W1, const float* b1, float* Y1);
int small_linear_bwd2(cudaStream_t s, int rows, int K, int N, const float* X0, const float* W0, const float* dY0, float* dW0, float* db0,
                      float* dX0, const float* X1, const float* W1, const float* dY1, float* dW1, float* db1, float* dX1);

// ---- bank write (train-mode forward only, no gradient)
int bank_write_reduce(cudaStream_t s, int dt, const void* tn, const void* cg, int ldcg, int B, int Nt, int d, int kb,
                      float* partial, int* n_partial);
int bank_write_apply(cudaStream_t s, const float* partial, int n_partial, int B, int d, int kb, float* bank_k,
                     float* bank_v, long long* update_count, int v1);

// ---- misc per-image kernels
int msda_pool_fwd(cudaStream_t s, int dt, const void* xn, int B, int Nt, int side, int C, const int* dil, int ndil,
                  int stride, int NM, void* xp);
int msda_pool_bwd(cudaStream_t s, int dt, const void* dxp, int B, int Nt, int side, int C, const int* dil, int ndil,
                  int stride, int NM, float* dxn);
int dwconv_fwd(cudaStream_t s, int dt, const void* x, int B, int side, int C, const float* w, const float* bias,
               const float* scale, void* y);
int dwconv_bwd(cudaStream_t s, int dt, const void* x, const void* dy, int B, int side, int C, const float* w,
               const float* bias, const float* scale, void* dx, float* dw, float* dbias, float* dscale);
int gelu_bwd(cudaStream_t s, int dt, const void* pre, const void* dact, long n, void* dpre);
// d_o = gamma * dout (* keep * rowscale when drop / rowscale are given: ids of drop_rows on [n / C, C]); dgamma += sum(dout * o).
// gamma == nullptr: plain cast (QAViT.py's CCFFFN has no gamma), dgamma untouched.
int gamma_bwd(cudaStream_t s, int dt, const float* dout, const void* o, long n, const float* gamma, void* d_o,
              float* dgamma, const DropP* drop = nullptr, const float* rowscale = nullptr, int rows_per_img = 1, int C = 0);
int cast_f32_to_t(cudaStream_t s, int dt, const float* x, long n, void* y);
int copy2_f32(cudaStream_t s, float* d0, const float* s0, int n0, float* d1, const float* s1, int n1);   // two small copies, one launch
int fusion_softmax(cudaStream_t s, const float* w, int n, float* alpha);
int fusion_bwd(cudaStream_t s, int dt, const void* dfused, const void* fused, long rows, int nb, int cw,
               const float* alpha, float* dalpha_raw);
int fusion_bwd_final(cudaStream_t s, const float* alpha, const float* dalpha_raw, int nb, float* dw);
int token_learner_fwd(cudaStream_t s, int dt, const float* x, const void* logits, int B, int N, int M, int C, float* S,
                      float* xc);
int token_learner_bwd(cudaStream_t s, int dt, const float* x, const float* S, const float* dxc, int B, int N, int M,
                      int C, void* dlogits, float* dx);
int token_upmix_fwd(cudaStream_t s, int dt, const float* xc, int B, int M, int N, int C, const float* W, const float* bias,
                    float* up);
int token_upmix_bwd(cudaStream_t s, int dt, const float* xc, const float* dup, int B, int M, int N, int C, const float* W,
                    float* dxc, float* dW, float* dbias);
// tensor-core (mma.sync bf16) flavours for bf16 runs with 16 learned tokens (tokens_mma.cu)
bool tokens_mma_ok(int M, int N, int C);
int tlm_fwd(cudaStream_t s, const float* x, const void* logits, int B, int N, int C, float* S, float* xc);
int tlm_bwd(cudaStream_t s, const float* x, const float* S, const float* dxc, int B, int N, int C, void* dlogits, float* dx);
int upm_fwd(cudaStream_t s, const float* xc, int B, int N, int C, const float* W, const float* bias, float* up);
int upm_bwd(cudaStream_t s, const float* xc, const float* dup, int B, int N, int C, const float* W, float* dxc, float* dW, float* dbias);
bool tokens_mma64_ok(int M, int N, int C);   // 32 .. 64 learned tokens, <= 256 stream tokens
int tlm64_fwd(cudaStream_t s, const float* x, const void* logits, int B, int N, int M, int C, float* S, float* xc);
int tlm64_bwd(cudaStream_t s, const float* x, const float* S, const float* dxc, int B, int N, int M, int C, void* dlogits, float* dx);
int upm64_fwd(cudaStream_t s, const float* xc, int B, int N, int M, int C, const float* W, const float* bias, float* up);
int upm64_bwd(cudaStream_t s, const float* xc, const float* dup, int B, int N, int M, int C, const float* W, float* dxc, float* dW,
              float* dbias);
// fused wrapper kernels for 16 learned / <= 64 stream tokens / 192 channels (tokens_fused.cu): LayerNorm + gate + softmax + pooling,
// up-mix + LayerNorm, and their backwards, one launch each, split-precision (bf16 hi + lo) MMAs, fp32 everywhere else
bool tokens_fused_ok(int M, int N, int C);
// (Z: the LayerNorm-normalised gate projection without its constant term, [B, N, 16] fp32, written by forward for backward)
int tlf_fwd(cudaStream_t s, const float* x, int B, int N, const float* gamma, const float* beta, const float* W, const float* bias,
            float eps, float* S, float* Z, float* xc);
int tlf_bwd(cudaStream_t s, const float* x, const float* S, const float* Z, const float* dxc, int B, int N, const float* gamma,
            const float* beta, const float* W, float eps, float* dx, float* dW, float* dbias, float* dgamma, float* dbeta);
int upf_fwd(cudaStream_t s, const float* xc, int B, int N, const float* W, const float* bias, const float* gamma, const float* beta,
            float eps, float* out, float* stats);
int upf_bwd(cudaStream_t s, const float* xc, const float* dout, const float* stats, int B, int N, const float* W, const float* bias,
            const float* gamma, float* dxc, float* dW, float* dgamma, float* dbeta);
// CCF-FFN mid-section GELU -> LayerNorm -> depthwise 3x3 (* scale) -> LayerNorm as one kernel per direction (ffn_mid.cu): bf16 runs,
// 4 x 4 token maps (register kernel, C a multiple of 32 up to 128) and 8 x 8 maps (shared-memory tile kernel, C a multiple of 8 <= 128).
// stats1 / stats2: (mean, rstd) per row of the two LayerNorms.
bool ffn_mid_ok(int side, int C);
int ffn_mid_fwd(cudaStream_t s, const void* h_pre, int B, int side, int C, const float* g1, const float* b1, const float* w, const float* bias,
                const float* scale, const float* g2, const float* b2, float eps, void* hn2, float* stats1, float* stats2);
int ffn_mid_bwd(cudaStream_t s, const void* h_pre, const void* d_hn2, const float* stats1, const float* stats2, int B, int side, int C,
                const float* g1, const float* b1, const float* w, const float* bias, const float* scale, const float* g2, void* d_hpre,
                float* dg1, float* db1, float* dw, float* dbias, float* dscale, float* dg2, float* db2);
// per-branch LayerNorm + compress Linear + fusion scale + concat for all 4 branches in one launch, and its backward (cmp_fused.cu);
// bf16 runs with d = 192, compress_dim = 48.  x[i]: branch outputs [R, 192] bf16; stats[i]: (mean, rstd) per row (written by fwd)
bool cmp_fused_ok(int d, int cd);
int cmpf_fwd(cudaStream_t s, long R, const void* const* x, const float* const* gamma, const float* const* beta, const float* const* W,
             const float* const* bias, const float* alpha, float eps, void* fused, float* const* stats);
int cmpf_bwd(cudaStream_t s, long R, const void* const* x, const float* const* stats, const float* const* gamma, const float* const* beta,
             const float* const* W, const float* alpha, const void* dfused, void* const* dx, float* const* dW, float* const* db,
             float* const* dgamma, float* const* dbeta, const DropP* drop);
bool bank_write_mma_ok(int Nt, int d, int kb, int ldcg);
int bank_write_reduce_mma(cudaStream_t s, const void* tn, const void* cg, int ldcg, int B, int Nt, int d, float* partial, int* n_partial);
// register-blocked flavours for 16 learned tokens (tokens.cu)
bool tokens16_ok(int M, int C);
int tl16_fwd(cudaStream_t s, int dt, const float* x, const void* logits, int B, int N, int C, float* S, float* xc);
int tl16_bwd(cudaStream_t s, int dt, const float* x, const float* S, const float* dxc, int B, int N, int C, void* dlogits, float* dx);
int up16_fwd(cudaStream_t s, const float* xc, int B, int N, int C, const float* W, const float* bias, float* up);
int up16_bwd(cudaStream_t s, const float* xc, const float* dup, int B, int N, int C, const float* W, float* dxc, float* dW, float* dbias);
// dt == bf16 with scratch (patch_embed_scratch_bytes): patch gather + tcgen05 GEMM; otherwise the fp32 SIMT kernels
size_t patch_embed_scratch_bytes(int B, int Cin, int S, int p, int d);
int patch_embed_fwd(cudaStream_t s, int dt, const float* img, int B, int Cin, int S, int p, int d, const float* W,
                    const float* bias, const float* gamma, const float* beta, const float* pos, float* pre, float* stats,
                    float* out, void* scratch);
int patch_embed_bwd(cudaStream_t s, int dt, const float* img, const float* dout, int B, int Cin, int S, int p, int d,
                    const float* pre, const float* stats, const float* gamma, float* dpre_scratch, float* dW,
                    float* dbias, float* dgamma, float* dbeta, float* dpos, void* scratch);
int head_fwd(cudaStream_t s, const float* x, int B, int N, int d, const float* gamma, const float* beta, const float* W,
             const float* bias, int ncls, float* stats, float* pooled, float* logits);
int head_bwd(cudaStream_t s, const float* x, const float* dlogits, int B, int N, int d, const float* gamma,
             const float* stats, const float* pooled, const float* W, int ncls, float* dpooled_scratch, float* dx,
             float* dgamma, float* dbeta, float* dW, float* dbias);
int ce_loss_fwd_bwd(cudaStream_t s, const float* logits, const long long* ya, const long long* yb, float lam, const float* lam_dev,
                    int B, int ncls, float smoothing, float* loss, float* dlogits, float* row_loss, int* err);
int scale_by_scalar(cudaStream_t s, const float* x, const float* scalar_dev, long n, float* y);   // y = x * (*scalar_dev)
