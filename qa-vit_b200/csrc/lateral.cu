// HQAViT's lateral CNN path (scope row f-1) and SplitFusion as native schedules over the kernels of this directory.
//   qavit_lateral_*     : CNNStemModel (H:742-793) -> 3 x LMFAdapter (H:799-849) -> 3 x RRCV (H:855-907): image in,
//                         the three refined token maps R2 / R3 / R4 out, one call per direction.
//                         stem_kind 1 = HQAViTv2_CIFAR100.py's stem (V:753-833: patchify conv, LayerNorm([C, H, W]), LayerScale
//                         ConvNeXt blocks with DropPath, LayerNorm + 1x1 conv downsampling), same adapters.
//   qavit_splitfusion_* : SplitFusion (H:913-965).
// Every pointwise / 1x1 / 3x3-stride-2 convolution and nn.Linear is a GEMM (tcgen05 in bf16 runs, fp32 SIMT in fp32
// runs) with bias / GELU / residual / GELU-backward epilogues; depthwise stencils, BatchNorm and the LayerNorm-based
// row operations are the HBM-bound kernels of dwconv_nhwc.cu / lateral_kernels.cu.  Activations are channels-last
// [rows = B * H * W, C] in the run's activation type T (bf16 or fp32); tokens ARE the NHWC feature map, so the
// reference's permutes / flattens (H:729, 736, 828, 888, 899) are no-ops here.
#include <stdio.h>
#include <string.h>

#include "../../include/qavit_b200.h"
#include "kernels.h"

namespace {

struct Bump {
  size_t off = 0;
  size_t take(size_t bytes) {
    size_t r = off;
    off += (bytes + 255) & ~(size_t)255;
    return r;
  }
};

// ------------------------------------------------------------------------------------------------ parameter table
// conv+bn unit: 7 entries; ConvNeXt block: 8; LMFAdapter: 8; RRCV: 7 + 8 * blocks
enum { CB_W, CB_B, BN_W, BN_B, BN_RM, BN_RV, BN_NBT, CB_N };
enum { CX_DW_W, CX_DW_B, CX_LN_W, CX_LN_B, CX_W1, CX_B1, CX_W2, CX_B2, CX_N };
enum { LM_DW3_W, LM_DW3_B, LM_DW5_W, LM_DW5_B, LM_PROJ_W, LM_PROJ_B, LM_LN_W, LM_LN_B, LM_N };
enum { RR_REV_W, RR_REV_B, RR_RE_W, RR_RE_B, RR_LN_W, RR_LN_B, RR_BETA, RR_N };

// stem_kind 1 (HQAViTv2): stem conv + spatial LN: 4 entries; ConvNeXt block with LayerScale: 9; downsample (spatial LN + 1x1 conv): 4
enum { S2_CONV_W, S2_CONV_B, S2_LN_W, S2_LN_B, S2_N };
enum { CX2_GAMMA = CX_N, CX2_N };
enum { DS_LN_W, DS_LN_B, DS_W, DS_B, DS_N };
constexpr int V2_BLOCKS = 7;                                   // 2 at c2, 3 at c3, 2 at c4 (V:770-800)
const int kV2Stage[V2_BLOCKS] = {0, 0, 1, 1, 1, 2, 2};         // feature stage of block j

int cb_base(int i) { return CB_N * i; }
int cx_base(int i) { return 4 * CB_N + CX_N * i; }
int v2_cx_base(int j) { return S2_N + CX2_N * j; }
int v2_ds_base(int i) { return S2_N + CX2_N * V2_BLOCKS + DS_N * i; }
int stem_count(const qavit_lateral_cfg& c) { return c.stem_kind == 1 ? S2_N + CX2_N * V2_BLOCKS + 2 * DS_N : 4 * CB_N + 3 * CX_N; }
int lm_base(const qavit_lateral_cfg& c, int i) { return stem_count(c) + LM_N * i; }
int rr_cx_n(const qavit_lateral_cfg& c) { return c.stem_kind == 1 ? CX2_N : CX_N; }   // HQAViTv2's RRCV blocks carry LayerScale too
int rr_stride(const qavit_lateral_cfg& c) { return RR_N + rr_cx_n(c) * c.rrcv_blocks; }
int rr_base(const qavit_lateral_cfg& c, int i) { return stem_count(c) + 3 * LM_N + rr_stride(c) * i; }
int rr_blk(const qavit_lateral_cfg& c, int i, int j) { return rr_base(c, i) + RR_N + rr_cx_n(c) * j; }
int param_count(const qavit_lateral_cfg& c) { return rr_base(c, 3); }

const char* kCbSfx[CB_N] = {"weight", "bias", "weight", "bias", "running_mean", "running_var", "num_batches_tracked"};
const char* kCxSfx[CX_N] = {"dwconv.weight", "dwconv.bias", "norm.weight", "norm.bias", "pwconv1.weight", "pwconv1.bias",
                            "pwconv2.weight", "pwconv2.bias"};
const char* kLmSfx[LM_N] = {"dwconv_3x3.weight", "dwconv_3x3.bias", "dwconv_5x5.weight", "dwconv_5x5.bias", "proj.weight",
                            "proj.bias", "norm.weight", "norm.bias"};
const char* kRrSfx[RR_N] = {"reverse_proj.weight", "reverse_proj.bias", "reembed_proj.weight", "reembed_proj.bias",
                            "norm.weight", "norm.bias", "beta"};

int check_cfg(const qavit_lateral_cfg& c) {
  QV_CHECK(c.batch > 0 && c.img_size % 4 == 0, "lateral: img_size %d must be a multiple of 4", c.img_size);
  QV_CHECK(c.grid >= 1 && c.grid <= c.img_size, "lateral: token grid %d", c.grid);
  QV_CHECK(c.c_stem % 8 == 0 && c.c2 % 8 == 0 && c.c3 % 8 == 0 && c.c4 % 8 == 0 && c.rrcv_channels % 8 == 0 && c.dim % 8 == 0,
           "lateral: channel counts must be multiples of 8");
  QV_CHECK(c.c2 <= 256 && c.c3 <= 256 && c.c4 <= 256 && c.rrcv_channels <= 256 && c.dim <= 256, "lateral: channel counts must be <= 256");
  QV_CHECK(c.rrcv_blocks >= 1 && c.rrcv_blocks <= 4, "lateral: rrcv_blocks %d (1..4)", c.rrcv_blocks);
  QV_CHECK(c.dtype == QV_F32 || c.dtype == QV_BF16, "lateral: dtype %d", c.dtype);
  QV_CHECK(c.stem_kind == 0 || c.stem_kind == 1, "lateral: stem_kind %d", c.stem_kind);
  if (c.stem_kind == 1) {
    const int hw = (c.img_size / 4) * (c.img_size / 4);
    QV_CHECK(sln_ok(hw, c.c2) && sln_ok(hw, c.c3) && sln_ok(hw, c.c4),
             "lateral: the HQAViTv2 stem's LayerNorm([C, %d, %d]) maps must hold 2048 / 4096 / 8192 / 16384 elements", c.img_size / 4,
             c.img_size / 4);
    QV_CHECK((c.in_channels * 16) % 8 == 0, "lateral: in_channels %d", c.in_channels);
    for (int j = 0; j < V2_BLOCKS; ++j)
      QV_CHECK(c.stem_drop_path[j] >= 0.f && c.stem_drop_path[j] < 1.f, "lateral: stem_drop_path[%d] = %f", j, c.stem_drop_path[j]);
  }
  return 0;
}
bool v2_droppath(const qavit_lateral_cfg& c) {
  if (c.stem_kind != 1 || !c.train) return false;
  for (int j = 0; j < V2_BLOCKS; ++j)
    if (c.stem_drop_path[j] > 0.f) return true;
  return false;
}

// ------------------------------------------------------------------------------------------------ plan
struct LinPlan { size_t wb = 0, wbt = 0; int N = 0, K = 0; };
struct CnxPlan { int C; size_t u, stats, n, hpre, h, out; LinPlan w1, w2; };
struct CbPlan { int Cin, Cout; size_t c, mr, a, wp; LinPlan w; int Kp; };
struct LmPlan { int C; size_t cat, pm, p, stats, A, A32; LinPlan w; };   // pm: projection at map resolution (only when it is resized)
struct RrPlan { size_t r0, r2, stats; CnxPlan blk[4]; size_t w2s[4], b2s[4]; LinPlan rev, re; };
struct Cx2Plan { CnxPlan x; size_t w2s, b2s; };                                  // + the LayerScale-folded fp32 pwconv2 copies
struct DsPlan { int Cin, Cout; size_t stats, n, out; LinPlan w; };
struct StemV2Plan {
  int K0;                                                                          // in_channels * 16
  size_t c0, st0, a0, rs, rng;
  LinPlan w0;
  Cx2Plan blk[V2_BLOCKS];
  DsPlan ds[2];
  size_t col, G, gb;                                                               // scratch
};
struct Plan {
  int dt, B, d, nb;
  int kind;
  StemV2Plan v2;
  size_t ts;
  long R1, R, Rt;        // rows at the stem resolution (img/2)^2, the feature-map resolution (img/4)^2 and the token grid
  int H1, H, Ht;         // side lengths; Ht != H: LMFAdapter resizes its projection bilinearly (H:839-843)
  bool resize;
  CbPlan cb[4];
  CnxPlan cx[3];
  LmPlan lm[3];
  RrPlan rr[3];
  size_t saved_total;
  // scratch
  size_t col0, col1, sums, wide, t1, t2, df[3], dA, da0, dc0, dwp0, dwp1, scratch_total;
};

void plan_lin(Bump& b, LinPlan& l, int N, int K, bool bf) {
  l.N = N; l.K = K;
  l.wb = b.take(bf ? (size_t)N * K * 2 : 0);
  l.wbt = b.take(bf ? (size_t)N * K * 2 : 0);
}
void plan_cnx(Bump& b, CnxPlan& x, int C, long R, size_t ts, bool bf) {
  x.C = C;
  x.u = b.take(R * C * ts);
  x.stats = b.take(R * 8);
  x.n = b.take(R * C * ts);
  x.hpre = b.take(R * 4 * C * ts);
  x.h = b.take(R * 4 * C * ts);
  x.out = b.take(R * C * ts);
  plan_lin(b, x.w1, 4 * C, C, bf);
  plan_lin(b, x.w2, C, 4 * C, bf);
}

void make_plan(const qavit_lateral_cfg& c, Plan* P) {
  Plan& p = *P;
  p.dt = c.dtype; p.B = c.batch; p.d = c.dim; p.nb = c.rrcv_blocks;
  p.ts = c.dtype == QV_BF16 ? 2 : 4;
  p.H1 = c.img_size / 2; p.H = c.img_size / 4;
  p.R1 = (long)c.batch * p.H1 * p.H1;
  p.R = (long)c.batch * p.H * p.H;
  p.Ht = c.grid;
  p.Rt = (long)c.batch * p.Ht * p.Ht;
  p.resize = p.Ht != p.H;
  const bool bf = c.dtype == QV_BF16;
  const size_t ts = p.ts;
  const int chans[5] = {c.in_channels, c.c_stem, c.c2, c.c3, c.c4};
  Bump b;
  p.kind = c.stem_kind;
  if (p.kind == 1) {
    StemV2Plan& v = p.v2;
    v.K0 = c.in_channels * 16;
    v.c0 = b.take(p.R * c.c2 * ts);
    v.st0 = b.take((size_t)c.batch * 8);
    v.a0 = b.take(p.R * c.c2 * ts);
    v.rs = b.take((size_t)V2_BLOCKS * c.batch * 4);
    v.rng = b.take(16);
    plan_lin(b, v.w0, c.c2, v.K0, bf);
    for (int j = 0; j < V2_BLOCKS; ++j) {
      const int C = chans[2 + kV2Stage[j]];
      plan_cnx(b, v.blk[j].x, C, p.R, ts, bf);
      v.blk[j].w2s = b.take((size_t)C * 4 * C * 4);
      v.blk[j].b2s = b.take((size_t)C * 4);
    }
    for (int i = 0; i < 2; ++i) {
      DsPlan& q = v.ds[i];
      q.Cin = chans[2 + i]; q.Cout = chans[3 + i];
      q.stats = b.take((size_t)c.batch * 8);
      q.n = b.take(p.R * q.Cin * ts);
      q.out = b.take(p.R * q.Cout * ts);
      plan_lin(b, q.w, q.Cout, q.Cin, bf);
    }
  }
  for (int i = 0; i < 4 && p.kind == 0; ++i) {
    CbPlan& q = p.cb[i];
    q.Cin = chans[i]; q.Cout = chans[i + 1];
    const long rows = i == 0 ? p.R1 : p.R;
    q.Kp = i < 2 ? ((9 * q.Cin + 31) / 32) * 32 : q.Cin;
    q.c = b.take(rows * q.Cout * ts);
    q.mr = b.take(2 * q.Cout * 4);
    q.a = b.take(rows * q.Cout * ts);
    q.wp = b.take(i < 2 ? (size_t)q.Cout * q.Kp * 4 : 0);
    plan_lin(b, q.w, q.Cout, q.Kp, bf);
    if (i >= 1) plan_cnx(b, p.cx[i - 1], q.Cout, p.R, ts, bf);
  }
  for (int i = 0; i < 3; ++i) {
    LmPlan& q = p.lm[i];
    q.C = chans[i + 2];
    q.cat = b.take(p.R * 3 * q.C * ts);
    q.pm = b.take(p.resize ? p.R * c.dim * ts : 0);
    q.p = b.take(p.Rt * c.dim * ts);
    q.stats = b.take(p.Rt * 8);
    q.A = b.take(p.Rt * c.dim * ts);
    q.A32 = bf ? b.take(p.Rt * c.dim * 4) : q.A;   // fp32 copy of A for RRCV's residual (autocast keeps LN outputs in fp32)
    plan_lin(b, q.w, c.dim, 3 * q.C, bf);
  }
  for (int i = 0; i < 3; ++i) {
    RrPlan& q = p.rr[i];
    q.r0 = b.take(p.Rt * c.rrcv_channels * ts);
    for (int j = 0; j < c.rrcv_blocks; ++j) {
      plan_cnx(b, q.blk[j], c.rrcv_channels, p.Rt, ts, bf);
      q.w2s[j] = b.take(c.stem_kind == 1 ? (size_t)c.rrcv_channels * 4 * c.rrcv_channels * 4 : 0);
      q.b2s[j] = b.take(c.stem_kind == 1 ? (size_t)c.rrcv_channels * 4 : 0);
    }
    q.r2 = b.take(p.Rt * c.dim * ts);
    q.stats = b.take(p.Rt * 8);
    plan_lin(b, q.rev, c.rrcv_channels, c.dim, bf);
    plan_lin(b, q.re, c.dim, c.rrcv_channels, bf);
  }
  // gradients of the three feature maps: written by the adapters' backward, consumed by the stem's backward -- they live in `saved`
  // (not scratch) because the phases may be separate calls (qavit_lateral_backward_parts)
  for (int i = 0; i < 3; ++i) p.df[i] = b.take(p.R * chans[i + 2] * ts);
  p.saved_total = b.off;

  Bump s;
  int cmax = c.dim;
  for (int i = 2; i < 5; ++i) cmax = chans[i] > cmax ? chans[i] : cmax;
  cmax = c.rrcv_channels > cmax ? c.rrcv_channels : cmax;
  int wide = 4 * cmax;
  if (p.kind == 0 && p.cb[1].Kp > wide) wide = p.cb[1].Kp;
  if (p.kind == 1) {
    p.v2.col = s.take(p.R * p.v2.K0 * ts);
    p.v2.G = s.take((size_t)cmax * 4 * cmax * 4);
    p.v2.gb = s.take((size_t)cmax * 4);
  }
  p.col0 = s.take(p.kind == 0 ? p.R1 * p.cb[0].Kp * ts : 0);
  p.col1 = s.take(p.kind == 0 ? p.R * p.cb[1].Kp * ts : 0);
  p.sums = s.take(2 * 2048 * 4);
  const long Rm = p.R > p.Rt ? p.R : p.Rt;
  p.wide = s.take(Rm * wide * ts);
  p.t1 = s.take(Rm * cmax * ts);
  p.t2 = s.take(Rm * cmax * ts);
  p.dA = s.take(Rm * c.dim * ts);
  p.da0 = s.take(p.kind == 0 ? p.R1 * c.c_stem * ts : 0);
  p.dc0 = s.take(p.kind == 0 ? p.R1 * c.c_stem * ts : 0);
  p.dwp0 = s.take(p.kind == 0 ? (size_t)p.cb[0].Cout * p.cb[0].Kp * 4 : 0);
  p.dwp1 = s.take(p.kind == 0 ? (size_t)p.cb[1].Cout * p.cb[1].Kp * 4 : 0);
  p.scratch_total = s.off;
}

struct Ctx {
  Plan P;
  const qavit_lateral_cfg* cfg;
  const void* const* params;
  float* const* grads;
  uint8_t* saved;
  uint8_t* scratch;
  cudaStream_t st;
  const float* pf(int i) const { return static_cast<const float*>(params[i]); }
  float* gf(int i) const { return grads ? grads[i] : nullptr; }
  void* sv(size_t off) const { return saved + off; }
  void* sc(size_t off) const { return scratch + off; }
  Weight W(const LinPlan& l, const float* w) const {
    Weight r;
    r.w = w; r.N = l.N; r.K = l.K;
    if (P.dt == QV_BF16) {
      r.wb = reinterpret_cast<const bf16*>(saved + l.wb);
      r.wbt = reinterpret_cast<const bf16*>(saved + l.wbt);
    }
    return r;
  }
};

GemmEpi epi(const Ctx& c, const float* bias, void* C, int ldc) {
  GemmEpi e;
  e.bias = bias; e.C = C; e.ldc = ldc; e.c_f32 = c.P.dt == QV_F32;
  return e;
}
// y = x W^T + b
int lin_fwd(const Ctx& c, const void* x, int ldx, long M, const LinPlan& l, const float* w, const float* bias, void* y) {
  return gemm_nt(c.st, c.P.dt, x, ldx, (int)M, c.W(l, w), epi(c, bias, y, l.N));
}
// dW += dy^T x, db += colsum(dy);  dx = dy W (+ resid) when dx != nullptr
int lin_bwd(const Ctx& c, const void* x, int ldx, const void* dy, long M, const LinPlan& l, const float* w, float* dw, float* db,
            void* dx, const void* resid, bool resid_f32 = false) {
  QV_TRY(gemm_tn(c.st, c.P.dt, dy, l.N, x, ldx, (int)M, l.N, l.K, dw, db, nullptr));
  if (dx) {
    GemmEpi e;
    if (resid) {
      e.resid = resid; e.ldr = l.K; e.r_bf16 = (c.P.dt == QV_BF16 && !resid_f32); e.C2 = dx; e.ldc2 = l.K; e.c2_f32 = c.P.dt == QV_F32;
    } else {
      e = epi(c, nullptr, dx, l.K);
    }
    QV_TRY(gemm_nn(c.st, c.P.dt, dy, l.N, (int)M, c.W(l, w), e));
  }
  return 0;
}

// LayerScale + DropPath of HQAViTv2's stem blocks (V:730-748): out = x + rowscale_b * gamma * pwconv2(...).  gamma is folded into
// the pwconv2 copies w2s / b2s (fp32, in `saved`, made by convert_all); rowscale == nullptr: no DropPath
struct Ls {
  size_t w2s, b2s;
  const float* rowscale;
  void* tmp;        // [R, C] scratch (fp32 runs with DropPath: the SIMT GEMM has no row-scale epilogue)
  float* G;         // [C, 4C] + [C] fp32 scratch for the gradients wrt the folded copies
  float* gb;
};

// ---- ConvNeXtBlock (H:718-739): out = x + pwconv2(gelu(pwconv1(LN(dwconv7(x)))))
int cnx_fwd(const Ctx& c, const CnxPlan& x, int pbase, int H, const void* in, void* out, const Ls* ls = nullptr) {
  const Plan& P = c.P;
  const int C = x.C, dt = P.dt;
  const long R = (long)P.B * H * H;
  DwP d{};
  d.x = in; d.ldx = C; d.B = P.B; d.H = H; d.W = H; d.C = C; d.K = 7; d.w = c.pf(pbase + CX_DW_W); d.bias = c.pf(pbase + CX_DW_B);
  d.y = c.sv(x.u); d.ldy = C;
  QV_TRY(dw2d_fwd(c.st, dt, d, false));
  QV_TRY(ln_fwd(c.st, dt, c.sv(x.u), C, (int)R, C, c.pf(pbase + CX_LN_W), c.pf(pbase + CX_LN_B), 1e-6f, 0, nullptr, nullptr, dt,
                c.sv(x.n), C, static_cast<float*>(c.sv(x.stats))));
  GemmEpi e1 = epi(c, c.pf(pbase + CX_B1), c.sv(x.hpre), 4 * C);
  e1.gelu = 1; e1.C2 = c.sv(x.h); e1.ldc2 = 4 * C; e1.c2_f32 = dt == QV_F32;
  e1.gelu_dgrad = dt == QV_BF16;   // bf16 runs keep gelu'(pre) in x.hpre: backward's epilogue is then one multiply (was 11 instructions)
  QV_TRY(gemm_nt(c.st, dt, c.sv(x.n), C, (int)R, c.W(x.w1, c.pf(pbase + CX_W1)), e1));
  GemmEpi e2;
  e2.bias = c.pf(pbase + CX_B2); e2.resid = in; e2.ldr = C; e2.r_bf16 = dt == QV_BF16; e2.C2 = out; e2.ldc2 = C; e2.c2_f32 = dt == QV_F32;
  const float* w2 = c.pf(pbase + CX_W2);
  if (ls) {
    w2 = static_cast<const float*>(c.sv(ls->w2s));
    e2.bias = static_cast<const float*>(c.sv(ls->b2s));
    if (ls->rowscale && dt == QV_BF16) { e2.rowscale = ls->rowscale; e2.rows_per_img = H * H; }
    if (ls->rowscale && dt == QV_F32) {
      GemmEpi e = epi(c, e2.bias, ls->tmp, C);
      QV_TRY(gemm_nt(c.st, dt, c.sv(x.h), 4 * C, (int)R, c.W(x.w2, w2), e));
      return scale_rows(c.st, dt, ls->tmp, R, C, ls->rowscale, H * H, in, out);
    }
  }
  QV_TRY(gemm_nt(c.st, dt, c.sv(x.h), 4 * C, (int)R, c.W(x.w2, w2), e2));
  return 0;
}
// dx may alias dout; wide: [R, 4C] scratch, tmp: [R, C] scratch
int cnx_bwd(const Ctx& c, const CnxPlan& x, int pbase, int H, const void* in, const void* dout, void* dx, void* wide, void* tmp,
            const Ls* ls = nullptr) {
  const Plan& P = c.P;
  const int C = x.C, dt = P.dt;
  const long R = (long)P.B * H * H;
  const void* dy = dout;                       // gradient of the (scaled) pwconv2 output
  const float* w2 = c.pf(pbase + CX_W2);
  float *dw2 = c.gf(pbase + CX_W2), *db2 = c.gf(pbase + CX_B2);
  if (ls) {
    if (ls->rowscale) {                        // DropPath: dy = rowscale_b * dout (tmp is free until the pwconv1 dX GEMM)
      QV_TRY(scale_rows(c.st, dt, dout, R, C, ls->rowscale, H * H, nullptr, tmp));
      dy = tmp;
    }
    w2 = static_cast<const float*>(c.sv(ls->w2s));
    dw2 = ls->G; db2 = ls->gb;
    QV_CUDA(cudaMemsetAsync(ls->G, 0, ((size_t)C * 4 * C + C) * 4, c.st));   // gb follows G (planned back to back)
  }
  // pwconv2: dW2 += dy^T h;  dhpre = (dy W2) * gelu'(hpre)
  QV_TRY(gemm_tn(c.st, dt, dy, C, c.sv(x.h), 4 * C, (int)R, C, 4 * C, dw2, db2, nullptr));
  if (ls)
    QV_TRY(layerscale_finish(c.st, ls->G, ls->gb, c.pf(pbase + CX_W2), c.pf(pbase + CX_B2), c.pf(pbase + CX2_GAMMA), C, 4 * C,
                             c.gf(pbase + CX_W2), c.gf(pbase + CX_B2), c.gf(pbase + CX2_GAMMA)));
  GemmEpi e = epi(c, nullptr, wide, 4 * C);
  e.gmul = c.sv(x.hpre); e.ldg = 4 * C; e.g_bf16 = dt == QV_BF16; e.gmul_raw = dt == QV_BF16;
  QV_TRY(gemm_nn(c.st, dt, dy, C, (int)R, c.W(x.w2, w2), e));
  // pwconv1
  QV_TRY(gemm_tn(c.st, dt, wide, 4 * C, c.sv(x.n), C, (int)R, 4 * C, C, c.gf(pbase + CX_W1), c.gf(pbase + CX_B1), nullptr));
  QV_TRY(gemm_nn(c.st, dt, wide, 4 * C, (int)R, c.W(x.w1, c.pf(pbase + CX_W1)), epi(c, nullptr, tmp, C)));
  // LayerNorm (in place on tmp)
  QV_TRY(ln_bwd(c.st, dt, c.sv(x.u), C, dt, tmp, C, (int)R, C, c.pf(pbase + CX_LN_W), static_cast<const float*>(c.sv(x.stats)), 0, dt,
                dt == QV_BF16 ? tmp : nullptr, dt == QV_F32 ? static_cast<float*>(tmp) : nullptr, nullptr, c.gf(pbase + CX_LN_W),
                c.gf(pbase + CX_LN_B)));
  // depthwise 7x7: weight gradient, then dx = dgrad(du) + dout
  QV_TRY(dw2d_wgrad(c.st, dt, 7, in, C, tmp, C, P.B, H, H, C, c.gf(pbase + CX_DW_W), c.gf(pbase + CX_DW_B)));
  if (dx) {
    DwP d{};
    d.x = tmp; d.ldx = C; d.B = P.B; d.H = H; d.W = H; d.C = C; d.K = 7; d.w = c.pf(pbase + CX_DW_W);
    d.y = dx; d.ldy = C; d.resid = dout; d.ldr = C;
    QV_TRY(dw2d_fwd(c.st, dt, d, true));
  }
  return 0;
}

int convert_all(const Ctx& c, bool train) {
  (void)train;
  const Plan& P = c.P;
  const qavit_lateral_cfg& cfg = *c.cfg;
  if (P.kind == 1) {   // LayerScale folded into fp32 copies of pwconv2 (the source of the bf16 copies below)
    for (int j = 0; j < V2_BLOCKS; ++j) {
      const int pb = v2_cx_base(j), C = P.v2.blk[j].x.C;
      QV_TRY(layerscale_prepare(c.st, c.pf(pb + CX_W2), c.pf(pb + CX_B2), c.pf(pb + CX2_GAMMA), C, 4 * C,
                                static_cast<float*>(c.sv(P.v2.blk[j].w2s)), static_cast<float*>(c.sv(P.v2.blk[j].b2s))));
    }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < cfg.rrcv_blocks; ++j) {
        const int pb = rr_blk(cfg, i, j), C = cfg.rrcv_channels;
        QV_TRY(layerscale_prepare(c.st, c.pf(pb + CX_W2), c.pf(pb + CX_B2), c.pf(pb + CX2_GAMMA), C, 4 * C,
                                  static_cast<float*>(c.sv(P.rr[i].w2s[j])), static_cast<float*>(c.sv(P.rr[i].b2s[j]))));
      }
  }
  // packed fp32 copies of the two 3x3 stride-2 conv weights (k = (ky, kx, cin) order, zero padded)
  for (int i = 0; i < 2 && P.kind == 0; ++i)
    QV_TRY(conv_w_pack(c.st, c.pf(cb_base(i) + CB_W), P.cb[i].Cout, P.cb[i].Cin, P.cb[i].Kp, static_cast<float*>(c.sv(P.cb[i].wp))));
  if (P.dt != QV_BF16) return 0;
  ConvertJobs jobs{};
  auto add = [&](const LinPlan& l, const float* w) -> int {
    jobs.j[jobs.n++] = ConvertJob{w, l.N, l.K, reinterpret_cast<bf16*>(c.sv(l.wb)), reinterpret_cast<bf16*>(c.sv(l.wbt))};
    if (jobs.n == 24) { QV_TRY(convert_weights_batched(c.st, jobs)); jobs.n = 0; }
    return 0;
  };
  if (P.kind == 1) {
    QV_TRY(add(P.v2.w0, c.pf(S2_CONV_W)));
    for (int j = 0; j < V2_BLOCKS; ++j) {
      QV_TRY(add(P.v2.blk[j].x.w1, c.pf(v2_cx_base(j) + CX_W1)));
      QV_TRY(add(P.v2.blk[j].x.w2, static_cast<const float*>(c.sv(P.v2.blk[j].w2s))));
    }
    for (int i = 0; i < 2; ++i) QV_TRY(add(P.v2.ds[i].w, c.pf(v2_ds_base(i) + DS_W)));
  }
  for (int i = 0; i < 4 && P.kind == 0; ++i)
    QV_TRY(add(P.cb[i].w, i < 2 ? static_cast<const float*>(c.sv(P.cb[i].wp)) : c.pf(cb_base(i) + CB_W)));
  for (int i = 0; i < 3; ++i) {
    if (P.kind == 0) {
      QV_TRY(add(P.cx[i].w1, c.pf(cx_base(i) + CX_W1)));
      QV_TRY(add(P.cx[i].w2, c.pf(cx_base(i) + CX_W2)));
    }
    QV_TRY(add(P.lm[i].w, c.pf(lm_base(cfg, i) + LM_PROJ_W)));
    QV_TRY(add(P.rr[i].rev, c.pf(rr_base(cfg, i) + RR_REV_W)));
    QV_TRY(add(P.rr[i].re, c.pf(rr_base(cfg, i) + RR_RE_W)));
    for (int j = 0; j < cfg.rrcv_blocks; ++j) {
      QV_TRY(add(P.rr[i].blk[j].w1, c.pf(rr_blk(cfg, i, j) + CX_W1)));
      QV_TRY(add(P.rr[i].blk[j].w2, P.kind == 1 ? static_cast<const float*>(c.sv(P.rr[i].w2s[j])) : c.pf(rr_blk(cfg, i, j) + CX_W2)));
    }
  }
  if (jobs.n) QV_TRY(convert_weights_batched(c.st, jobs));
  return 0;
}

int init_ctx(Ctx* c, const qavit_lateral_cfg* cfg, const void* const* params, float* const* grads, const void* saved, void* scratch,
             void* stream) {
  QV_CHECK(cfg && params && saved && scratch, "lateral: null argument");
  QV_TRY(check_cfg(*cfg));
  make_plan(*cfg, &c->P);
  c->cfg = cfg; c->params = params; c->grads = grads;
  c->saved = static_cast<uint8_t*>(const_cast<void*>(saved));
  c->scratch = static_cast<uint8_t*>(scratch);
  c->st = static_cast<cudaStream_t>(stream);
  return 0;
}

// weight used by conv unit i: the packed copy for the 3x3 convs, the parameter itself for the 1x1 convs
const float* cb_weight(const Ctx& c, int i) {
  return i < 2 ? static_cast<const float*>(c.sv(c.P.cb[i].wp)) : c.pf(cb_base(i) + CB_W);
}

}  // namespace

extern "C" int qavit_lateral_param_count(const qavit_lateral_cfg* cfg) {
  if (!cfg || check_cfg(*cfg)) return -1;
  return param_count(*cfg);
}

extern "C" const char* qavit_lateral_param_name(const qavit_lateral_cfg* cfg, int index) {
  static thread_local char buf[128];
  if (!cfg || index < 0 || index >= param_count(*cfg)) return nullptr;
  static const char* kCb[4][2] = {{"cnn_stem.stem.0", "cnn_stem.stem.1"}, {"cnn_stem.stage1.0", "cnn_stem.stage1.1"},
                                  {"cnn_stem.stage2.0", "cnn_stem.stage2.1"}, {"cnn_stem.stage3.0", "cnn_stem.stage3.1"}};
  static const char* kCx[3] = {"cnn_stem.stage1.3", "cnn_stem.stage2.2", "cnn_stem.stage3.2"};
  static const char* kV2Blk[V2_BLOCKS] = {"cnn_stem.stage2.0", "cnn_stem.stage2.1", "cnn_stem.stage3.0", "cnn_stem.stage3.1",
                                          "cnn_stem.stage3.2", "cnn_stem.stage4.0", "cnn_stem.stage4.1"};
  if (cfg->stem_kind == 1 && index < stem_count(*cfg)) {
    static const char* kS2[S2_N] = {"cnn_stem.stem.0.weight", "cnn_stem.stem.0.bias", "cnn_stem.stem.1.weight", "cnn_stem.stem.1.bias"};
    static const char* kDs[DS_N] = {"0.weight", "0.bias", "1.weight", "1.bias"};
    if (index < S2_N) {
      snprintf(buf, sizeof(buf), "%s", kS2[index]);
    } else if (index < v2_ds_base(0)) {
      const int j = (index - S2_N) / CX2_N, k = (index - S2_N) % CX2_N;
      snprintf(buf, sizeof(buf), "%s.%s", kV2Blk[j], k == CX2_GAMMA ? "gamma" : kCxSfx[k]);
    } else {
      const int i = (index - v2_ds_base(0)) / DS_N, k = (index - v2_ds_base(0)) % DS_N;
      snprintf(buf, sizeof(buf), "cnn_stem.downsample%d.%s", i + 2, kDs[k]);
    }
    return buf;
  }
  if (cfg->stem_kind == 1) {
    if (index < rr_base(*cfg, 0)) {
      const int i = (index - lm_base(*cfg, 0)) / LM_N, k = (index - lm_base(*cfg, 0)) % LM_N;
      snprintf(buf, sizeof(buf), "lmfa%d.%s", i + 2, kLmSfx[k]);
    } else {
      const int rel = index - rr_base(*cfg, 0), i = rel / rr_stride(*cfg), k = rel % rr_stride(*cfg);
      if (k < RR_N) snprintf(buf, sizeof(buf), "rrcv%d.%s", i + 2, kRrSfx[k]);
      else snprintf(buf, sizeof(buf), "rrcv%d.blocks.%d.%s", i + 2, (k - RR_N) / CX2_N,
                    (k - RR_N) % CX2_N == CX2_GAMMA ? "gamma" : kCxSfx[(k - RR_N) % CX2_N]);
    }
    return buf;
  }
  if (index < cx_base(0)) {
    const int i = index / CB_N, k = index % CB_N;
    snprintf(buf, sizeof(buf), "%s.%s", kCb[i][k < BN_W ? 0 : 1], kCbSfx[k]);
  } else if (index < lm_base(*cfg, 0)) {
    const int i = (index - cx_base(0)) / CX_N, k = (index - cx_base(0)) % CX_N;
    snprintf(buf, sizeof(buf), "%s.%s", kCx[i], kCxSfx[k]);
  } else if (index < rr_base(*cfg, 0)) {
    const int i = (index - lm_base(*cfg, 0)) / LM_N, k = (index - lm_base(*cfg, 0)) % LM_N;
    snprintf(buf, sizeof(buf), "lmfa%d.%s", i + 2, kLmSfx[k]);
  } else {
    const int rel = index - rr_base(*cfg, 0), i = rel / rr_stride(*cfg), k = rel % rr_stride(*cfg);
    if (k < RR_N) snprintf(buf, sizeof(buf), "rrcv%d.%s", i + 2, kRrSfx[k]);
    else snprintf(buf, sizeof(buf), "rrcv%d.blocks.%d.%s", i + 2, (k - RR_N) / CX_N, kCxSfx[(k - RR_N) % CX_N]);
  }
  return buf;
}

extern "C" int qavit_lateral_workspace(const qavit_lateral_cfg* cfg, size_t* saved_bytes, size_t* scratch_bytes) {
  QV_CHECK(cfg, "null cfg");
  QV_TRY(check_cfg(*cfg));
  Plan P;
  make_plan(*cfg, &P);
  if (saved_bytes) *saved_bytes = P.saved_total + 256;
  if (scratch_bytes) *scratch_bytes = P.scratch_total + 256;
  return 0;
}

extern "C" int qavit_lateral_forward(const qavit_lateral_cfg* cfg, const void* const* params, const float* img, float* R2, float* R3,
                                     float* R4, void* saved, void* scratch, void* stream) {
  return qavit_lateral_forward_parts(cfg, params, img, R2, R3, R4, saved, scratch, stream, QAVIT_LATERAL_ALL);
}

extern "C" int qavit_lateral_forward_parts(const qavit_lateral_cfg* cfg, const void* const* params, const float* img, float* R2, float* R3,
                                           float* R4, void* saved, void* scratch, void* stream, unsigned parts) {
  QV_RANGE("qavit_lateral_forward");
  Ctx c;
  QV_TRY(init_ctx(&c, cfg, params, nullptr, saved, scratch, stream));
  const Plan& P = c.P;
  const int dt = P.dt, d = P.d;
  const bool train = cfg->train != 0;
  cudaStream_t st = c.st;
  const bool do_stem = (parts & QAVIT_LATERAL_STEM) != 0;
  if (do_stem) QV_TRY(convert_all(c, train));      // every GEMM weight of the path, adapters included
  float* sums = static_cast<float*>(c.sc(P.sums));

  const void* feat[3] = {c.sv(P.cx[0].out), c.sv(P.cx[1].out), c.sv(P.cx[2].out)};
  if (P.kind == 1) { feat[0] = c.sv(P.v2.blk[1].x.out); feat[1] = c.sv(P.v2.blk[4].x.out); feat[2] = c.sv(P.v2.blk[6].x.out); }
  const void* prev = nullptr;
  if (P.kind == 1 && do_stem) {
    // ---- HQAViTv2 stem (V:811-827): patchify conv -> LN([c2, g, g]) -> 2 blocks -> [LN + 1x1 conv -> 3 blocks] -> [... -> 2 blocks]
    const StemV2Plan& v = P.v2;
    const int HW = P.H * P.H;
    const bool dp = v2_droppath(*cfg);
    float* rs = static_cast<float*>(c.sv(v.rs));
    if (dp) {
      QV_CHECK(cfg->rng, "lateral: stem DropPath needs the rng state");
      QV_TRY(rng_snapshot_advance(st, cfg->rng, static_cast<unsigned long long*>(c.sv(v.rng))));
      QV_TRY(stem_droppath_scales(st, static_cast<const unsigned long long*>(c.sv(v.rng)), 0x5C00u, V2_BLOCKS, P.B, cfg->stem_drop_path, rs));
    }
    QV_TRY(patch_rows(st, dt, img, P.B, cfg->in_channels, cfg->img_size, 4, c.sc(v.col)));
    QV_TRY(lin_fwd(c, c.sc(v.col), v.K0, P.R, v.w0, c.pf(S2_CONV_W), c.pf(S2_CONV_B), c.sv(v.c0)));
    QV_TRY(sln_fwd(st, dt, c.sv(v.c0), P.B, HW, cfg->c2, c.pf(S2_LN_W), c.pf(S2_LN_B), 1e-6f, c.sv(v.a0), static_cast<float*>(c.sv(v.st0))));
    prev = c.sv(v.a0);
    for (int j = 0; j < V2_BLOCKS; ++j) {
      if (j > 0 && kV2Stage[j] != kV2Stage[j - 1]) {           // downsample between stages
        const DsPlan& q = v.ds[kV2Stage[j] - 1];
        const int pb = v2_ds_base(kV2Stage[j] - 1);
        QV_TRY(sln_fwd(st, dt, prev, P.B, HW, q.Cin, c.pf(pb + DS_LN_W), c.pf(pb + DS_LN_B), 1e-6f, c.sv(q.n), static_cast<float*>(c.sv(q.stats))));
        QV_TRY(lin_fwd(c, c.sv(q.n), q.Cin, P.R, q.w, c.pf(pb + DS_W), c.pf(pb + DS_B), c.sv(q.out)));
        prev = c.sv(q.out);
      }
      Ls ls{v.blk[j].w2s, v.blk[j].b2s, (dp && cfg->stem_drop_path[j] > 0.f) ? rs + (size_t)j * P.B : nullptr, c.sc(P.t1), nullptr, nullptr};
      QV_TRY(cnx_fwd(c, v.blk[j].x, v2_cx_base(j), P.H, prev, c.sv(v.blk[j].x.out), &ls));
      prev = c.sv(v.blk[j].x.out);
    }
  }
  // ---- CNNStemModel: (conv -> BN [-> GELU]) x 4 with a ConvNeXt block after units 1..3
  for (int i = 0; i < 4 && P.kind == 0 && do_stem; ++i) {
    const CbPlan& q = P.cb[i];
    const int pb = cb_base(i);
    const long rows = i == 0 ? P.R1 : P.R;
    const void* gin;
    int ldin;
    if (i == 0) {
      QV_TRY(im2col_img(st, dt, img, P.B, q.Cin, cfg->img_size, q.Kp, c.sc(P.col0)));
      gin = c.sc(P.col0); ldin = q.Kp;
    } else if (i == 1) {
      QV_TRY(im2col_nhwc(st, dt, prev, P.B, P.H1, q.Cin, c.sc(P.col1)));
      gin = c.sc(P.col1); ldin = q.Kp;
    } else {
      gin = prev; ldin = q.Cin;
    }
    QV_TRY(lin_fwd(c, gin, ldin, rows, q.w, cb_weight(c, i), c.pf(pb + CB_B), c.sv(q.c)));
    QV_TRY(bn_fwd(st, dt, c.sv(q.c), rows, q.Cout, c.pf(pb + BN_W), c.pf(pb + BN_B), cfg->bn_eps, cfg->bn_momentum, train,
                  const_cast<float*>(c.pf(pb + BN_RM)), const_cast<float*>(c.pf(pb + BN_RV)),
                  static_cast<long long*>(const_cast<void*>(params[pb + BN_NBT])), i < 2, sums, static_cast<float*>(c.sv(q.mr)),
                  c.sv(q.a)));
    prev = c.sv(q.a);
    if (i >= 1) {
      QV_TRY(cnx_fwd(c, P.cx[i - 1], cx_base(i - 1), P.H, prev, c.sv(P.cx[i - 1].out)));
      prev = c.sv(P.cx[i - 1].out);
    }
  }
  // ---- LMFAdapter + RRCV per fusion stage
  float* Rout[3] = {R2, R3, R4};
  for (int i = 0; i < 3; ++i) {
    if (!(parts & (QAVIT_LATERAL_ADAPTER2 << i))) continue;
    QV_CHECK(Rout[i], "lateral_forward: null output for adapter %d", i + 2);
    const LmPlan& q = P.lm[i];
    const int pb = lm_base(*cfg, i), C = q.C;
    DwP a{};
    a.x = feat[i]; a.ldx = C; a.B = P.B; a.H = P.H; a.W = P.H; a.C = C; a.K = 3; a.w = c.pf(pb + LM_DW3_W); a.bias = c.pf(pb + LM_DW3_B);
    a.y = c.sv(q.cat); a.ldy = 3 * C;
    QV_TRY(dw2d_fwd(st, dt, a, false));
    a.K = 5; a.w = c.pf(pb + LM_DW5_W); a.bias = c.pf(pb + LM_DW5_B);
    a.y = static_cast<uint8_t*>(c.sv(q.cat)) + (size_t)C * P.ts;
    a.copy = static_cast<uint8_t*>(c.sv(q.cat)) + (size_t)2 * C * P.ts; a.ldcp = 3 * C;
    QV_TRY(dw2d_fwd(st, dt, a, false));
    QV_TRY(lin_fwd(c, c.sv(q.cat), 3 * C, P.R, q.w, c.pf(pb + LM_PROJ_W), c.pf(pb + LM_PROJ_B), c.sv(P.resize ? q.pm : q.p)));
    if (P.resize) QV_TRY(resize_bilinear_fwd(st, dt, c.sv(q.pm), P.B, P.H, P.H, P.Ht, P.Ht, d, c.sv(q.p)));
    QV_TRY(rowln_fwd(st, dt, c.sv(q.p), P.Rt, d, c.pf(pb + LM_LN_W), c.pf(pb + LM_LN_B), 1e-5f, 1, dt, nullptr, nullptr, c.sv(q.A),
                     dt == QV_BF16 ? static_cast<float*>(c.sv(q.A32)) : nullptr, static_cast<float*>(c.sv(q.stats))));
    // RRCV
    const RrPlan& r = P.rr[i];
    const int rb = rr_base(*cfg, i);
    QV_TRY(lin_fwd(c, c.sv(q.A), d, P.Rt, r.rev, c.pf(rb + RR_REV_W), c.pf(rb + RR_REV_B), c.sv(r.r0)));
    const void* cur = c.sv(r.r0);
    for (int j = 0; j < cfg->rrcv_blocks; ++j) {
      Ls ls{r.w2s[j], r.b2s[j], nullptr, nullptr, nullptr, nullptr};
      QV_TRY(cnx_fwd(c, r.blk[j], rr_blk(*cfg, i, j), P.Ht, cur, c.sv(r.blk[j].out), P.kind == 1 ? &ls : nullptr));
      cur = c.sv(r.blk[j].out);
    }
    QV_TRY(lin_fwd(c, cur, cfg->rrcv_channels, P.Rt, r.re, c.pf(rb + RR_RE_W), c.pf(rb + RR_RE_B), c.sv(r.r2)));
    QV_TRY(rowln_fwd(st, dt, c.sv(r.r2), P.Rt, d, c.pf(rb + RR_LN_W), c.pf(rb + RR_LN_B), 1e-5f, 0, QV_F32, c.sv(q.A32),
                     c.pf(rb + RR_BETA), Rout[i], nullptr, static_cast<float*>(c.sv(r.stats))));
  }
  return 0;
}

extern "C" int qavit_lateral_backward(const qavit_lateral_cfg* cfg, const void* const* params, float* const* grads, const float* img,
                                      const float* dR2, const float* dR3, const float* dR4, const void* saved, void* scratch,
                                      void* stream) {
  return qavit_lateral_backward_parts(cfg, params, grads, img, dR2, dR3, dR4, saved, scratch, stream, QAVIT_LATERAL_ALL);
}

extern "C" int qavit_lateral_backward_parts(const qavit_lateral_cfg* cfg, const void* const* params, float* const* grads, const float* img,
                                            const float* dR2, const float* dR3, const float* dR4, const void* saved, void* scratch,
                                            void* stream, unsigned parts) {
  QV_RANGE("qavit_lateral_backward");
  Ctx c;
  QV_TRY(init_ctx(&c, cfg, params, grads, saved, scratch, stream));
  QV_CHECK(grads, "lateral_backward: null grads");
  const Plan& P = c.P;
  const int dt = P.dt, d = P.d, rc = cfg->rrcv_channels;
  const bool train = cfg->train != 0;
  cudaStream_t st = c.st;
  float* sums = static_cast<float*>(c.sc(P.sums));
  void* wide = c.sc(P.wide);
  void* t1 = c.sc(P.t1);
  void* t2 = c.sc(P.t2);
  const float* dRs[3] = {dR2, dR3, dR4};
  const void* feat[3] = {c.sv(P.cx[0].out), c.sv(P.cx[1].out), c.sv(P.cx[2].out)};
  if (P.kind == 1) { feat[0] = c.sv(P.v2.blk[1].x.out); feat[1] = c.sv(P.v2.blk[4].x.out); feat[2] = c.sv(P.v2.blk[6].x.out); }

  // ---- RRCV + LMFAdapter backward per stage: dR_i -> df_i
  for (int i = 2; i >= 0; --i) {
    if (!(parts & (QAVIT_LATERAL_ADAPTER2 << i))) continue;
    const LmPlan& q = P.lm[i];
    const RrPlan& r = P.rr[i];
    const int rb = rr_base(*cfg, i), pb = lm_base(*cfg, i), C = q.C;
    void* df = c.sv(P.df[i]);
    void* dA = c.sc(P.dA);
    if (dRs[i] == nullptr) {   // this stage's fusion is not part of the graph
      QV_CUDA(cudaMemsetAsync(df, 0, (size_t)P.R * C * P.ts, st));
      continue;
    }
    // R = A + beta * LN(r2)
    QV_TRY(rowln_bwd(st, dt, c.sv(r.r2), QV_F32, dRs[i], P.Rt, d, c.pf(rb + RR_LN_W), c.pf(rb + RR_LN_B), static_cast<const float*>(c.sv(r.stats)),
                     0, c.pf(rb + RR_BETA), c.gf(rb + RR_BETA), t1, c.gf(rb + RR_LN_W), c.gf(rb + RR_LN_B)));
    const void* last = c.sv(r.blk[cfg->rrcv_blocks - 1].out);
    QV_TRY(lin_bwd(c, last, rc, t1, P.Rt, r.re, c.pf(rb + RR_RE_W), c.gf(rb + RR_RE_W), c.gf(rb + RR_RE_B), t2, nullptr));
    for (int j = cfg->rrcv_blocks - 1; j >= 0; --j) {
      const void* in = j == 0 ? c.sv(r.r0) : c.sv(r.blk[j - 1].out);
      float* G = P.kind == 1 ? static_cast<float*>(c.sc(P.v2.G)) : nullptr;
      Ls ls{r.w2s[j], r.b2s[j], nullptr, nullptr, G, G ? G + (size_t)rc * 4 * rc : nullptr};
      QV_TRY(cnx_bwd(c, r.blk[j], rr_blk(*cfg, i, j), P.Ht, in, t2, t2, wide, t1, P.kind == 1 ? &ls : nullptr));
    }
    // dA = dR + dr0 W_rev
    QV_TRY(lin_bwd(c, c.sv(q.A), d, t2, P.Rt, r.rev, c.pf(rb + RR_REV_W), c.gf(rb + RR_REV_W), c.gf(rb + RR_REV_B), dA, dRs[i], true));
    // A = gelu(LN(p))
    QV_TRY(rowln_bwd(st, dt, c.sv(q.p), dt, dA, P.Rt, d, c.pf(pb + LM_LN_W), c.pf(pb + LM_LN_B), static_cast<const float*>(c.sv(q.stats)), 1,
                     nullptr, nullptr, t1, c.gf(pb + LM_LN_W), c.gf(pb + LM_LN_B)));
    const void* dproj = t1;
    if (P.resize) {   // t2 is free again: gradient of the map-resolution projection
      QV_TRY(resize_bilinear_bwd(st, dt, t1, P.B, P.H, P.H, P.Ht, P.Ht, d, t2));
      dproj = t2;
    }
    QV_TRY(lin_bwd(c, c.sv(q.cat), 3 * C, dproj, P.R, q.w, c.pf(pb + LM_PROJ_W), c.gf(pb + LM_PROJ_W), c.gf(pb + LM_PROJ_B), wide, nullptr));
    // dcat = [d(dw3) | d(dw5) | d(identity)] -> df
    const uint8_t* g1 = static_cast<const uint8_t*>(wide);
    const uint8_t* g2 = g1 + (size_t)C * P.ts;
    const uint8_t* g3 = g1 + (size_t)2 * C * P.ts;
    QV_TRY(dw2d_wgrad(st, dt, 3, feat[i], C, g1, 3 * C, P.B, P.H, P.H, C, c.gf(pb + LM_DW3_W), c.gf(pb + LM_DW3_B)));
    QV_TRY(dw2d_wgrad(st, dt, 5, feat[i], C, g2, 3 * C, P.B, P.H, P.H, C, c.gf(pb + LM_DW5_W), c.gf(pb + LM_DW5_B)));
    DwP a{};
    a.x = g1; a.ldx = 3 * C; a.B = P.B; a.H = P.H; a.W = P.H; a.C = C; a.K = 3; a.w = c.pf(pb + LM_DW3_W); a.y = df; a.ldy = C;
    a.resid = g3; a.ldr = 3 * C;
    QV_TRY(dw2d_fwd(st, dt, a, true));
    a.x = g2; a.K = 5; a.w = c.pf(pb + LM_DW5_W); a.resid = df; a.ldr = C;
    QV_TRY(dw2d_fwd(st, dt, a, true));
  }

  if (!(parts & QAVIT_LATERAL_STEM)) return 0;
  if (P.kind == 1) {
    // ---- HQAViTv2 stem backward: blocks 6 .. 0; the gradient of a stage's feature map (df[stage], from its LMFAdapter) joins at
    // the stage's last block, the downsample's gradient is added to the previous stage's df by the LayerNorm backward
    const StemV2Plan& v = P.v2;
    const int HW = P.H * P.H;
    const bool dp = v2_droppath(*cfg);
    const float* rs = static_cast<const float*>(c.sv(v.rs));
    float* G = static_cast<float*>(c.sc(v.G));
    for (int j = V2_BLOCKS - 1; j >= 0; --j) {
      const int stg = kV2Stage[j], C = v.blk[j].x.C;
      const bool last_of_stage = j == V2_BLOCKS - 1 || kV2Stage[j + 1] != stg;
      const bool first_of_stage = j == 0 || kV2Stage[j - 1] != stg;
      const void* in = first_of_stage ? (stg == 0 ? c.sv(v.a0) : c.sv(v.ds[stg - 1].out)) : c.sv(v.blk[j - 1].x.out);
      const void* dout = last_of_stage ? c.sv(P.df[stg]) : t2;
      Ls ls{v.blk[j].w2s, v.blk[j].b2s, (dp && cfg->stem_drop_path[j] > 0.f) ? rs + (size_t)j * P.B : nullptr, nullptr, G, G + (size_t)C * 4 * C};
      QV_TRY(cnx_bwd(c, v.blk[j].x, v2_cx_base(j), P.H, in, dout, t2, wide, t1, &ls));
      if (first_of_stage && stg > 0) {        // downsample: 1x1 conv, then LN([Cin, g, g]) of the previous stage's feature map
        const DsPlan& q = v.ds[stg - 1];
        const int pb = v2_ds_base(stg - 1);
        QV_TRY(lin_bwd(c, c.sv(q.n), q.Cin, t2, P.R, q.w, c.pf(pb + DS_W), c.gf(pb + DS_W), c.gf(pb + DS_B), t1, nullptr));
        void* dprev = c.sv(P.df[stg - 1]);
        QV_TRY(sln_bwd(st, dt, feat[stg - 1], t1, P.B, HW, q.Cin, c.pf(pb + DS_LN_W), static_cast<const float*>(c.sv(q.stats)), dprev, dprev,
                       c.gf(pb + DS_LN_W), c.gf(pb + DS_LN_B)));
      }
    }
    // stem: LN([c2, g, g]) backward, then the patchify conv's weight / bias gradient (the image receives none)
    QV_TRY(sln_bwd(st, dt, c.sv(v.c0), t2, P.B, HW, cfg->c2, c.pf(S2_LN_W), static_cast<const float*>(c.sv(v.st0)), nullptr, t1,
                   c.gf(S2_LN_W), c.gf(S2_LN_B)));
    QV_TRY(patch_rows(st, dt, img, P.B, cfg->in_channels, cfg->img_size, 4, c.sc(v.col)));
    QV_TRY(lin_bwd(c, c.sc(v.col), v.K0, t1, P.R, v.w0, c.pf(S2_CONV_W), c.gf(S2_CONV_W), c.gf(S2_CONV_B), nullptr, nullptr));
    return 0;
  }
  // ---- CNN stem backward: units 3, 2, 1 (ConvNeXt -> BN -> conv), then unit 0
  for (int i = 3; i >= 1; --i) {
    const CbPlan& q = P.cb[i];
    const int pb = cb_base(i);
    void* df = c.sv(P.df[i - 1]);
    QV_TRY(cnx_bwd(c, P.cx[i - 1], cx_base(i - 1), P.H, c.sv(q.a), df, t2, wide, t1));
    QV_TRY(bn_bwd(st, dt, c.sv(q.c), t2, P.R, q.Cout, c.pf(pb + BN_W), c.pf(pb + BN_B), static_cast<const float*>(c.sv(q.mr)), train,
                  i < 2, sums, t1, c.gf(pb + BN_W), c.gf(pb + BN_B)));
    if (i >= 2) {   // 1x1 conv: the input is the previous stage's feature map, its gradient accumulates into df[i - 2]
      void* dprev = c.sv(P.df[i - 2]);
      QV_TRY(lin_bwd(c, feat[i - 2], q.Cin, t1, P.R, q.w, c.pf(pb + CB_W), c.gf(pb + CB_W), c.gf(pb + CB_B), dprev, dprev));
    } else {        // 3x3 stride-2 conv on the stem output
      float* dwp = static_cast<float*>(c.sc(P.dwp1));
      QV_CUDA(cudaMemsetAsync(dwp, 0, (size_t)q.Cout * q.Kp * 4, st));
      QV_TRY(im2col_nhwc(st, dt, c.sv(P.cb[0].a), P.B, P.H1, q.Cin, c.sc(P.col1)));
      QV_TRY(lin_bwd(c, c.sc(P.col1), q.Kp, t1, P.R, q.w, cb_weight(c, 1), dwp, c.gf(pb + CB_B), wide, nullptr));
      QV_TRY(conv_w_unpack_add(st, dwp, q.Cout, q.Cin, q.Kp, c.gf(pb + CB_W)));
      QV_TRY(col2im_nhwc(st, dt, wide, P.B, P.H1, q.Cin, c.sc(P.da0)));
    }
  }
  {
    const CbPlan& q = P.cb[0];
    const int pb = cb_base(0);
    QV_TRY(bn_bwd(st, dt, c.sv(q.c), c.sc(P.da0), P.R1, q.Cout, c.pf(pb + BN_W), c.pf(pb + BN_B), static_cast<const float*>(c.sv(q.mr)),
                  train, 1, sums, c.sc(P.dc0), c.gf(pb + BN_W), c.gf(pb + BN_B)));
    float* dwp = static_cast<float*>(c.sc(P.dwp0));
    QV_CUDA(cudaMemsetAsync(dwp, 0, (size_t)q.Cout * q.Kp * 4, st));
    QV_TRY(im2col_img(st, dt, img, P.B, q.Cin, cfg->img_size, q.Kp, c.sc(P.col0)));
    QV_TRY(lin_bwd(c, c.sc(P.col0), q.Kp, c.sc(P.dc0), P.R1, q.w, cb_weight(c, 0), dwp, c.gf(pb + CB_B), nullptr, nullptr));
    QV_TRY(conv_w_unpack_add(st, dwp, q.Cout, q.Cin, q.Kp, c.gf(pb + CB_W)));
  }
  return 0;
}

// =================================================================================================== SplitFusion
namespace {
enum { SF_GN_W, SF_GN_B, SF_GFC_W, SF_GFC_B, SF_CAT_W, SF_CAT_B, SF_CLN_W, SF_CLN_B, SF_FW, SF_FN_W, SF_FN_B, SF_N };
const char* kSfNames[SF_N] = {"gate_norm.weight", "gate_norm.bias", "gate_fc.weight", "gate_fc.bias", "cat_mlp.0.weight",
                              "cat_mlp.0.bias", "cat_mlp.1.weight", "cat_mlp.1.bias", "fusion_weights", "final_norm.weight",
                              "final_norm.bias"};
struct SfPlan {
  size_t g_in, cat, glin, cpre, gstats, cstats, fstats, alpha, rng;
  LinPlan wg, wc;
  size_t saved_total;
  size_t dglin, dcpre, dg_in, dcat, draw, scratch_total;
};
void make_sf_plan(const qavit_splitfusion_cfg& c, SfPlan* p) {
  const size_t ts = c.dtype == QV_BF16 ? 2 : 4;
  const long R = c.rows;
  const int C = c.dim;
  Bump b;
  p->g_in = b.take(R * C * ts);
  p->cat = b.take(R * 2 * C * ts);
  p->glin = b.take(R * C * ts);
  p->cpre = b.take(R * C * ts);
  p->gstats = b.take(R * 8);
  p->cstats = b.take(R * 8);
  p->fstats = b.take(R * 8);
  p->alpha = b.take(16);
  p->rng = b.take(16);
  plan_lin(b, p->wg, C, C, c.dtype == QV_BF16);
  plan_lin(b, p->wc, C, 2 * C, c.dtype == QV_BF16);
  p->saved_total = b.off;
  Bump s;
  p->dglin = s.take(R * C * ts);
  p->dcpre = s.take(R * C * ts);
  p->dg_in = s.take(R * C * ts);
  p->dcat = s.take(R * 2 * C * ts);
  p->draw = s.take(16);
  p->scratch_total = s.off;
}
int check_sf(const qavit_splitfusion_cfg& c) {
  QV_CHECK(c.rows > 0 && c.dim % 8 == 0 && c.dim <= 256, "splitfusion: rows=%lld dim=%d unsupported", (long long)c.rows, c.dim);
  QV_CHECK(c.dtype == QV_F32 || c.dtype == QV_BF16, "splitfusion: dtype %d", c.dtype);
  QV_CHECK(c.drop_p >= 0.f && c.drop_p < 1.f, "splitfusion: dropout p=%f", c.drop_p);
  return 0;
}
}  // namespace

extern "C" const char* qavit_splitfusion_param_name(int index) { return index >= 0 && index < SF_N ? kSfNames[index] : nullptr; }

extern "C" int qavit_splitfusion_workspace(const qavit_splitfusion_cfg* cfg, size_t* saved_bytes, size_t* scratch_bytes) {
  QV_CHECK(cfg, "null cfg");
  QV_TRY(check_sf(*cfg));
  SfPlan p;
  make_sf_plan(*cfg, &p);
  if (saved_bytes) *saved_bytes = p.saved_total + 256;
  if (scratch_bytes) *scratch_bytes = p.scratch_total + 256;
  return 0;
}

extern "C" int qavit_splitfusion_forward(const qavit_splitfusion_cfg* cfg, const void* const* params, unsigned long long* rng,
                                         const float* T_in, const float* R, float* out, void* saved, void* scratch, void* stream) {
  QV_CHECK(cfg && params && T_in && R && out && saved, "splitfusion: null argument");
  QV_TRY(check_sf(*cfg));
  (void)scratch;
  SfPlan p;
  make_sf_plan(*cfg, &p);
  Ctx c{};
  c.P.dt = cfg->dtype;
  c.params = params; c.saved = static_cast<uint8_t*>(saved); c.st = static_cast<cudaStream_t>(stream);
  const int dt = cfg->dtype, C = cfg->dim;
  const long rows = cfg->rows;
  const bool drop = cfg->train && cfg->drop_p > 0.f;
  QV_CHECK(!drop || rng, "splitfusion: train-mode dropout needs the rng state");
  if (dt == QV_BF16) {
    ConvertJobs jobs{};
    jobs.j[jobs.n++] = ConvertJob{c.pf(SF_GFC_W), C, C, reinterpret_cast<bf16*>(c.sv(p.wg.wb)), reinterpret_cast<bf16*>(c.sv(p.wg.wbt))};
    jobs.j[jobs.n++] = ConvertJob{c.pf(SF_CAT_W), C, 2 * C, reinterpret_cast<bf16*>(c.sv(p.wc.wb)), reinterpret_cast<bf16*>(c.sv(p.wc.wbt))};
    QV_TRY(convert_weights_batched(c.st, jobs));
  }
  if (drop) QV_TRY(rng_snapshot_advance(c.st, rng, static_cast<unsigned long long*>(c.sv(p.rng))));
  QV_TRY(fusion_softmax(c.st, c.pf(SF_FW), 2, static_cast<float*>(c.sv(p.alpha))));
  QV_TRY(sf_pre_fwd(c.st, dt, T_in, R, rows, C, c.pf(SF_GN_W), c.pf(SF_GN_B), c.sv(p.g_in), c.sv(p.cat), static_cast<float*>(c.sv(p.gstats))));
  QV_TRY(lin_fwd(c, c.sv(p.g_in), C, rows, p.wg, c.pf(SF_GFC_W), c.pf(SF_GFC_B), c.sv(p.glin)));
  QV_TRY(lin_fwd(c, c.sv(p.cat), 2 * C, rows, p.wc, c.pf(SF_CAT_W), c.pf(SF_CAT_B), c.sv(p.cpre)));
  SfArgs a{T_in, R, c.sv(p.glin), c.sv(p.cpre), c.pf(SF_CLN_W), c.pf(SF_CLN_B), c.pf(SF_FN_W), c.pf(SF_FN_B), c.pf(SF_FW),
           drop ? cfg->drop_p : 0.f, static_cast<const unsigned long long*>(c.sv(p.rng)), 0x5F01u, rows, C};
  QV_TRY(sf_post_fwd(c.st, dt, a, out, static_cast<float*>(c.sv(p.cstats)), static_cast<float*>(c.sv(p.fstats))));
  return 0;
}

extern "C" int qavit_splitfusion_backward(const qavit_splitfusion_cfg* cfg, const void* const* params, float* const* grads,
                                          const float* T_in, const float* R, const float* dout, float* dT, float* dR, const void* saved,
                                          void* scratch, void* stream) {
  QV_CHECK(cfg && params && grads && T_in && R && dout && dT && dR && saved && scratch, "splitfusion: null argument");
  QV_TRY(check_sf(*cfg));
  SfPlan p;
  make_sf_plan(*cfg, &p);
  Ctx c{};
  c.P.dt = cfg->dtype;
  c.params = params; c.grads = grads;
  c.saved = static_cast<uint8_t*>(const_cast<void*>(saved)); c.scratch = static_cast<uint8_t*>(scratch);
  c.st = static_cast<cudaStream_t>(stream);
  const int dt = cfg->dtype, C = cfg->dim;
  const long rows = cfg->rows;
  const bool drop = cfg->train && cfg->drop_p > 0.f;
  float* draw = static_cast<float*>(c.sc(p.draw));
  QV_CUDA(cudaMemsetAsync(draw, 0, 16, c.st));
  SfArgs a{T_in, R, c.sv(p.glin), c.sv(p.cpre), c.pf(SF_CLN_W), c.pf(SF_CLN_B), c.pf(SF_FN_W), c.pf(SF_FN_B), c.pf(SF_FW),
           drop ? cfg->drop_p : 0.f, static_cast<const unsigned long long*>(c.sv(p.rng)), 0x5F01u, rows, C};
  QV_TRY(sf_post_bwd(c.st, dt, a, dout, static_cast<const float*>(c.sv(p.cstats)), static_cast<const float*>(c.sv(p.fstats)), dT, dR,
                     c.sc(p.dglin), c.sc(p.dcpre), c.gf(SF_FN_W), c.gf(SF_FN_B), c.gf(SF_CLN_W), c.gf(SF_CLN_B), draw));
  QV_TRY(fusion_bwd_final(c.st, static_cast<const float*>(c.sv(p.alpha)), draw, 2, c.gf(SF_FW)));
  QV_TRY(lin_bwd(c, c.sv(p.cat), 2 * C, c.sc(p.dcpre), rows, p.wc, c.pf(SF_CAT_W), c.gf(SF_CAT_W), c.gf(SF_CAT_B), c.sc(p.dcat), nullptr));
  QV_TRY(lin_bwd(c, c.sv(p.g_in), C, c.sc(p.dglin), rows, p.wg, c.pf(SF_GFC_W), c.gf(SF_GFC_W), c.gf(SF_GFC_B), c.sc(p.dg_in), nullptr));
  QV_TRY(sf_pre_bwd(c.st, dt, T_in, R, c.sc(p.dg_in), c.sc(p.dcat), rows, C, c.pf(SF_GN_W), static_cast<const float*>(c.sv(p.gstats)), dT,
                    dR, c.gf(SF_GN_W), c.gf(SF_GN_B)));
  return 0;
}
