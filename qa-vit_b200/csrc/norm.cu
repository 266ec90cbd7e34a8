// LayerNorm forward / backward: one warp per row, the row held in registers, two-pass statistics in fp32.
// Every lane owns EPL contiguous channels (8 when C > 128, else 4) so all global traffic is 8 / 16 B vectors
// (C <= 256, C % EPL == 0).  HBM-bound: algorithmic bytes = rows * C * (in + out) (+ 8 B/row of statistics).
//   gelu_in : the LN input is gelu(x) of the stored pre-activation x (CCF-FFN fc1 -> GELU -> LN, H:704-706)
//   chain   : y = LN2(LN1(x))  (branch .norm followed by bank write_norm, H:468 + H:301)
#include "kernels.h"

namespace {

template <typename TI, typename TO, int EPL, bool GELU_IN>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const TI* __restrict__ x, int ldx, int rows, int C,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     float eps, const float* __restrict__ gamma2,
                                                     const float* __restrict__ beta2, TO* __restrict__ y, int ldy,
                                                     float* __restrict__ stats) {
  QV_PDL_ENTRY();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int c0 = lane * EPL;
  const bool act = c0 < C;
  const float invC = 1.f / (float)C;
  float gm[EPL], bt[EPL], gm2[EPL], bt2[EPL];
#pragma unroll
  for (int i = 0; i < EPL; ++i) { gm[i] = bt[i] = gm2[i] = bt2[i] = 0.f; }
  if (act) {
    load_vec<EPL>(gamma + c0, gm);
    load_vec<EPL>(beta + c0, bt);
    if (gamma2) { load_vec<EPL>(gamma2 + c0, gm2); load_vec<EPL>(beta2 + c0, bt2); }
  }
  // two rows per iteration: both rows' loads are issued before either row's reductions (a warp otherwise serialises
  // load latency -> shuffle reductions -> store for each of its ~10 rows)
  auto finish = [&](float* v, long row) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) { if (GELU_IN) v[i] = gelu_f(v[i]); s += v[i]; }
    const float mean = warp_sum(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) { const float d = act ? v[i] - mean : 0.f; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) * invC + eps);
    if (stats && lane == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
#pragma unroll
    for (int i = 0; i < EPL; ++i) v[i] = (v[i] - mean) * rstd * gm[i] + bt[i];
    if (gamma2) {  // chained second LayerNorm (no statistics kept: the bank write path has no backward)
      float s2 = 0.f;
#pragma unroll
      for (int i = 0; i < EPL; ++i) s2 += act ? v[i] : 0.f;
      const float m2 = warp_sum(s2) * invC;
      float q2 = 0.f;
#pragma unroll
      for (int i = 0; i < EPL; ++i) { const float d = act ? v[i] - m2 : 0.f; q2 += d * d; }
      const float r2 = rsqrtf(warp_sum(q2) * invC + eps);
#pragma unroll
      for (int i = 0; i < EPL; ++i) v[i] = (v[i] - m2) * r2 * gm2[i] + bt2[i];
    }
    if (act) store_vec<EPL>(y + row * ldy + c0, v);
  };
  const long stride = (long)gridDim.x * wpb;
  for (long row = (long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += 2 * stride) {
    const long row2 = row + stride;
    float v[EPL], w[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) v[i] = w[i] = 0.f;
    if (act) {
      load_vec<EPL>(x + row * ldx + c0, v);
      if (row2 < rows) load_vec<EPL>(x + row2 * ldx + c0, w);
    }
    finish(v, row);
    if (row2 < rows) finish(w, row2);
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) [+ resid],  g = dy * gamma;  dgamma += dy * xhat, dbeta += dy.
// dx is written as T (dx_t) and / or fp32 (dx_f32).  resid must not alias an output.
#ifndef LNB_M
#define LNB_M 3
#endif
template <typename TX, typename TDY, typename TO, int EPL, bool GELU_IN>
__global__ void __launch_bounds__(256, LNB_M) ln_bwd_kernel(const TX* __restrict__ x, int ldx, const TDY* __restrict__ dy,
                                                     int lddy, int rows, int C, const float* __restrict__ gamma,
                                                     const float* __restrict__ stats, TO* __restrict__ dx_t,
                                                     float* __restrict__ dx_f32, const float* __restrict__ resid,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta, DropP drop,
                                                     const float* __restrict__ rowscale, int rows_per_img) {
  QV_PDL_ENTRY();
  __shared__ __align__(16) float red[2][8][256];
  const bool masked = EPL == 8 && drop.p > 0.f, scaled = EPL == 8 && rowscale != nullptr;
  DropState dst{};
  if (masked) dst = drop_state(drop);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int c0 = lane * EPL;
  const bool act = c0 < C;
  const float invC = 1.f / (float)C;
  float gm[EPL], ag[EPL], ab[EPL];
#pragma unroll
  for (int i = 0; i < EPL; ++i) gm[i] = ag[i] = ab[i] = 0.f;
  if (act) load_vec<EPL>(gamma + c0, gm);
  for (long row = (long)blockIdx.x * wpb + warp; row < rows; row += (long)gridDim.x * wpb) {
    float xv[EPL], dv[EPL], rs[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) xv[i] = dv[i] = rs[i] = 0.f;
    if (act) {
      load_vec<EPL>(x + row * ldx + c0, xv);
      load_vec<EPL>(dy + row * lddy + c0, dv);
      if (resid) load_vec<EPL>(resid + row * C + c0, rs);
    }
    const float mean = stats[2 * row], rstd = stats[2 * row + 1];
    float s1 = 0.f, s2 = 0.f, u[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      u[i] = xv[i];
      if (GELU_IN) xv[i] = gelu_f(xv[i]);
      xv[i] = act ? (xv[i] - mean) * rstd : 0.f;        // xhat
      ag[i] += dv[i] * xv[i];
      ab[i] += dv[i];
      dv[i] *= gm[i];                                   // g
      s1 += dv[i];
      s2 += dv[i] * xv[i];
    }
    const float m1 = warp_sum(s1) * invC, m2 = warp_sum(s2) * invC;
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      float d = rstd * (dv[i] - m1 - xv[i] * m2);
      if (GELU_IN) d *= gelu_grad_f(u[i]);
      dv[i] = d + rs[i];
    }
    if (act) {
      if (dx_f32) store_vec<EPL>(dx_f32 + row * C + c0, dv);
      if (dx_t) {
        if (EPL == 8 && (masked || scaled)) {   // gradient w.r.t. the pre-dropout activation (dx_f32 stays the stream gradient)
          const float rsc = scaled ? rowscale[row / rows_per_img] : 1.f;
          float k[8];
          if (masked) drop_keep8(dst, (unsigned long long)(row * C + c0) >> 3, k);
#pragma unroll
          for (int i = 0; i < EPL; ++i) dv[i] *= (masked ? k[i & 7] : 1.f) * rsc;
        }
        store_vec<EPL>(dx_t + row * C + c0, dv);
      }
    }
  }
  if (dgamma == nullptr) return;
#pragma unroll
  for (int i = 0; i < EPL; ++i) { red[0][warp][lane * EPL + i] = ag[i]; red[1][warp][lane * EPL + i] = ab[i]; }
  __syncthreads();
  // 16 B vector reductions (C is a multiple of EPL >= 4): every CTA ends with 2 C / 4 atomic operations instead of 2 C
  for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    for (int w = 0; w < wpb; ++w) {
      const float4 x0 = *reinterpret_cast<const float4*>(&red[0][w][c]), x1 = *reinterpret_cast<const float4*>(&red[1][w][c]);
      a.x += x0.x; a.y += x0.y; a.z += x0.z; a.w += x0.w;
      b.x += x1.x; b.y += x1.y; b.z += x1.z; b.w += x1.w;
    }
    red_add_v4(dgamma + c, a.x, a.y, a.z, a.w);
    red_add_v4(dbeta + c, b.x, b.y, b.z, b.w);
  }
}

int pick_epl(int C) { return C > 128 ? 8 : 4; }

}  // namespace

int ln_fwd(cudaStream_t s, int dt_in, const void* x, int ldx, int rows, int C, const float* gamma, const float* beta,
           float eps, int gelu_in, const float* gamma2, const float* beta2, int dt_out, void* y, int ldy, float* stats) {
  if (rows <= 0) return 0;
  const int epl = pick_epl(C);
  QV_CHECK(C <= 256 && C % epl == 0 && ldx % epl == 0 && ldy % epl == 0, "ln_fwd: C=%d ldx=%d ldy=%d unsupported", C, ldx, ldy);
  const int grid = max(1, min(cdiv(rows, 8), qv_num_sms() * 8));
#define LN_F3(TI, TO, E, G) \
  qv_launch(ln_fwd_kernel<TI, TO, E, G>, grid, 256, 0, s, (const TI*)x, ldx, rows, C, gamma, beta, eps, gamma2, beta2, (TO*)y, ldy, stats)
#define LN_F2(TI, TO, G) do { if (epl == 8) LN_F3(TI, TO, 8, G); else LN_F3(TI, TO, 4, G); } while (0)
#define LN_F(TI, TO) do { if (gelu_in) LN_F2(TI, TO, true); else LN_F2(TI, TO, false); } while (0)
  if (dt_in == QV_F32 && dt_out == QV_F32) LN_F(float, float);
  else if (dt_in == QV_F32 && dt_out == QV_BF16) LN_F(float, bf16);
  else if (dt_in == QV_BF16 && dt_out == QV_BF16) LN_F(bf16, bf16);
  else LN_F(bf16, float);
#undef LN_F
#undef LN_F2
#undef LN_F3
  QV_LAUNCH_CHECK();
  return 0;
}

int ln_bwd(cudaStream_t s, int dt_x, const void* x, int ldx, int dt_dy, const void* dy, int lddy, int rows, int C,
           const float* gamma, const float* stats, int gelu_in, int dt_out, void* dx_t, float* dx_f32,
           const float* resid, float* dgamma, float* dbeta, const DropP* drop, const float* rowscale, int rows_per_img) {
  if (rows <= 0) return 0;
  const int epl = pick_epl(C);
  QV_CHECK(C <= 256 && C % epl == 0 && ldx % epl == 0 && lddy % epl == 0, "ln_bwd: C=%d ldx=%d lddy=%d unsupported", C, ldx, lddy);
  const DropP dp = drop ? *drop : DropP();
  QV_CHECK((dp.p == 0.f && !rowscale) || (epl == 8 && dx_t), "ln_bwd: fused dropout needs C > 128 (8 elements per lane) and a T output");
  QV_CHECK(resid == nullptr || (resid != dx_f32 && (const void*)resid != dx_t), "ln_bwd: resid must not alias an output");
  QV_CHECK((((uintptr_t)dgamma | (uintptr_t)dbeta) & 15) == 0, "ln_bwd: dgamma / dbeta must be 16 B aligned (vector reductions)");
  // every CTA ends with 2 * C atomics: keep >= 32 rows per warp before adding CTAs
  const int grid = max(1, min(cdiv(rows, 8 * 32), qv_num_sms() * 6));
#define LN_B4(TX, TDY, TO, E, G)                                                                                      \
  qv_launch(ln_bwd_kernel<TX, TDY, TO, E, G>, grid, 256, 0, s, (const TX*)x, ldx, (const TDY*)dy, lddy, rows, C, gamma, stats, \
                                                        (TO*)dx_t, dx_f32, resid, dgamma, dbeta, dp, rowscale, rows_per_img)
#define LN_B3(TX, TDY, TO, G) do { if (epl == 8) LN_B4(TX, TDY, TO, 8, G); else LN_B4(TX, TDY, TO, 4, G); } while (0)
#define LN_B2(TX, TDY, TO) do { if (gelu_in) LN_B3(TX, TDY, TO, true); else LN_B3(TX, TDY, TO, false); } while (0)
  const int key = dt_x * 4 + dt_dy * 2 + dt_out;
  switch (key) {
    case 0: LN_B2(float, float, float); break;
    case 1: LN_B2(float, float, bf16); break;
    case 2: LN_B2(float, bf16, float); break;
    case 3: LN_B2(float, bf16, bf16); break;
    case 4: LN_B2(bf16, float, float); break;
    case 5: LN_B2(bf16, float, bf16); break;
    case 6: LN_B2(bf16, bf16, float); break;
    default: LN_B2(bf16, bf16, bf16); break;
  }
#undef LN_B2
#undef LN_B3
#undef LN_B4
  QV_LAUNCH_CHECK();
  return 0;
}
