// LayerNorm forward / backward: one warp per row, row held in registers (C <= 256, C even),
// two-pass statistics in fp32.  HBM-bound kernels: algorithmic bytes = rows*C*(in + out) (+8 B/row stats).
//   gelu_in : the LN input is gelu(x) of the stored pre-activation x (CCF-FFN fc1 -> GELU -> LN, H:704-706)
//   chain   : y = LN2(LN1(x))  (branch .norm followed by bank write_norm, H:468 + H:301)
#include "kernels.h"

namespace {

constexpr int MAXV = 8;  // values per lane (C <= 256)

template <typename TI, typename TO, bool GELU_IN>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const TI* __restrict__ x, int ldx, int rows, int C,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     float eps, const float* __restrict__ gamma2,
                                                     const float* __restrict__ beta2, TO* __restrict__ y, int ldy,
                                                     float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int nper = (C + 31) / 32;
  const float invC = 1.f / (float)C;
  for (long row = (long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long)gridDim.x * wpb) {
    float v[MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + i * 32;
      v[i] = 0.f;
      if (i < nper && c < C) {
        float t = ldf(x + row * ldx + c);
        if (GELU_IN) t = gelu_f(t);
        v[i] = t;
        s += t;
      }
    }
    const float mean = warp_sum(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + i * 32;
      if (i < nper && c < C) { const float d = v[i] - mean; q += d * d; }
    }
    const float rstd = rsqrtf(warp_sum(q) * invC + eps);
    if (stats && lane == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + i * 32;
      if (i < nper && c < C) v[i] = (v[i] - mean) * rstd * gamma[c] + beta[c];
    }
    if (gamma2) {  // chained second LayerNorm (no stats kept: the write path has no backward)
      float s2 = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) { const int c = lane + i * 32; if (i < nper && c < C) s2 += v[i]; }
      const float m2 = warp_sum(s2) * invC;
      float q2 = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) { const int c = lane + i * 32; if (i < nper && c < C) { const float d = v[i] - m2; q2 += d * d; } }
      const float r2 = rsqrtf(warp_sum(q2) * invC + eps);
#pragma unroll
      for (int i = 0; i < MAXV; ++i) { const int c = lane + i * 32; if (i < nper && c < C) v[i] = (v[i] - m2) * r2 * gamma2[c] + beta2[c]; }
    }
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + i * 32;
      if (i < nper && c < C) stf(y + row * ldy + c, v[i]);
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) [+ resid],  g = dy * gamma;  dgamma += dy * xhat, dbeta += dy.
// dx is written as T (dx_t) and/or fp32 (dx_f32); resid may alias dx_f32 (same element, same thread).
template <typename TX, typename TDY, typename TO, bool GELU_IN>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const TX* __restrict__ x, int ldx, const TDY* __restrict__ dy,
                                                     int lddy, int rows, int C, const float* __restrict__ gamma,
                                                     const float* __restrict__ stats, TO* __restrict__ dx_t,
                                                     float* dx_f32, const float* resid,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float red[2][8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const int nper = (C + 31) / 32;
  const float invC = 1.f / (float)C;
  float ag[MAXV], ab[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) ag[i] = ab[i] = 0.f;
  for (long row = (long)blockIdx.x * wpb + warp; row < rows; row += (long)gridDim.x * wpb) {
    const float mean = stats[2 * row], rstd = stats[2 * row + 1];
    float xh[MAXV], g[MAXV], u[MAXV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + i * 32;
      xh[i] = g[i] = u[i] = 0.f;
      if (i < nper && c < C) {
        float t = ldf(x + row * ldx + c);
        if (GELU_IN) { u[i] = t; t = gelu_f(t); }
        const float d = ldf(dy + row * lddy + c);
        xh[i] = (t - mean) * rstd;
        g[i] = d * gamma[c];
        ag[i] += d * xh[i];
        ab[i] += d;
        s1 += g[i];
        s2 += g[i] * xh[i];
      }
    }
    const float c1 = warp_sum(s1) * invC, c2 = warp_sum(s2) * invC;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + i * 32;
      if (i < nper && c < C) {
        float d = rstd * (g[i] - c1 - xh[i] * c2);
        if (GELU_IN) d *= gelu_grad_f(u[i]);
        if (resid) d += resid[row * C + c];
        if (dx_t) stf(dx_t + row * C + c, d);
        if (dx_f32) dx_f32[row * C + c] = d;
      }
    }
  }
  if (dgamma == nullptr) return;
  // block reduction of the per-warp partial dgamma / dbeta, then one atomic per channel per block
#pragma unroll
  for (int i = 0; i < MAXV; ++i) { red[0][warp][lane + i * 32] = ag[i]; red[1][warp][lane + i * 32] = ab[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < wpb; ++w) { a += red[0][w][c]; b += red[1][w][c]; }
    atomicAdd(dgamma + c, a);
    atomicAdd(dbeta + c, b);
  }
}

int grid_for(int rows) { return max(1, min(cdiv(rows, 8), qv_num_sms() * 4)); }

}  // namespace

int ln_fwd(cudaStream_t s, int dt_in, const void* x, int ldx, int rows, int C, const float* gamma, const float* beta,
           float eps, int gelu_in, const float* gamma2, const float* beta2, int dt_out, void* y, int ldy, float* stats) {
  if (rows <= 0) return 0;
  QV_CHECK(C <= 256, "ln_fwd: C=%d > 256", C);
  const int grid = grid_for(rows);
#define LN_F(TI, TO, G) \
  ln_fwd_kernel<TI, TO, G><<<grid, 256, 0, s>>>((const TI*)x, ldx, rows, C, gamma, beta, eps, gamma2, beta2, (TO*)y, ldy, stats)
  if (dt_in == QV_F32 && dt_out == QV_F32) { if (gelu_in) LN_F(float, float, true); else LN_F(float, float, false); }
  else if (dt_in == QV_F32 && dt_out == QV_BF16) { if (gelu_in) LN_F(float, bf16, true); else LN_F(float, bf16, false); }
  else if (dt_in == QV_BF16 && dt_out == QV_BF16) { if (gelu_in) LN_F(bf16, bf16, true); else LN_F(bf16, bf16, false); }
  else if (dt_in == QV_BF16 && dt_out == QV_F32) { if (gelu_in) LN_F(bf16, float, true); else LN_F(bf16, float, false); }
#undef LN_F
  QV_LAUNCH_CHECK();
  return 0;
}

int ln_bwd(cudaStream_t s, int dt_x, const void* x, int ldx, int dt_dy, const void* dy, int lddy, int rows, int C,
           const float* gamma, const float* stats, int gelu_in, int dt_out, void* dx_t, float* dx_f32,
           const float* resid, float* dgamma, float* dbeta) {
  if (rows <= 0) return 0;
  QV_CHECK(C <= 256, "ln_bwd: C=%d > 256", C);
  const int grid = max(1, min(cdiv(rows, 8 * 4), qv_num_sms() * 2));
#define LN_B(TX, TDY, TO, G)                                                                                      \
  ln_bwd_kernel<TX, TDY, TO, G><<<grid, 256, 0, s>>>((const TX*)x, ldx, (const TDY*)dy, lddy, rows, C, gamma, stats, \
                                                     (TO*)dx_t, dx_f32, resid, dgamma, dbeta)
#define LN_B2(TX, TDY, TO) do { if (gelu_in) LN_B(TX, TDY, TO, true); else LN_B(TX, TDY, TO, false); } while (0)
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
#undef LN_B
  QV_LAUNCH_CHECK();
  return 0;
}
