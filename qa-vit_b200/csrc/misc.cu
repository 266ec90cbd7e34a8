// Per-image and elementwise kernels of the quad block: bank write, MSDA pooling, depthwise 3x3,
// TokenLearner / TokenUpMix, fusion weights, GELU / gamma backward, column sums, weight conversion.
// All HBM-bound byte movers: coalesced along the channel axis, fp32 math, T = float | bf16 storage.
#include "kernels.h"

#define DISPATCH_T(dt, ...)                         \
  do {                                              \
    if ((dt) == QV_F32) { typedef float T; __VA_ARGS__; } \
    else { typedef bf16 T; __VA_ARGS__; }           \
  } while (0)

// =============================================================================== bank write (H:296-321)
namespace {
// tn[B*Nt, d] (already write_norm'ed), cg[B*Nt, d + kb] = [write_compression(tn) | write_gate(tn)].
// partial[cta][0] = sum_b S_b^T c_b, partial[cta][1] = sum_b S_b^T tn_b with S_b = softmax over tokens of the gate.
// CTA strides over images; thread = channel (its 2 x KB accumulators stay in registers for all images of the CTA);
// the softmaxed gate tile [Nt][KB] sits in shared memory and is read as broadcast float4s (8 FMAs per shared load);
// the gate logits of the next image are prefetched into registers while the current image is accumulated.
// NQ = gate-tile entries per thread: Nt * KB <= NQ * 256 (4 covers the 16 / 64-token blocks, 16 the 196 tokens of 224 / 16).
template <typename T, int KB, int NQ>
__global__ void __launch_bounds__(256) bank_write_reduce_kernel(const T* __restrict__ tn, const T* __restrict__ cg, int ldcg,
                                                                int B, int Nt, int d, float* __restrict__ partial) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) float g[];  // [Nt][KB]
  const int c = threadIdx.x;
  const int ng = Nt * KB;                     // <= NQ * blockDim.x (checked by the launcher)
  float ak[KB], av[KB];
#pragma unroll
  for (int s = 0; s < KB; ++s) ak[s] = av[s] = 0.f;
  float pre[NQ];
  auto fetch = [&](int b) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int idx = threadIdx.x + q * blockDim.x;
      pre[q] = idx < ng ? ldf(cg + ((long)b * Nt + idx / KB) * ldcg + d + idx % KB) : 0.f;
    }
  };
  if ((int)blockIdx.x < B) fetch(blockIdx.x);
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();                          // previous image's accumulation is done with g
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int idx = threadIdx.x + q * blockDim.x;
      if (idx < ng) g[idx] = pre[q];
    }
    __syncthreads();
    // softmax over the token axis, every thread normalises its own entries (column statistics recomputed per thread)
    float e[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int idx = threadIdx.x + q * blockDim.x;
      e[q] = 0.f;
      if (idx < ng) {
        const int sl = idx % KB;
        float m = -INFINITY;
        for (int n = 0; n < Nt; ++n) m = fmaxf(m, g[n * KB + sl]);
        float z = 0.f;
        for (int n = 0; n < Nt; ++n) z += __expf(g[n * KB + sl] - m);
        e[q] = __expf(g[idx] - m) / z;
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int idx = threadIdx.x + q * blockDim.x;
      if (idx < ng) g[idx] = e[q];
    }
    if (b + (int)gridDim.x < B) fetch(b + gridDim.x);
    __syncthreads();
    if (c < d) {
      // loads are independent: issue 8 tokens' worth before the FMAs
      for (int n0 = 0; n0 < Nt; n0 += 8) {
        float cv[8], tv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int n = min(n0 + k, Nt - 1);
          cv[k] = ldf(cg + ((long)b * Nt + n) * ldcg + c);
          tv[k] = ldf(tn + ((long)b * Nt + n) * d + c);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (n0 + k < Nt) {
#pragma unroll
            for (int q = 0; q < KB / 4; ++q) {
              const float4 w = *reinterpret_cast<const float4*>(g + (n0 + k) * KB + 4 * q);
              ak[4 * q] = fmaf(w.x, cv[k], ak[4 * q]); ak[4 * q + 1] = fmaf(w.y, cv[k], ak[4 * q + 1]);
              ak[4 * q + 2] = fmaf(w.z, cv[k], ak[4 * q + 2]); ak[4 * q + 3] = fmaf(w.w, cv[k], ak[4 * q + 3]);
              av[4 * q] = fmaf(w.x, tv[k], av[4 * q]); av[4 * q + 1] = fmaf(w.y, tv[k], av[4 * q + 1]);
              av[4 * q + 2] = fmaf(w.z, tv[k], av[4 * q + 2]); av[4 * q + 3] = fmaf(w.w, tv[k], av[4 * q + 3]);
            }
          }
        }
      }
    }
  }
  if (c < d) {
    float* pk = partial + (long)blockIdx.x * 2 * KB * d;
#pragma unroll
    for (int s = 0; s < KB; ++s) { pk[s * d + c] = ak[s]; pk[KB * d + s * d + c] = av[s]; }
  }
}

// mean over batch -> clamp -> bank += rate * u -> clamp -> update_count += 1.  One thread per bank element, the
// partials are summed in a fixed order (deterministic); the last CTA to finish bumps the write counter, which every
// CTA has read before taking its ticket.
__device__ unsigned int g_bank_ticket = 0;
__global__ void __launch_bounds__(256) bank_write_apply_kernel(const float* __restrict__ partial, int n_partial, int B,
                                                               int n, float* __restrict__ bank_k,
                                                               float* __restrict__ bank_v,
                                                               long long* __restrict__ update_count, int v1) {
  QV_PDL_ENTRY();
  float uclamp, rate, bclamp;
  if (v1) { uclamp = 0.1f; rate = 0.01f; bclamp = 1.0f; }                    // QAViT.py:217-224
  else { uclamp = 0.05f; bclamp = 0.5f; rate = (*update_count < 1000) ? 0.005f : 0.01f; }   // H:310-319
  const float invB = 1.f / (float)B;
  // CTA = 32 bank elements x 8 partial groups: coalesced 128 B reads, the 8 group sums are combined in a fixed order
  __shared__ float gs[8][33];
  const int e = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + e;
  float acc = 0.f;
  if (i < 2 * n)
    for (int p = grp; p < n_partial; p += 8) acc += partial[(long)p * 2 * n + i];
  gs[grp][e] = acc;
  __syncthreads();
  if (grp == 0 && i < 2 * n) {
    float sum = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) sum += gs[q][e];
    const float u = fminf(fmaxf(sum * invB, -uclamp), uclamp);
    float* dst = (i < n) ? bank_k + i : bank_v + (i - n);
    *dst = fminf(fmaxf(*dst + rate * u, -bclamp), bclamp);
  }
  if (v1) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(&g_bank_ticket, 1u);
    if (t == gridDim.x - 1) {
      *update_count += 1;
      g_bank_ticket = 0;
    }
  }
}
}  // namespace

int bank_write_reduce(cudaStream_t s, int dt, const void* tn, const void* cg, int ldcg, int B, int Nt, int d, int kb,
                      float* partial, int* n_partial) {
  if (dt == QV_BF16 && bank_write_mma_ok(Nt, d, kb, ldcg)) return bank_write_reduce_mma(s, tn, cg, ldcg, B, Nt, d, partial, n_partial);
  QV_CHECK(kb == 16, "bank write kernel is instantiated for bank size 16 (got %d)", kb);
  QV_CHECK(d <= 256, "bank write: d=%d > 256", d);
  QV_CHECK(Nt * kb <= 16 * 256, "bank write: %d tokens x %d slots exceed the gate tile (<= 256 tokens)", Nt, kb);
  const int grid = max(1, min(cdiv(B, 4), min(592, qv_num_sms() * 4)));   // >= 4 images per CTA; 4 CTAs / SM for latency hiding
  *n_partial = grid;
  const size_t smem = (size_t)Nt * kb * sizeof(float);
  if (Nt * kb <= 4 * 256) {
    DISPATCH_T(dt, (qv_launch(bank_write_reduce_kernel<T, 16, 4>, grid, 256, smem, s, (const T*)tn, (const T*)cg, ldcg, B, Nt, d, partial)));
  } else {
    DISPATCH_T(dt, (qv_launch(bank_write_reduce_kernel<T, 16, 16>, grid, 256, smem, s, (const T*)tn, (const T*)cg, ldcg, B, Nt, d, partial)));
  }
  QV_LAUNCH_CHECK();
  return 0;
}

int bank_write_apply(cudaStream_t s, const float* partial, int n_partial, int B, int d, int kb, float* bank_k,
                     float* bank_v, long long* update_count, int v1) {
  qv_launch(bank_write_apply_kernel, cdiv(2 * kb * d, 32), 256, 0, s, partial, n_partial, B, kb * d, bank_k, bank_v, update_count, v1);
  QV_LAUNCH_CHECK();
  return 0;
}

// =============================================================================== MSDA pooling (H:489-501)
namespace {
struct DilP { int dil[4]; int ndil; };
// multi-scale token m -> source token index in the side x side grid
__device__ __forceinline__ int msda_src(int m, int side, const DilP& dp) {
  for (int i = 0; i < dp.ndil; ++i) {
    const int d = dp.dil[i], nd = (side + d - 1) / d, cnt = nd * nd;
    if (m < cnt) return (m / nd) * d * side + (m % nd) * d;
    m -= cnt;
  }
  return 0;
}
// xp[b, t, :] = mean_k xn[b, src(t * stride + k), :].  The source-token table is built once per CTA in shared memory;
// a thread moves 4 channels of one pooled token.
template <typename T>
__global__ void __launch_bounds__(256) msda_pool_fwd_kernel(const T* __restrict__ xn, int B, int Nt, int side, int C, DilP dp,
                                                            int stride, int NM, T* __restrict__ xp) {
  QV_PDL_ENTRY();
  extern __shared__ int src[];   // [NM * stride]
  for (int i = threadIdx.x; i < NM * stride; i += blockDim.x) src[i] = msda_src(i, side, dp);
  __syncthreads();
  const int cv = C / 4;
  const long total = (long)B * NM * cv;
  const float inv = 1.f / (float)stride;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % cv) * 4;
    const long r = idx / cv;
    const int t = (int)(r % NM);
    const long b = r / NM;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < stride; ++k) {
      float v[4];
      load_vec<4>(xn + (b * Nt + src[t * stride + k]) * C + c4, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) a[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] *= inv;
    store_vec<4>(xp + r * C + c4, a);
  }
}
// dxn[b, src(t * stride + k), :] += dxp[b, t, :] / stride.  A thread owns 4 channels of one image and walks the pooled
// tokens sequentially (several pooled tokens can hit the same source token): no write races.
template <typename T>
__global__ void __launch_bounds__(256) msda_pool_bwd_kernel(const T* __restrict__ dxp, int B, int Nt, int side, int C, DilP dp,
                                                            int stride, int NM, float* __restrict__ dxn) {
  QV_PDL_ENTRY();
  extern __shared__ int src[];   // [NM * stride]
  for (int i = threadIdx.x; i < NM * stride; i += blockDim.x) src[i] = msda_src(i, side, dp);
  __syncthreads();
  const int cv = C / 4;
  const long total = (long)B * cv;
  const float inv = 1.f / (float)stride;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % cv) * 4;
    const long b = idx / cv;
    for (int t = 0; t < NM; ++t) {
      float g[4];
      load_vec<4>(dxp + (b * NM + t) * C + c4, g);
      for (int k = 0; k < stride; ++k) {
        float4* d = reinterpret_cast<float4*>(dxn + (b * Nt + src[t * stride + k]) * C + c4);
        float4 o = *d;
        o.x += g[0] * inv; o.y += g[1] * inv; o.z += g[2] * inv; o.w += g[3] * inv;
        *d = o;
      }
    }
  }
}
int make_dilp(const int* dil, int ndil, DilP* dp) {
  QV_CHECK(ndil >= 1 && ndil <= 4, "msda: %d dilation factors (1..4 supported)", ndil);
  dp->ndil = ndil;
  for (int i = 0; i < 4; ++i) dp->dil[i] = i < ndil ? dil[i] : 1;
  return 0;
}
}  // namespace

int msda_pool_fwd(cudaStream_t s, int dt, const void* xn, int B, int Nt, int side, int C, const int* dil, int ndil,
                  int stride, int NM, void* xp) {
  DilP dp;
  QV_TRY(make_dilp(dil, ndil, &dp));
  const long total = (long)B * NM * C;
  if (total <= 0) return 0;
  QV_CHECK(C % 4 == 0, "msda pooling: C=%d must be a multiple of 4", C);
  const int grid = (int)max(1L, min((long)qv_num_sms() * 8, (total / 4 + 255) / 256));
  DISPATCH_T(dt, (qv_launch(msda_pool_fwd_kernel<T>, grid, 256, (size_t)NM * stride * sizeof(int), s, (const T*)xn, B, Nt, side, C, dp, stride, NM, (T*)xp)));
  QV_LAUNCH_CHECK();
  return 0;
}

int msda_pool_bwd(cudaStream_t s, int dt, const void* dxp, int B, int Nt, int side, int C, const int* dil, int ndil,
                  int stride, int NM, float* dxn) {
  DilP dp;
  QV_TRY(make_dilp(dil, ndil, &dp));
  if (B <= 0) return 0;
  QV_CHECK(C % 4 == 0, "msda pooling: C=%d must be a multiple of 4", C);
  const int grid = (int)max(1L, min((long)qv_num_sms() * 8, ((long)B * (C / 4) + 255) / 256));
  DISPATCH_T(dt, (qv_launch(msda_pool_bwd_kernel<T>, grid, 256, (size_t)NM * stride * sizeof(int), s, (const T*)dxp, B, Nt, side, C, dp, stride, NM, dxn)));
  QV_LAUNCH_CHECK();
  return 0;
}

// =============================================================================== depthwise 3x3 (H:670-675)
namespace {
// y[b, p, c] = (sum_taps w[c, ky, kx] x[b, (py+ky-1, px+kx-1), c] + bias[c]) * scale[c]
template <typename T>
__global__ void dwconv_fwd_kernel(const T* __restrict__ x, int B, int side, int C, const float* __restrict__ w,
                                  const float* __restrict__ bias, const float* __restrict__ scale, T* __restrict__ y) {
  QV_PDL_ENTRY();
  const int Nt = side * side;
  const long total = (long)B * Nt * C;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c = idx % C;
    const long r = idx / C;
    const int p = r % Nt;
    const long b = r / Nt;
    const int py = p / side, px = p % side;
    float a = bias ? bias[c] : 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = py + ky - 1;
      if (yy < 0 || yy >= side) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = px + kx - 1;
        if (xx < 0 || xx >= side) continue;
        a = fmaf(w[c * 9 + ky * 3 + kx], ldf(x + (b * Nt + yy * side + xx) * C + c), a);
      }
    }
    stf(y + idx, scale ? a * scale[c] : a);
  }
}
// thread = channel; CTA strides over images; dw / dbias / dscale accumulate in registers, one atomic flush.
template <typename T>
__global__ void dwconv_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, int B, int side, int C,
                                  const float* __restrict__ w, const float* __restrict__ bias,
                                  const float* __restrict__ scale, T* __restrict__ dx, float* __restrict__ dw,
                                  float* __restrict__ dbias, float* __restrict__ dscale) {
  QV_PDL_ENTRY();
  const int c = threadIdx.x;
  if (c >= C) return;
  const int Nt = side * side;
  float wr[9], aw[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) { wr[t] = w[c * 9 + t]; aw[t] = 0.f; }
  const float sc = scale ? scale[c] : 1.f;
  const float bs = bias ? bias[c] : 0.f;
  float ab = 0.f, as = 0.f;
  for (long b = blockIdx.x; b < B; b += gridDim.x) {
    for (int p = 0; p < Nt; ++p) {
      const int py = p / side, px = p % side;
      // dx at p gathers d_raw from neighbours; conv at p (for dscale) gathers x from neighbours
      float gx = 0.f, conv = bs;
      const float draw_p = ldf(dy + (b * Nt + p) * C + c) * sc;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int yy = py + ky - 1, xx = px + kx - 1;       // input position feeding output p through tap (ky,kx)
          if (yy >= 0 && yy < side && xx >= 0 && xx < side) {
            const float xv = ldf(x + (b * Nt + yy * side + xx) * C + c);
            conv = fmaf(wr[ky * 3 + kx], xv, conv);
            aw[ky * 3 + kx] = fmaf(draw_p, xv, aw[ky * 3 + kx]);
          }
          const int oy = py - ky + 1, ox = px - kx + 1;       // output position that reads input p through tap (ky,kx)
          if (oy >= 0 && oy < side && ox >= 0 && ox < side)
            gx = fmaf(wr[ky * 3 + kx], ldf(dy + (b * Nt + oy * side + ox) * C + c) * sc, gx);
        }
      }
      stf(dx + (b * Nt + p) * C + c, gx);
      ab += draw_p;
      as = fmaf(ldf(dy + (b * Nt + p) * C + c), conv, as);
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) atomicAdd(dw + c * 9 + t, aw[t]);
  if (dbias) atomicAdd(dbias + c, ab);
  if (dscale) atomicAdd(dscale + c, as);
}
}  // namespace

int dwconv_fwd(cudaStream_t s, int dt, const void* x, int B, int side, int C, const float* w, const float* bias,
               const float* scale, void* y) {
  const long total = (long)B * side * side * C;
  if (total <= 0) return 0;
  if (C % 2 == 0 && (side == 4 || side == 8 || side == 16)) {   // shared-memory tiled stencil (dwconv_nhwc.cu)
    DwP p{};
    p.x = x; p.ldx = C; p.B = B; p.H = side; p.W = side; p.C = C; p.K = 3; p.w = w; p.bias = bias; p.scale = scale; p.y = y; p.ldy = C;
    return dw2d_fwd(s, dt, p, false);
  }
  const int grid = (int)min((long)qv_num_sms() * 8, (total + 255) / 256);
  DISPATCH_T(dt, (qv_launch(dwconv_fwd_kernel<T>, grid, 256, 0, s, (const T*)x, B, side, C, w, bias, scale, (T*)y)));
  QV_LAUNCH_CHECK();
  return 0;
}

int dwconv_bwd(cudaStream_t s, int dt, const void* x, const void* dy, int B, int side, int C, const float* w,
               const float* bias, const float* scale, void* dx, float* dw, float* dbias, float* dscale) {
  if (B <= 0) return 0;
  if (C % 2 == 0 && (side == 4 || side == 8 || side == 16)) {
    DwScale sc;
    sc.w = w; sc.bias = bias; sc.scale = scale; sc.dscale = dscale;
    QV_TRY(dw2d_wgrad(s, dt, 3, x, C, dy, C, B, side, side, C, dw, dbias, sc));
    DwP p{};
    p.x = dy; p.ldx = C; p.B = B; p.H = side; p.W = side; p.C = C; p.K = 3; p.w = w; p.scale = scale; p.y = dx; p.ldy = C;
    return dw2d_fwd(s, dt, p, true);
  }
  QV_CHECK(C <= 256, "dwconv_bwd: C=%d > 256", C);
  const int grid = min(B, qv_num_sms() * 8);
  const int threads = ((C + 31) / 32) * 32;
  DISPATCH_T(dt, (qv_launch(dwconv_bwd_kernel<T>, grid, threads, 0, s, (const T*)x, (const T*)dy, B, side, C, w, bias, scale,
                                                             (T*)dx, dw, dbias, dscale)));
  QV_LAUNCH_CHECK();
  return 0;
}

// =============================================================================== elementwise
namespace {
template <typename T>
__global__ void gelu_bwd_kernel(const T* __restrict__ pre, const T* __restrict__ dact, long n, T* __restrict__ dpre) {
  QV_PDL_ENTRY();
  for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += (long)gridDim.x * blockDim.x * 2) {
    const float2 u = ld2(pre + i), g = ld2(dact + i);
    st2(dpre + i, make_float2(g.x * gelu_grad_f(u.x), g.y * gelu_grad_f(u.y)));
  }
}
// d_o = gamma * dout (T);  dgamma += sum(dout * o)
template <typename T>
__global__ void gamma_bwd_kernel(const float* __restrict__ dout, const T* __restrict__ o, long n,
                                 const float* __restrict__ gamma, T* __restrict__ d_o, float* __restrict__ dgamma) {
  QV_PDL_ENTRY();
  __shared__ float red[32];
  const float g = *gamma;
  float acc = 0.f;
  for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += (long)gridDim.x * blockDim.x * 2) {
    const float2 d = ld2(dout + i), ov = ld2(o + i);
    acc = fmaf(d.x, ov.x, fmaf(d.y, ov.y, acc));
    st2(d_o + i, make_float2(g * d.x, g * d.y));
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(dgamma, v);
  }
}
// same with a dropout site / DropPath scale on d_o (8 elements per thread: one Philox call); gamma == nullptr: cast only
template <typename T>
__global__ void __launch_bounds__(256) gamma_bwd_drop_kernel(const float* __restrict__ dout, const T* __restrict__ o, long n,
                                                             const float* __restrict__ gamma, T* __restrict__ d_o,
                                                             float* __restrict__ dgamma, DropP drop,
                                                             const float* __restrict__ rowscale, int rows_per_img, int C) {
  QV_PDL_ENTRY();
  __shared__ float red[32];
  const float g = gamma ? *gamma : 1.f;
  const bool masked = drop.p > 0.f;
  DropState dst{};
  if (masked) dst = drop_state(drop);
  float acc = 0.f;
  for (long i8 = (long)blockIdx.x * blockDim.x + threadIdx.x; i8 * 8 < n; i8 += (long)gridDim.x * blockDim.x) {
    const long i = i8 * 8;
    float d[8], k[8];
    load_vec<8>(dout + i, d);
    if (gamma) {
      float ov[8];
      load_vec<8>(o + i, ov);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(d[j], ov[j], acc);
    }
    const float rs = rowscale ? rowscale[(i / C) / rows_per_img] : 1.f;
    if (masked) drop_keep8(dst, (unsigned long long)i8, k);
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] *= g * rs * (masked ? k[j] : 1.f);
    store_vec<8>(d_o + i, d);
  }
  if (!gamma) return;
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(dgamma, v);
  }
}
template <typename T>
__global__ void cast_kernel(const float* __restrict__ x, long n, T* __restrict__ y) {
  QV_PDL_ENTRY();
  for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += (long)gridDim.x * blockDim.x * 2)
    st2(y + i, ld2(x + i));
}
__global__ void fusion_softmax_kernel(const float* w, int n, float* alpha) {
  QV_PDL_ENTRY();
  if (threadIdx.x == 0) {
    float m = -INFINITY, z = 0.f;
    for (int i = 0; i < n; ++i) m = fmaxf(m, w[i]);
    for (int i = 0; i < n; ++i) z += expf(w[i] - m);
    for (int i = 0; i < n; ++i) alpha[i] = expf(w[i] - m) / z;
  }
}
// raw[i] += sum over rows and the i-th column slice of dfused * fused   (fused_i = alpha_i * b_i)
template <typename T>
__global__ void __launch_bounds__(256) fusion_bwd_kernel(const T* __restrict__ df, const T* __restrict__ f, long rows, int nb, int cw,
                                                         float* __restrict__ raw) {
  QV_PDL_ENTRY();
  __shared__ float red[4][8];
  const int C = nb * cw, v4 = C / 4;            // 4-element vectors never straddle a branch slice (cw % 4 == 0)
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const long total = rows * v4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int br = (int)(i % v4) * 4 / cw;
    float a[4], g[4];
    load_vec<4>(df + i * 4, a);
    load_vec<4>(f + i * 4, g);
    const float v = a[0] * g[0] + a[1] * g[1] + a[2] * g[2] + a[3] * g[3];
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] += (br == k) ? v : 0.f;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float v = warp_sum(acc[k]);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < nb) {
    float v = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[threadIdx.x][w];
    atomicAdd(raw + threadIdx.x, v);
  }
}
// dalpha_i = raw_i / alpha_i ; dw_j += alpha_j (dalpha_j - sum_i alpha_i dalpha_i)
__global__ void fusion_bwd_final_kernel(const float* alpha, const float* raw, int nb, float* dw) {
  QV_PDL_ENTRY();
  if (threadIdx.x == 0) {
    float dot = 0.f;
    for (int i = 0; i < nb; ++i) dot += raw[i];   // alpha_i * (raw_i / alpha_i)
    for (int j = 0; j < nb; ++j) dw[j] += raw[j] - alpha[j] * dot;
  }
}
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ dY, int ldy, int M, int N, int rows_per_block,
                              float* __restrict__ db, const float* __restrict__ scale) {
  QV_PDL_ENTRY();
  __shared__ float red[8][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  const long r0 = (long)blockIdx.y * rows_per_block;
  const long r1 = min((long)M, r0 + rows_per_block);
  float acc = 0.f;
  if (col < N)
    for (long r = r0 + threadIdx.y; r < r1; r += 8) acc += ldf(dY + r * ldy + col);
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && col < N) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v += red[i][threadIdx.x];
    atomicAdd(db + col, v * (scale ? *scale : 1.f));
  }
}
__global__ void convert_weight_kernel(const float* __restrict__ w, int N, int K, bf16* __restrict__ wb,
                                      bf16* __restrict__ wbt) {
  QV_PDL_ENTRY();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * K) return;
  const bf16 v = __float2bfloat16_rn(w[idx]);
  if (wb) wb[idx] = v;
  if (wbt) wbt[(long)(idx % K) * N + idx / K] = v;
}
__global__ void convert_weights_batched_kernel(ConvertJobs jobs) {
  QV_PDL_ENTRY();
  const ConvertJob jb = jobs.j[blockIdx.y];
  const int total = jb.N * jb.K;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const bf16 v = __float2bfloat16_rn(jb.w[idx]);
    jb.wb[idx] = v;
    if (jb.wbt) jb.wbt[(long)(idx % jb.K) * (jb.ldt ? jb.ldt : jb.N) + idx / jb.K] = v;
  }
  if (jb.bdst && blockIdx.x == 0)
    for (int i = threadIdx.x; i < jb.bn; i += blockDim.x) jb.bdst[i] = jb.bsrc[i];
}
int ew_grid(long n) { return (int)max(1L, min((long)qv_num_sms() * 8, (n / 2 + 255) / 256)); }
}  // namespace

int gelu_bwd(cudaStream_t s, int dt, const void* pre, const void* dact, long n, void* dpre) {
  if (n <= 0) return 0;
  DISPATCH_T(dt, (qv_launch(gelu_bwd_kernel<T>, ew_grid(n), 256, 0, s, (const T*)pre, (const T*)dact, n, (T*)dpre)));
  QV_LAUNCH_CHECK();
  return 0;
}
int gamma_bwd(cudaStream_t s, int dt, const float* dout, const void* o, long n, const float* gamma, void* d_o, float* dgamma,
              const DropP* drop, const float* rowscale, int rows_per_img, int C) {
  if (n <= 0) return 0;
  const DropP dp = drop ? *drop : DropP();
  if (dp.p > 0.f || rowscale || !gamma) {
    QV_CHECK(n % 8 == 0 && C > 0 && C % 8 == 0, "gamma_bwd: fused dropout needs C %% 8 == 0 (C = %d)", C);
    const int grid = (int)max(1L, min((long)qv_num_sms() * 16, (n / 8 + 255) / 256));
    DISPATCH_T(dt, (qv_launch(gamma_bwd_drop_kernel<T>, grid, 256, 0, s, dout, (const T*)o, n, gamma, (T*)d_o, dgamma, dp, rowscale, rows_per_img, C)));
    QV_LAUNCH_CHECK();
    return 0;
  }
  DISPATCH_T(dt, (qv_launch(gamma_bwd_kernel<T>, ew_grid(n), 256, 0, s, dout, (const T*)o, n, gamma, (T*)d_o, dgamma)));
  QV_LAUNCH_CHECK();
  return 0;
}
// two small fp32 copies as one launch (bank snapshots, stacked biases: a memcpy node each was 2 graph nodes per site, 80 per step)
namespace {
__global__ void copy2_f32_kernel(float* __restrict__ d0, const float* __restrict__ s0, int n0, float* __restrict__ d1,
                                 const float* __restrict__ s1, int n1) {
  QV_PDL_ENTRY();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n0 + n1; i += gridDim.x * blockDim.x) {
    if (i < n0) d0[i] = s0[i];
    else d1[i - n0] = s1[i - n0];
  }
}
}  // namespace
int copy2_f32(cudaStream_t s, float* d0, const float* s0, int n0, float* d1, const float* s1, int n1) {
  if (n0 + n1 <= 0) return 0;
  qv_launch(copy2_f32_kernel, min(64, (n0 + n1 + 255) / 256), 256, 0, s, d0, s0, n0, d1, s1, n1);
  QV_LAUNCH_CHECK();
  return 0;
}

int cast_f32_to_t(cudaStream_t s, int dt, const float* x, long n, void* y) {
  if (n <= 0) return 0;
  DISPATCH_T(dt, (qv_launch(cast_kernel<T>, ew_grid(n), 256, 0, s, x, n, (T*)y)));
  QV_LAUNCH_CHECK();
  return 0;
}
int fusion_softmax(cudaStream_t s, const float* w, int n, float* alpha) {
  qv_launch(fusion_softmax_kernel, 1, 32, 0, s, w, n, alpha);
  QV_LAUNCH_CHECK();
  return 0;
}
int fusion_bwd(cudaStream_t s, int dt, const void* dfused, const void* fused, long rows, int nb, int cw,
               const float* alpha, float* dalpha_raw) {
  (void)alpha;
  QV_CHECK(nb <= 4 && cw % 4 == 0, "fusion_bwd: %d branches (<= 4) of width %d (multiple of 4)", nb, cw);
  if (rows <= 0) return 0;
  const int grid = (int)max(1L, min((long)qv_num_sms() * 4, (rows * nb * cw + 255) / 256));
  DISPATCH_T(dt, (qv_launch(fusion_bwd_kernel<T>, grid, 256, 0, s, (const T*)dfused, (const T*)fused, rows, nb, cw, dalpha_raw)));
  QV_LAUNCH_CHECK();
  return 0;
}
int fusion_bwd_final(cudaStream_t s, const float* alpha, const float* dalpha_raw, int nb, float* dw) {
  qv_launch(fusion_bwd_final_kernel, 1, 32, 0, s, alpha, dalpha_raw, nb, dw);
  QV_LAUNCH_CHECK();
  return 0;
}
int colsum_accum(cudaStream_t s, int dt, const void* dY, int ldy, int M, int N, float* db, const float* scale) {
  if (M <= 0 || N <= 0) return 0;
  const int gx = cdiv(N, 32);
  int gy = max(1, min(cdiv(M, 64), cdiv(qv_num_sms() * 4, gx)));
  const int rpb = cdiv(M, gy);
  gy = cdiv(M, rpb);
  DISPATCH_T(dt, (qv_launch(colsum_kernel<T>, dim3(gx, gy), dim3(32, 8), 0, s, (const T*)dY, ldy, M, N, rpb, db, scale)));
  QV_LAUNCH_CHECK();
  return 0;
}
int convert_weights_batched(cudaStream_t s, const ConvertJobs& jobs) {
  if (jobs.n <= 0) return 0;
  qv_launch(convert_weights_batched_kernel, dim3(32, jobs.n), 256, 0, s, jobs);
  QV_LAUNCH_CHECK();
  return 0;
}
int convert_weight(cudaStream_t s, const float* w, int N, int K, bf16* wb, bf16* wbt) {
  qv_launch(convert_weight_kernel, cdiv((long)N * K, 256), 256, 0, s, w, N, K, wb, wbt);
  QV_LAUNCH_CHECK();
  return 0;
}

// =============================================================================== TokenLearner (H:985-1002)
namespace {
// S = softmax over tokens of logits[B*N, M];  xc[b, m, :] = sum_n S[b, n, m] x[b, n, :].  One CTA per image.
template <typename T>
__global__ void token_learner_fwd_kernel(const float* __restrict__ x, const T* __restrict__ logits, int B, int N, int M,
                                         int C, float* __restrict__ S, float* __restrict__ xc) {
  QV_PDL_ENTRY();
  extern __shared__ float sS[];  // [N][M]
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < N * M; idx += blockDim.x) sS[idx] = ldf(logits + (long)b * N * M + idx);
    __syncthreads();
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
      float mx = -INFINITY;
      for (int n = 0; n < N; ++n) mx = fmaxf(mx, sS[n * M + m]);
      float z = 0.f;
      for (int n = 0; n < N; ++n) { const float e = __expf(sS[n * M + m] - mx); sS[n * M + m] = e; z += e; }
      z = 1.f / z;
      for (int n = 0; n < N; ++n) sS[n * M + m] *= z;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < N * M; idx += blockDim.x) S[(long)b * N * M + idx] = sS[idx];
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      for (int m0 = 0; m0 < M; m0 += 16) {
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.f;
        for (int n = 0; n < N; ++n) {
          const float xv = x[((long)b * N + n) * C + c];
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[i] = fmaf(sS[n * M + m0 + i], xv, acc[i]);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) xc[((long)b * M + m0 + i) * C + c] = acc[i];
      }
    }
  }
}
// dS[n,m] = x[n,:] . dxc[m,:];  dlogits = S (dS - colsum_n(S dS));  dx[n,:] = sum_m S[n,m] dxc[m,:]
template <typename T>
__global__ void token_learner_bwd_kernel(const float* __restrict__ x, const float* __restrict__ S,
                                         const float* __restrict__ dxc, int B, int N, int M, int C,
                                         T* __restrict__ dlogits, float* __restrict__ dx) {
  QV_PDL_ENTRY();
  extern __shared__ float sm[];
  float* sS = sm;                 // [N][M]
  float* sdS = sS + N * M;        // [N][M]
  float* sD = sdS + N * M;        // [M][C + 1]
  float* st = sD + M * (C + 1);   // [M]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < N * M; idx += blockDim.x) sS[idx] = S[(long)b * N * M + idx];
    for (int idx = threadIdx.x; idx < M * C; idx += blockDim.x) sD[(idx / C) * (C + 1) + idx % C] = dxc[(long)b * M * C + idx];
    __syncthreads();
    for (int n = warp; n < N; n += nwarp) {
      float xv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { const int c = lane + 32 * i; xv[i] = c < C ? x[((long)b * N + n) * C + c] : 0.f; }
      for (int m = 0; m < M; ++m) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const int c = lane + 32 * i; if (c < C) a = fmaf(xv[i], sD[m * (C + 1) + c], a); }
        a = warp_sum(a);
        if (lane == 0) sdS[n * M + m] = a;
      }
    }
    __syncthreads();
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
      float t = 0.f;
      for (int n = 0; n < N; ++n) t = fmaf(sS[n * M + m], sdS[n * M + m], t);
      st[m] = t;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < N * M; idx += blockDim.x)
      stf(dlogits + (long)b * N * M + idx, sS[idx] * (sdS[idx] - st[idx % M]));
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      for (int n = 0; n < N; ++n) {
        float a = 0.f;
        for (int m = 0; m < M; ++m) a = fmaf(sS[n * M + m], sD[m * (C + 1) + c], a);
        dx[((long)b * N + n) * C + c] = a;
      }
    }
  }
}

// ---- TokenUpMix (H:1016-1031): up[b, n, c] = sum_m W[n, m] xc[b, m, c] + bias[n]
__global__ void token_upmix_fwd_kernel(const float* __restrict__ xc, int B, int M, int N, int C,
                                       const float* __restrict__ W, const float* __restrict__ bias,
                                       float* __restrict__ up) {
  QV_PDL_ENTRY();
  extern __shared__ float sm[];
  float* sW = sm;            // [N][M]
  float* sX = sm + N * M;    // [M][C]
  for (int idx = threadIdx.x; idx < N * M; idx += blockDim.x) sW[idx] = W[idx];
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < M * C; idx += blockDim.x) sX[idx] = xc[(long)b * M * C + idx];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      for (int n = 0; n < N; ++n) {
        float a = bias[n];
        for (int m = 0; m < M; ++m) a = fmaf(sW[n * M + m], sX[m * C + c], a);
        up[((long)b * N + n) * C + c] = a;
      }
    }
  }
}
// dxc[b,m,c] = sum_n W[n,m] dup[b,n,c];  dW[n,m] += sum_{b,c} dup[b,n,c] xc[b,m,c];  dbias[n] += sum_{b,c} dup[b,n,c]
__global__ void token_upmix_bwd_kernel(const float* __restrict__ xc, const float* __restrict__ dup, int B, int M, int N,
                                       int C, const float* __restrict__ W, float* __restrict__ dxc,
                                       float* __restrict__ dW, float* __restrict__ dbias) {
  QV_PDL_ENTRY();
  extern __shared__ float sm[];
  constexpr int NC = 16;                 // dup rows staged per pass
  float* sW = sm;                        // [N][M]
  float* sX = sW + N * M;                // [M][C + 1]
  float* sG = sX + M * (C + 1);          // [NC][C + 1]
  for (int idx = threadIdx.x; idx < N * M; idx += blockDim.x) sW[idx] = W[idx];
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < M * C; idx += blockDim.x) sX[(idx / C) * (C + 1) + idx % C] = xc[(long)b * M * C + idx];
    for (int idx = threadIdx.x; idx < M * C; idx += blockDim.x) dxc[(long)b * M * C + idx] = 0.f;
    for (int n0 = 0; n0 < N; n0 += NC) {
      __syncthreads();
      for (int idx = threadIdx.x; idx < NC * C; idx += blockDim.x)
        sG[(idx / C) * (C + 1) + idx % C] = dup[((long)b * N + n0 + idx / C) * C + idx % C];
      __syncthreads();
      for (int c = threadIdx.x; c < C; c += blockDim.x) {
        for (int m = 0; m < M; ++m) {
          float a = 0.f;
#pragma unroll
          for (int i = 0; i < NC; ++i) a = fmaf(sW[(n0 + i) * M + m], sG[i * (C + 1) + c], a);
          dxc[((long)b * M + m) * C + c] += a;
        }
      }
      for (int idx = threadIdx.x; idx < NC * M; idx += blockDim.x) {
        const int i = idx / M, m = idx % M;
        float a = 0.f;
        for (int c = 0; c < C; ++c) a = fmaf(sG[i * (C + 1) + c], sX[m * (C + 1) + c], a);
        atomicAdd(dW + (n0 + i) * M + m, a);
      }
      for (int i = threadIdx.x; i < NC; i += blockDim.x) {
        float a = 0.f;
        for (int c = 0; c < C; ++c) a += sG[i * (C + 1) + c];
        atomicAdd(dbias + n0 + i, a);
      }
    }
  }
}
template <typename K>
int opt_in_smem(K kernel, size_t bytes) {
  QV_CHECK(bytes <= 227 * 1024, "kernel needs %zu B of shared memory (> 227 KB): config not supported", bytes);
  if (bytes > 48 * 1024) QV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}
}  // namespace

int token_learner_fwd(cudaStream_t s, int dt, const float* x, const void* logits, int B, int N, int M, int C, float* S,
                      float* xc) {
  if (B <= 0) return 0;
  if (dt == QV_BF16 && tokens_mma_ok(M, N, C)) return tlm_fwd(s, x, logits, B, N, C, S, xc);
  if (dt == QV_BF16 && tokens_mma64_ok(M, N, C)) return tlm64_fwd(s, x, logits, B, N, M, C, S, xc);
  if (tokens16_ok(M, C)) return tl16_fwd(s, dt, x, logits, B, N, C, S, xc);
  QV_CHECK(M % 16 == 0, "token_learner: M=%d must be a multiple of 16", M);
  const size_t smem = (size_t)N * M * sizeof(float);
  const int grid = min(B, qv_num_sms() * 8);
  if (dt == QV_F32) { QV_TRY(opt_in_smem(token_learner_fwd_kernel<float>, smem)); qv_launch(token_learner_fwd_kernel<float>, grid, 192, smem, s, x, (const float*)logits, B, N, M, C, S, xc); }
  else { QV_TRY(opt_in_smem(token_learner_fwd_kernel<bf16>, smem)); qv_launch(token_learner_fwd_kernel<bf16>, grid, 192, smem, s, x, (const bf16*)logits, B, N, M, C, S, xc); }
  QV_LAUNCH_CHECK();
  return 0;
}
int token_learner_bwd(cudaStream_t s, int dt, const float* x, const float* S, const float* dxc, int B, int N, int M,
                      int C, void* dlogits, float* dx) {
  if (B <= 0) return 0;
  if (dt == QV_BF16 && tokens_mma_ok(M, N, C)) return tlm_bwd(s, x, S, dxc, B, N, C, dlogits, dx);
  if (dt == QV_BF16 && tokens_mma64_ok(M, N, C)) return tlm64_bwd(s, x, S, dxc, B, N, M, C, dlogits, dx);
  if (tokens16_ok(M, C)) return tl16_bwd(s, dt, x, S, dxc, B, N, C, dlogits, dx);
  QV_CHECK(C <= 256, "token_learner_bwd: C=%d > 256", C);
  const size_t smem = (size_t)(2 * N * M + M * (C + 1) + M) * sizeof(float);
  const int grid = min(B, qv_num_sms() * 4);
  if (dt == QV_F32) { QV_TRY(opt_in_smem(token_learner_bwd_kernel<float>, smem)); qv_launch(token_learner_bwd_kernel<float>, grid, 256, smem, s, x, S, dxc, B, N, M, C, (float*)dlogits, dx); }
  else { QV_TRY(opt_in_smem(token_learner_bwd_kernel<bf16>, smem)); qv_launch(token_learner_bwd_kernel<bf16>, grid, 256, smem, s, x, S, dxc, B, N, M, C, (bf16*)dlogits, dx); }
  QV_LAUNCH_CHECK();
  return 0;
}
int token_upmix_fwd(cudaStream_t s, int dt, const float* xc, int B, int M, int N, int C, const float* W, const float* bias, float* up) {
  if (B <= 0) return 0;
  if (dt == QV_BF16 && tokens_mma_ok(M, N, C)) return upm_fwd(s, xc, B, N, C, W, bias, up);
  if (dt == QV_BF16 && tokens_mma64_ok(M, N, C)) return upm64_fwd(s, xc, B, N, M, C, W, bias, up);
  if (tokens16_ok(M, C)) return up16_fwd(s, xc, B, N, C, W, bias, up);
  const size_t smem = (size_t)(N * M + M * C) * sizeof(float);
  QV_TRY(opt_in_smem(token_upmix_fwd_kernel, smem));
  qv_launch(token_upmix_fwd_kernel, min(B, qv_num_sms() * 4), 192, smem, s, xc, B, M, N, C, W, bias, up);
  QV_LAUNCH_CHECK();
  return 0;
}
int token_upmix_bwd(cudaStream_t s, int dt, const float* xc, const float* dup, int B, int M, int N, int C, const float* W,
                    float* dxc, float* dW, float* dbias) {
  if (B <= 0) return 0;
  if (dt == QV_BF16 && tokens_mma_ok(M, N, C)) return upm_bwd(s, xc, dup, B, N, C, W, dxc, dW, dbias);
  if (dt == QV_BF16 && tokens_mma64_ok(M, N, C)) return upm64_bwd(s, xc, dup, B, N, M, C, W, dxc, dW, dbias);
  if (tokens16_ok(M, C)) return up16_bwd(s, xc, dup, B, N, C, W, dxc, dW, dbias);
  QV_CHECK(N % 16 == 0, "token_upmix_bwd: N=%d must be a multiple of 16", N);
  const size_t smem = (size_t)(N * M + M * (C + 1) + 16 * (C + 1)) * sizeof(float);
  QV_TRY(opt_in_smem(token_upmix_bwd_kernel, smem));
  qv_launch(token_upmix_bwd_kernel, min(B, qv_num_sms() * 4), 192, smem, s, xc, dup, B, M, N, C, W, dxc, dW, dbias);
  QV_LAUNCH_CHECK();
  return 0;
}
