// CCF-FFN mid-section (H:704-709, QAViTv2.py:870-882) for blocks of 16 tokens (4 x 4 map), bf16 runs:
//
//   fwd : h_pre[R, C] -> GELU -> LayerNorm(dwconv_norm) -> depthwise 3x3 (+ bias) * scale -> LayerNorm(post_dwconv_norm) -> hn2[R, C]
//   bwd : d_hn2 -> d_hpre, and the gradients of both LayerNorms, the nine taps, the conv bias and the channel scale
//
// one launch per direction in place of ln_fwd + dw_fwd + ln_fwd (76 us at B = 4736, C = 96) and ln_bwd + dw_wgrad + dw_fwd(flip) +
// ln_bwd (226 us): those exchanged four [R, C] bf16 tensors (14.5 MB each) through HBM and were bound by launch latency, not bytes.
//
// Thread = one channel of one image: the channel's 4 x 4 map sits in 16 registers, so the stencil and its transpose are fully
// unrolled register code with the out-of-range taps pruned at compile time, and the parameter gradients accumulate in registers
// over the CTA's images (one atomic flush at the end).  LayerNorm needs sums over the C channels of each token: each warp
// reduces its 32 channels for all 16 tokens at once by recursive halving (16 shuffles: lanes trade halves of the token array, so
// lane 2t ends with token t's total), the C / 32 warps of an image meet through shared memory.  Everything between the bf16 input
// and the bf16 output is fp32 (the unfused chain rounded hn and cs to bf16 on the way); backward recomputes the forward from h_pre
// and the saved (mean, rstd) pairs, so hn and cs are never written.
#include "kernels.h"

namespace {

constexpr int MT = 16;   // tokens of an image: 4 x 4

struct FfnMidP {
  const bf16* h_pre; const bf16* d_hn2;
  const float *g1, *b1, *w, *bias, *scale, *g2, *b2;
  float *st1, *st2;                     // (mean, rstd) per row of the two LayerNorms: written by forward, read by backward
  bf16* hn2; bf16* d_hpre;
  float *dg1, *db1, *dw, *dbias, *dscale, *dg2, *db2;
  int B; float eps;
};

// the NW warps of ONE image meet on their own named barrier (ids 1 .. IPC): the images of a CTA never wait for each other
__device__ __forceinline__ void img_sync(int img, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(img + 1), "r"(nthreads) : "memory");
}
#ifndef FMID_FAST_GELU
#define FMID_FAST_GELU 0   // 1: tanh-form GELU on the MUFU unit (the GEMM epilogues' flavour) instead of exact erf -- A/B knob
#endif
__device__ __forceinline__ float mid_gelu(float x) { return FMID_FAST_GELU ? gelu_fast_f(x) : gelu_f(x); }
// value and derivative from one erf
__device__ __forceinline__ void mid_gelu_both(float x, float& g, float& dg) {
  if (FMID_FAST_GELU) { gelu_both_fast_f(x, g, dg); return; }
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  g = x * cdf;
  dg = fmaf(x, 0.39894228040143268f * __expf(-0.5f * x * x), cdf);
}

// v[t] summed over the 32 lanes for all 16 tokens: lane l returns the total of token l >> 1 (v is consumed)
__device__ __forceinline__ float seg16(float (&v)[MT], int lane) {
#pragma unroll
  for (int h = 8, bit = 16; h >= 1; h >>= 1, bit >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = up ? v[i] : v[i + h], keep = up ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// per-token totals of two quantities over the NW warps of an image: red[stat][warp][token] -> a[t], b[t]
template <int NW>
__device__ __forceinline__ void totals(const float* red, float (&a)[MT], float (&b)[MT]) {
#pragma unroll
  for (int q = 0; q < MT / 4; ++q) {
    float4 sa = *reinterpret_cast<const float4*>(red + q * 4), sb = *reinterpret_cast<const float4*>(red + NW * MT + q * 4);
#pragma unroll
    for (int w = 1; w < NW; ++w) {
      const float4 ta = *reinterpret_cast<const float4*>(red + w * MT + q * 4);
      const float4 tb = *reinterpret_cast<const float4*>(red + (NW + w) * MT + q * 4);
      sa.x += ta.x; sa.y += ta.y; sa.z += ta.z; sa.w += ta.w;
      sb.x += tb.x; sb.y += tb.y; sb.z += tb.z; sb.w += tb.w;
    }
    a[4 * q] = sa.x; a[4 * q + 1] = sa.y; a[4 * q + 2] = sa.z; a[4 * q + 3] = sa.w;
    b[4 * q] = sb.x; b[4 * q + 1] = sb.y; b[4 * q + 2] = sb.z; b[4 * q + 3] = sb.w;
  }
}

// LayerNorm statistics from the per-warp partial sums: lane l derives (mean, rstd) of token l & 15 once, 32 shuffles hand them to
// every lane (each thread deriving all 16 pairs itself cost ~10 instructions per token, a fifth of the forward kernel)
template <int NW>
__device__ __forceinline__ void ln_stats(const float* red, int lane, float invC, float eps, float (&m)[MT], float (&r)[MT],
                                         float& my_m, float& my_r) {
  const int t = lane & (MT - 1);
  float a = 0.f, q = 0.f;
#pragma unroll
  for (int w = 0; w < NW; ++w) { a += red[w * MT + t]; q += red[(NW + w) * MT + t]; }
  my_m = a * invC;
  my_r = rsqrtf(fmaxf(q * invC - my_m * my_m, 0.f) + eps);
#pragma unroll
  for (int i = 0; i < MT; ++i) { m[i] = __shfl_sync(0xffffffffu, my_m, i); r[i] = __shfl_sync(0xffffffffu, my_r, i); }
}

// y[p] = bias + sum_taps w[ky][kx] x[py + ky - 1][px + kx - 1] on the 4 x 4 map (cross-correlation, zero padding)
__device__ __forceinline__ void conv3(const float (&x)[MT], const float (&w)[9], float bias, float (&y)[MT]) {
#pragma unroll
  for (int py = 0; py < 4; ++py)
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      float a = bias;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int yy = py + ky - 1, xx = px + kx - 1;
          if (yy >= 0 && yy < 4 && xx >= 0 && xx < 4) a = fmaf(w[ky * 3 + kx], x[yy * 4 + xx], a);
        }
      y[py * 4 + px] = a;
    }
}

template <int NW, int IPC>
__global__ void __launch_bounds__(32 * NW * IPC) ffn_mid_fwd_kernel(FfnMidP p) {
  QV_PDL_ENTRY();
  constexpr int C = 32 * NW;
  __shared__ __align__(16) float red[2][IPC][2 * NW * MT];
  const int img = threadIdx.x / C, c = threadIdx.x % C, wi = c >> 5, lane = threadIdx.x & 31;
  float w[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) w[k] = p.w[c * 9 + k];
  const float g1 = p.g1[c], b1 = p.b1[c], g2 = p.g2[c], b2 = p.b2[c], sc = p.scale ? p.scale[c] : 1.f, bs = p.bias ? p.bias[c] : 0.f;
  const float invC = 1.f / C;
  const long ngroups = ((long)p.B + IPC - 1) / IPC;
  for (long g = blockIdx.x; g < ngroups; g += gridDim.x) {
    const long b = g * IPC + img;
    const bool act = b < p.B;
    float x[MT], cs[MT];
    if (act) {
      const bf16* src = p.h_pre + b * MT * C + c;
#pragma unroll
      for (int t = 0; t < MT; ++t) x[t] = __bfloat162float(src[t * C]);
#pragma unroll
      for (int t = 0; t < MT; ++t) x[t] = mid_gelu(x[t]);
    } else {
#pragma unroll
      for (int t = 0; t < MT; ++t) x[t] = 0.f;
    }
    {   // LayerNorm 1 over the channels of each token
      float s[MT], q[MT];
#pragma unroll
      for (int t = 0; t < MT; ++t) { s[t] = x[t]; q[t] = x[t] * x[t]; }
      const float ts = seg16(s, lane), tq = seg16(q, lane);
      if (!(lane & 1)) { red[0][img][wi * MT + (lane >> 1)] = ts; red[0][img][(NW + wi) * MT + (lane >> 1)] = tq; }
      img_sync(img, C);
      float mm, mr;
      ln_stats<NW>(red[0][img], lane, invC, p.eps, s, q, mm, mr);      // s <- mean, q <- rstd
      if (act && c < MT) *reinterpret_cast<float2*>(p.st1 + 2 * (b * MT + c)) = make_float2(mm, mr);
#pragma unroll
      for (int t = 0; t < MT; ++t) x[t] = fmaf((x[t] - s[t]) * q[t], g1, b1);
    }
    conv3(x, w, bs, cs);
#pragma unroll
    for (int t = 0; t < MT; ++t) cs[t] *= sc;
    {   // LayerNorm 2
      float s[MT], q[MT];
#pragma unroll
      for (int t = 0; t < MT; ++t) { s[t] = cs[t]; q[t] = cs[t] * cs[t]; }
      const float ts = seg16(s, lane), tq = seg16(q, lane);
      if (!(lane & 1)) { red[1][img][wi * MT + (lane >> 1)] = ts; red[1][img][(NW + wi) * MT + (lane >> 1)] = tq; }
      img_sync(img, C);
      float mm, mr;
      ln_stats<NW>(red[1][img], lane, invC, p.eps, s, q, mm, mr);
      if (act && c < MT) *reinterpret_cast<float2*>(p.st2 + 2 * (b * MT + c)) = make_float2(mm, mr);
      if (act) {
        bf16* dst = p.hn2 + b * MT * C + c;
#pragma unroll
        for (int t = 0; t < MT; ++t) dst[t * C] = __float2bfloat16_rn(fmaf((cs[t] - s[t]) * q[t], g2, b2));
      }
    }
  }
}

template <int NW, int IPC>
__global__ void __launch_bounds__(32 * NW * IPC) ffn_mid_bwd_kernel(FfnMidP p) {
  QV_PDL_ENTRY();
  constexpr int C = 32 * NW;
  __shared__ __align__(16) float red[2][IPC][2 * NW * MT];
  __shared__ __align__(16) float4 st[2][IPC][MT];                 // (mean1, rstd1, mean2, rstd2) per token, double-buffered by iteration
  const int img = threadIdx.x / C, c = threadIdx.x % C, wi = c >> 5, lane = threadIdx.x & 31;
  float w[9], aw[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { w[k] = p.w[c * 9 + k]; aw[k] = 0.f; }
  const float g1 = p.g1[c], b1 = p.b1[c], g2 = p.g2[c], sc = p.scale ? p.scale[c] : 1.f, bs = p.bias ? p.bias[c] : 0.f;
  float a_g1 = 0.f, a_b1 = 0.f, a_g2 = 0.f, a_b2 = 0.f, a_sc = 0.f, a_bs = 0.f;
  const float invC = 1.f / C;
  const long ngroups = ((long)p.B + IPC - 1) / IPC;
  int par = 0;
  for (long g = blockIdx.x; g < ngroups; g += gridDim.x, par ^= 1) {
    const long b = g * IPC + img;
    const bool act = b < p.B;
    float x[MT], dy[MT], xh1[MT], cv[MT];
    if (act) {
      const bf16* src = p.h_pre + b * MT * C + c;
      const bf16* dsrc = p.d_hn2 + b * MT * C + c;
#pragma unroll
      for (int t = 0; t < MT; ++t) x[t] = __bfloat162float(src[t * C]);
#pragma unroll
      for (int t = 0; t < MT; ++t) dy[t] = __bfloat162float(dsrc[t * C]);
      if (c < MT) {
        const float2 s1 = *reinterpret_cast<const float2*>(p.st1 + 2 * (b * MT + c));
        const float2 s2 = *reinterpret_cast<const float2*>(p.st2 + 2 * (b * MT + c));
        st[par][img][c] = make_float4(s1.x, s1.y, s2.x, s2.y);
      }
    } else {
#pragma unroll
      for (int t = 0; t < MT; ++t) { x[t] = 0.f; dy[t] = 0.f; }
      if (c < MT) st[par][img][c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    img_sync(img, C);
    // ---- recompute the forward: xh1 = LN1-normalised gelu(x), cv = conv(hn) + bias, xh2 (kept in cv's place after dscale)
    {
      float hn[MT];
#pragma unroll
      for (int t = 0; t < MT; ++t) {
        const float4 s = st[par][img][t];
        float gv;
        mid_gelu_both(x[t], gv, x[t]);      // x <- gelu'(x) for the last step
        xh1[t] = (gv - s.x) * s.y;
        hn[t] = fmaf(xh1[t], g1, b1);
      }
      conv3(hn, w, bs, cv);
    }
    // ---- LayerNorm 2 backward
    float d[MT];      // dxh2 -> dcs -> dconv
    {
      float sa[MT], sb[MT];
#pragma unroll
      for (int t = 0; t < MT; ++t) {
        const float4 s = st[par][img][t];
        const float xh2 = (cv[t] * sc - s.z) * s.w;
        a_g2 = fmaf(dy[t], xh2, a_g2);
        a_b2 += dy[t];
        d[t] = dy[t] * g2;
        sa[t] = d[t];
        sb[t] = d[t] * xh2;
      }
      const float ta = seg16(sa, lane), tb = seg16(sb, lane);
      if (!(lane & 1)) { red[0][img][wi * MT + (lane >> 1)] = ta; red[0][img][(NW + wi) * MT + (lane >> 1)] = tb; }
      img_sync(img, C);
      totals<NW>(red[0][img], sa, sb);
#pragma unroll
      for (int t = 0; t < MT; ++t) {
        const float4 s = st[par][img][t];
        const float xh2 = (cv[t] * sc - s.z) * s.w;
        const float dcs = s.w * (d[t] - sa[t] * invC - xh2 * sb[t] * invC);
        a_sc = fmaf(dcs, cv[t], a_sc);
        d[t] = dcs * sc;
        a_bs += d[t];
      }
    }
    // ---- depthwise backward: taps, then d_hn = transposed stencil of dconv
    float dh[MT];
    {
      float hn[MT];
#pragma unroll
      for (int t = 0; t < MT; ++t) hn[t] = fmaf(xh1[t], g1, b1);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          float a = 0.f;
#pragma unroll
          for (int py = 0; py < 4; ++py)
#pragma unroll
            for (int px = 0; px < 4; ++px) {
              const int yy = py + ky - 1, xx = px + kx - 1;
              if (yy >= 0 && yy < 4 && xx >= 0 && xx < 4) a = fmaf(d[py * 4 + px], hn[yy * 4 + xx], a);
            }
          aw[ky * 3 + kx] += a;
        }
#pragma unroll
      for (int py = 0; py < 4; ++py)
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          float a = 0.f;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const int oy = py - ky + 1, ox = px - kx + 1;   // the output position that read input (py, px) through tap (ky, kx)
              if (oy >= 0 && oy < 4 && ox >= 0 && ox < 4) a = fmaf(w[ky * 3 + kx], d[oy * 4 + ox], a);
            }
          dh[py * 4 + px] = a;
        }
    }
    // ---- LayerNorm 1 backward and gelu'
    {
      float sa[MT], sb[MT];
#pragma unroll
      for (int t = 0; t < MT; ++t) {
        a_g1 = fmaf(dh[t], xh1[t], a_g1);
        a_b1 += dh[t];
        dh[t] *= g1;
        sa[t] = dh[t];
        sb[t] = dh[t] * xh1[t];
      }
      const float ta = seg16(sa, lane), tb = seg16(sb, lane);
      if (!(lane & 1)) { red[1][img][wi * MT + (lane >> 1)] = ta; red[1][img][(NW + wi) * MT + (lane >> 1)] = tb; }
      img_sync(img, C);
      totals<NW>(red[1][img], sa, sb);
      if (act) {
        bf16* dst = p.d_hpre + b * MT * C + c;
#pragma unroll
        for (int t = 0; t < MT; ++t) {
          const float r1 = st[par][img][t].y;
          const float dg = r1 * (dh[t] - sa[t] * invC - xh1[t] * sb[t] * invC);
          dst[t * C] = __float2bfloat16_rn(dg * x[t]);
        }
      }
    }
  }
  // the CTA's images are summed in shared memory first: one atomic per CTA and value (the per-thread flush was 852 k atomics onto
  // 1440 addresses at B = 4736 -- a third of the kernel by ncu's stall samples, all of it in the exit drain)
  __shared__ float accs[IPC][15][C];
  accs[img][0][c] = a_g1; accs[img][1][c] = a_b1; accs[img][2][c] = a_g2; accs[img][3][c] = a_b2; accs[img][4][c] = a_bs; accs[img][5][c] = a_sc;
#pragma unroll
  for (int k = 0; k < 9; ++k) accs[img][6 + k][c] = aw[k];
  __syncthreads();
  for (int i = threadIdx.x; i < 15 * C; i += 32 * NW * IPC) {
    const int k = i / C, cc = i % C;
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < IPC; ++j) v += accs[j][k][cc];
    float* dst = k == 0 ? p.dg1 + cc : k == 1 ? p.db1 + cc : k == 2 ? p.dg2 + cc : k == 3 ? p.db2 + cc
               : k == 4 ? (p.dbias ? p.dbias + cc : nullptr) : k == 5 ? (p.dscale ? p.dscale + cc : nullptr) : p.dw + cc * 9 + (k - 6);
    if (dst) atomicAdd(dst, v);
  }
}

// ------------------------------------------------------------------------------------------------ 8 x 8 maps (64 tokens)
// MEASURED SLOWER than the ln_fwd + dw_fwd + ln_fwd chain it would replace (266 + 612 us against 177 + 362 us per block at B = 4736,
// C = 96: eight serial token iterations per warp with two shuffle reductions each, erf per element, nine shared loads per stencil
// output), so block.cu routes 8 x 8 maps here only with QV_FMID_TILE=1; kept with its parity test as the starting point for a
// register-tiled version.
// Same operation for 64-token blocks (QAViTv2, HQAViT-TinyImageNet): a channel's map no longer fits a thread's registers, so the
// image lives in shared memory as fp32 [token][channel] tiles and every phase is "a warp takes a token, its lanes stride over the
// channels": conflict-free shared memory, coalesced global rows, LayerNorm sums by one warp reduction per token, and the stencil reads
// its neighbours from the tile.  One CTA per image (8 warps), persistent over images; parameters staged in shared memory once; per-lane
// channel accumulators for the parameter gradients, reduced over the CTA's warps in shared memory before the atomics.
constexpr int TW = 8;                      // warps per CTA
constexpr int TCH = 4;                     // channels per lane: C <= 128

struct TilePar {                           // parameters in shared memory: [g1 | b1 | g2 | b2 | scale | bias | w (9 C)]
  const float *g1, *b1, *g2, *b2, *sc, *bs, *w;
  __device__ TilePar(const float* base, int C) : g1(base), b1(base + C), g2(base + 2 * C), b2(base + 3 * C), sc(base + 4 * C), bs(base + 5 * C), w(base + 6 * C) {}
};
__device__ __forceinline__ void stage_params(float* base, const FfnMidP& p, int C) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    base[c] = p.g1[c]; base[C + c] = p.b1 ? p.b1[c] : 0.f; base[2 * C + c] = p.g2[c]; base[3 * C + c] = p.b2 ? p.b2[c] : 0.f;
    base[4 * C + c] = p.scale ? p.scale[c] : 1.f; base[5 * C + c] = p.bias ? p.bias[c] : 0.f;
  }
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) base[6 * C + i] = p.w[i];
}

template <int SIDE>
__global__ void __launch_bounds__(TW * 32) ffn_mid_tile_fwd_kernel(FfnMidP p, int C) {
  QV_PDL_ENTRY();
  constexpr int NT = SIDE * SIDE;
  extern __shared__ __align__(16) float sm[];
  float* A = sm;                            // gelu(x), then hn
  float* Bt = A + NT * C;                   // the raw bf16 image first, then cs
  float* par = Bt + NT * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_params(par, p, C);
  const TilePar P(par, C);
  const float invC = 1.f / C;
  for (long b = blockIdx.x; b < p.B; b += gridDim.x) {
    __syncthreads();                         // parameters staged / previous image's tiles consumed
    // the image comes in as ONE wave of independent 16 B loads (a global load per token iteration cost a memory round trip each)
    const bf16* src = reinterpret_cast<const bf16*>(Bt);
    for (int i = threadIdx.x; i < NT * C / 8; i += TW * 32)
      reinterpret_cast<uint4*>(Bt)[i] = reinterpret_cast<const uint4*>(p.h_pre + b * NT * C)[i];
    __syncthreads();
    for (int t = warp; t < NT; t += TW) {    // gelu + LayerNorm 1 statistics
      float s = 0.f, q = 0.f;
      for (int c = lane; c < C; c += 32) {
        const float v = mid_gelu(__bfloat162float(src[t * C + c]));
        A[t * C + c] = v; s += v; q = fmaf(v, v, q);
      }
      s = warp_sum(s); q = warp_sum(q);
      const float m = s * invC, r = rsqrtf(fmaxf(q * invC - m * m, 0.f) + p.eps);
      for (int c = lane; c < C; c += 32) A[t * C + c] = fmaf((A[t * C + c] - m) * r, P.g1[c], P.b1[c]);   // hn (own elements)
      if (lane == 0) *reinterpret_cast<float2*>(p.st1 + 2 * (b * NT + t)) = make_float2(m, r);
    }
    __syncthreads();
    for (int t = warp; t < NT; t += TW) {    // stencil * scale + LayerNorm 2 statistics
      const int py = t / SIDE, px = t % SIDE;
      float s = 0.f, q = 0.f;
      for (int c = lane; c < C; c += 32) {
        float a = P.bs[c];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int yy = py + ky - 1, xx = px + kx - 1;
            if (yy >= 0 && yy < SIDE && xx >= 0 && xx < SIDE) a = fmaf(P.w[c * 9 + ky * 3 + kx], A[(yy * SIDE + xx) * C + c], a);
          }
        a *= P.sc[c];
        Bt[t * C + c] = a; s += a; q = fmaf(a, a, q);
      }
      s = warp_sum(s); q = warp_sum(q);
      const float m = s * invC, r = rsqrtf(fmaxf(q * invC - m * m, 0.f) + p.eps);
      bf16* dst = p.hn2 + (b * NT + t) * C;
      for (int c = lane; c < C; c += 32) dst[c] = __float2bfloat16_rn(fmaf((Bt[t * C + c] - m) * r, P.g2[c], P.b2[c]));
      if (lane == 0) *reinterpret_cast<float2*>(p.st2 + 2 * (b * NT + t)) = make_float2(m, r);
    }
  }
}

template <int SIDE>
__global__ void __launch_bounds__(TW * 32) ffn_mid_tile_bwd_kernel(FfnMidP p, int C) {
  QV_PDL_ENTRY();
  constexpr int NT = SIDE * SIDE;
  extern __shared__ __align__(16) float sm[];
  float* A = sm;                            // xh1 (LayerNorm-1-normalised gelu(x))
  float* Bt = A + NT * C;                   // the raw bf16 dy first, then cv = conv(hn) + bias
  float* Dt = Bt + NT * C;                  // dy, then dconv
  float* par = Dt + NT * C;
  float4* sst = reinterpret_cast<float4*>(par + 15 * C + (4 - (15 * C) % 4) % 4);   // (mean1, rstd1, mean2, rstd2) per token
  bf16* Xr = reinterpret_cast<bf16*>(sst + NT);                                     // the raw bf16 x of the image (gelu' at the end)
  float* red;                               // [TW][15][C] only at the very end: overlays the tiles (see below)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_params(par, p, C);
  const TilePar P(par, C);
  const float invC = 1.f / C;
  // per-lane accumulators for channels lane, lane + 32, ...: dg1, db1, dg2, db2, dbias, dscale, 9 taps
  float acc[15][TCH];
#pragma unroll
  for (int k = 0; k < 15; ++k)
#pragma unroll
    for (int j = 0; j < TCH; ++j) acc[k][j] = 0.f;
  for (long b = blockIdx.x; b < p.B; b += gridDim.x) {
    __syncthreads();
    // x, dy and the statistics come in as ONE wave of independent loads (per-token global loads cost a memory round trip each)
    const bf16* src = Xr;
    const bf16* dsrc = reinterpret_cast<const bf16*>(Bt);
    for (int i = threadIdx.x; i < NT * C / 8; i += TW * 32) {
      reinterpret_cast<uint4*>(Xr)[i] = reinterpret_cast<const uint4*>(p.h_pre + b * NT * C)[i];
      reinterpret_cast<uint4*>(Bt)[i] = reinterpret_cast<const uint4*>(p.d_hn2 + b * NT * C)[i];
    }
    for (int t = threadIdx.x; t < NT; t += TW * 32) {
      const float2 a1 = *reinterpret_cast<const float2*>(p.st1 + 2 * (b * NT + t)), a2 = *reinterpret_cast<const float2*>(p.st2 + 2 * (b * NT + t));
      sst[t] = make_float4(a1.x, a1.y, a2.x, a2.y);
    }
    __syncthreads();
    for (int t = warp; t < NT; t += TW) {    // xh1, dy
      const float4 s4 = sst[t];
      const float2 s1 = make_float2(s4.x, s4.y);
      for (int c = lane; c < C; c += 32) {
        A[t * C + c] = (mid_gelu(__bfloat162float(src[t * C + c])) - s1.x) * s1.y;
        Dt[t * C + c] = __bfloat162float(dsrc[t * C + c]);
      }
    }
    __syncthreads();
    for (int t = warp; t < NT; t += TW) {    // cv = stencil(hn) + bias
      const int py = t / SIDE, px = t % SIDE;
      for (int c = lane; c < C; c += 32) {
        float a = P.bs[c];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int yy = py + ky - 1, xx = px + kx - 1;
            if (yy >= 0 && yy < SIDE && xx >= 0 && xx < SIDE)
              a = fmaf(P.w[c * 9 + ky * 3 + kx], fmaf(A[(yy * SIDE + xx) * C + c], P.g1[c], P.b1[c]), a);
          }
        Bt[t * C + c] = a;
      }
    }
    __syncthreads();
    for (int t = warp; t < NT; t += TW) {    // LayerNorm 2 backward: Dt <- dconv
      const float2 s2 = make_float2(sst[t].z, sst[t].w);
      float sa = 0.f, sb = 0.f;
#pragma unroll
      for (int j = 0; j < TCH; ++j) {
        const int c = lane + 32 * j;
        if (c < C) {
          const float dy = Dt[t * C + c], xh2 = (Bt[t * C + c] * P.sc[c] - s2.x) * s2.y, d = dy * P.g2[c];
          acc[2][j] = fmaf(dy, xh2, acc[2][j]); acc[3][j] += dy;
          sa += d; sb = fmaf(d, xh2, sb);
        }
      }
      sa = warp_sum(sa) * invC; sb = warp_sum(sb) * invC;
#pragma unroll
      for (int j = 0; j < TCH; ++j) {
        const int c = lane + 32 * j;
        if (c < C) {
          const float cv = Bt[t * C + c], xh2 = (cv * P.sc[c] - s2.x) * s2.y;
          const float dcs = s2.y * (Dt[t * C + c] * P.g2[c] - sa - xh2 * sb);
          acc[5][j] = fmaf(dcs, cv, acc[5][j]);
          const float dc = dcs * P.sc[c];
          acc[4][j] += dc;
          Dt[t * C + c] = dc;
        }
      }
    }
    __syncthreads();
    for (int t = warp; t < NT; t += TW) {    // stencil backward (taps + transposed stencil), LayerNorm 1 backward, gelu'
      const int py = t / SIDE, px = t % SIDE;
      const float2 s1 = make_float2(sst[t].x, sst[t].y);
      float dh[TCH], sa = 0.f, sb = 0.f;
#pragma unroll
      for (int j = 0; j < TCH; ++j) {
        const int c = lane + 32 * j;
        dh[j] = 0.f;
        if (c < C) {
          const float dc = Dt[t * C + c];
          float a = 0.f;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const int yy = py + ky - 1, xx = px + kx - 1;     // input position this output read through tap (ky, kx)
              if (yy >= 0 && yy < SIDE && xx >= 0 && xx < SIDE)
                acc[6 + ky * 3 + kx][j] = fmaf(dc, fmaf(A[(yy * SIDE + xx) * C + c], P.g1[c], P.b1[c]), acc[6 + ky * 3 + kx][j]);
              const int oy = py - ky + 1, ox = px - kx + 1;     // output position that read this input through tap (ky, kx)
              if (oy >= 0 && oy < SIDE && ox >= 0 && ox < SIDE) a = fmaf(P.w[c * 9 + ky * 3 + kx], Dt[(oy * SIDE + ox) * C + c], a);
            }
          const float xh1 = A[t * C + c];
          acc[0][j] = fmaf(a, xh1, acc[0][j]); acc[1][j] += a;
          dh[j] = a * P.g1[c];
          sa += dh[j]; sb = fmaf(dh[j], xh1, sb);
        }
      }
      sa = warp_sum(sa) * invC; sb = warp_sum(sb) * invC;
      bf16* dst = p.d_hpre + (b * NT + t) * C;
#pragma unroll
      for (int j = 0; j < TCH; ++j) {
        const int c = lane + 32 * j;
        if (c < C) {
          float gv, dg;
          mid_gelu_both(__bfloat162float(src[t * C + c]), gv, dg);
          dst[c] = __float2bfloat16_rn(s1.y * (dh[j] - sa - A[t * C + c] * sb) * dg);
        }
      }
    }
  }
  // ---- the warps' channel accumulators meet in shared memory (over the tiles, which are dead now), one atomic per value and CTA
  __syncthreads();
  red = sm;
#pragma unroll
  for (int k = 0; k < 15; ++k)
#pragma unroll
    for (int j = 0; j < TCH; ++j) {
      const int c = lane + 32 * j;
      if (c < C) red[(warp * 15 + k) * C + c] = acc[k][j];
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 15 * C; i += TW * 32) {
    const int k = i / C, c = i % C;
    float v = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < TW; ++w2) v += red[(w2 * 15 + k) * C + c];
    float* dst = k == 0 ? p.dg1 + c : k == 1 ? p.db1 + c : k == 2 ? p.dg2 + c : k == 3 ? p.db2 + c
               : k == 4 ? (p.dbias ? p.dbias + c : nullptr) : k == 5 ? (p.dscale ? p.dscale + c : nullptr) : p.dw + c * 9 + (k - 6);
    if (dst) atomicAdd(dst, v);
  }
}

template <int SIDE>
int launch_tile(cudaStream_t s, const FfnMidP& p, int C, bool bwd) {
  constexpr int NT = SIDE * SIDE;
  const size_t fwd_b = ((size_t)2 * NT * C + 15 * C) * 4, bwd_b = ((size_t)3 * NT * C + 15 * C + 4 + 4 * NT) * 4 + (size_t)NT * C * 2;
  const size_t smem = bwd ? max(bwd_b, (size_t)TW * 15 * C * 4) : fwd_b;
  QV_CHECK(smem <= 200 * 1024, "ffn_mid: %zu B of shared memory", smem);
  auto kf = ffn_mid_tile_fwd_kernel<SIDE>;
  auto kb = ffn_mid_tile_bwd_kernel<SIDE>;
  if (smem > 48 * 1024) {
    if (bwd) QV_CUDA(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else QV_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int per_sm = max(1, min(8, (int)(220 * 1024 / (smem + 1024))));
  const int grid = (int)max(1L, min((long)p.B, (long)qv_num_sms() * per_sm));
  if (bwd) qv_launch(kb, grid, TW * 32, smem, s, p, C);
  else qv_launch(kf, grid, TW * 32, smem, s, p, C);
  QV_LAUNCH_CHECK();
  return 0;
}

template <typename K>
int grid_for(K kernel, int threads, long ngroups) {
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  return (int)max(1L, min(ngroups, (long)qv_num_sms() * per_sm));
}

template <int NW, int IPC>
int launch(cudaStream_t s, const FfnMidP& p, bool bwd) {
  const long ngroups = ((long)p.B + IPC - 1) / IPC;
  if (bwd) qv_launch(ffn_mid_bwd_kernel<NW, IPC>, grid_for(ffn_mid_bwd_kernel<NW, IPC>, 32 * NW * IPC, ngroups), 32 * NW * IPC, 0, s, p);
  else qv_launch(ffn_mid_fwd_kernel<NW, IPC>, grid_for(ffn_mid_fwd_kernel<NW, IPC>, 32 * NW * IPC, ngroups), 32 * NW * IPC, 0, s, p);
  QV_LAUNCH_CHECK();
  return 0;
}
int dispatch(cudaStream_t s, const FfnMidP& p, int C, bool bwd) {
  switch (C / 32) {
    case 1: return launch<1, 8>(s, p, bwd);
    case 2: return launch<2, 4>(s, p, bwd);
    case 3: return launch<3, 4>(s, p, bwd);
    default: return launch<4, 2>(s, p, bwd);
  }
}

}  // namespace

bool ffn_mid_ok(int side, int C) { return (side == 4 && C % 32 == 0 && C >= 32 && C <= 128) || (side == 8 && C % 8 == 0 && C >= 8 && C <= 128); }

int ffn_mid_fwd(cudaStream_t s, const void* h_pre, int B, int side, int C, const float* g1, const float* b1, const float* w, const float* bias,
                const float* scale, const float* g2, const float* b2, float eps, void* hn2, float* stats1, float* stats2) {
  if (B <= 0) return 0;
  QV_CHECK(ffn_mid_ok(side, C), "ffn_mid: side=%d C=%d not covered (4 x 4 maps: C a multiple of 32 in [32, 128]; 8 x 8 maps: C a multiple of 8 <= 128)", side, C);
  FfnMidP p{};
  p.h_pre = static_cast<const bf16*>(h_pre); p.g1 = g1; p.b1 = b1; p.w = w; p.bias = bias; p.scale = scale; p.g2 = g2; p.b2 = b2;
  p.st1 = stats1; p.st2 = stats2; p.hn2 = static_cast<bf16*>(hn2); p.B = B; p.eps = eps;
  return side == 8 ? launch_tile<8>(s, p, C, false) : dispatch(s, p, C, false);
}

int ffn_mid_bwd(cudaStream_t s, const void* h_pre, const void* d_hn2, const float* stats1, const float* stats2, int B, int side, int C,
                const float* g1, const float* b1, const float* w, const float* bias, const float* scale, const float* g2, void* d_hpre,
                float* dg1, float* db1, float* dw, float* dbias, float* dscale, float* dg2, float* db2) {
  if (B <= 0) return 0;
  QV_CHECK(ffn_mid_ok(side, C), "ffn_mid: side=%d C=%d not covered (4 x 4 maps: C a multiple of 32 in [32, 128]; 8 x 8 maps: C a multiple of 8 <= 128)", side, C);
  FfnMidP p{};
  p.h_pre = static_cast<const bf16*>(h_pre); p.d_hn2 = static_cast<const bf16*>(d_hn2); p.g1 = g1; p.b1 = b1; p.w = w; p.bias = bias;
  p.scale = scale; p.g2 = g2; p.st1 = const_cast<float*>(stats1); p.st2 = const_cast<float*>(stats2);
  p.d_hpre = static_cast<bf16*>(d_hpre); p.dg1 = dg1; p.db1 = db1; p.dw = dw; p.dbias = dbias; p.dscale = dscale; p.dg2 = dg2; p.db2 = db2;
  p.B = B;
  return side == 8 ? launch_tile<8>(s, p, C, true) : dispatch(s, p, C, true);
}
