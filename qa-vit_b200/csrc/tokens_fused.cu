// TokenLearner and TokenUpMix of HQAViT's block wrapper (H:971-1031, 1091-1123) as ONE kernel per direction each, for the
// bf16 run with 16 learned tokens, <= 64 stream tokens and 192 channels (HQAViT CIFAR-100: 64 -> 16 -> 64):
//
//   tlf_fwd : x[N, C] -> LayerNorm -> Linear C -> 16 -> softmax over tokens -> xc[16, C] = S^T x      (H:985-1002)
//   tlf_bwd : dxc -> dx = S dxc + LayerNorm-backward(dlogits W), dW / db / dgamma / dbeta of the gate
//   upf_fwd : xc'[16, C] -> Linear over the token axis 16 -> N (+ bias) -> LayerNorm(C)              (H:1016-1031)
//   upf_bwd : dout -> LayerNorm-backward -> dxc' = W^T dup, dW += dup xc'^T, dgamma / dbeta
//
// They replace ln_fwd + GEMM + tlm_fwd, tlm_bwd + 2 GEMMs + ln_bwd, upm_fwd + ln_fwd and ln_bwd + upm_bwd (12 launches per
// block -> 4) and every intermediate those exchanged through HBM (the bf16 LayerNorm output, the gate logits, the
// [B N, C] fp32 up-mixed tensor): each kernel makes one pass over the full-resolution stream tensor.
//
// Precision.  The wrapper has no residual connection around it -- the block's output REPLACES the stream -- so rounding
// here goes straight into the fp32 stream (the live reference under autocast is 5e-2 / 9e-2 away from its own fp32 run on
// logits / gradients; the gate softmax over tokens is the amplifier).  Every product runs on mma.sync.m16n8k16 with both
// operands kept as bf16 PAIRS (hi + lo, ~16 mantissa bits) and three MMAs per product (hi hi + lo hi + hi lo), accumulated
// in fp32; LayerNorm statistics, softmax and all elementwise math are fp32.  The gate logits never leave the SM.
//
// LayerNorm is folded into the gate GEMM algebraically: with Wg[m, k] = gamma_k W[m, k], c1[m] = sum_k Wg[m, k],
// c0[m] = sum_k beta_k W[m, k] + b[m]:  logits[n, m] = rstd_n (x_n . Wg_m - mean_n c1[m]) + c0[m], so the x tile is the
// only large MMA operand.  Backward uses the same identity: with dl' = dlogits * rstd, P = dl'^T x, U = dl'^T mean,
// T = colsum(dlogits):  dW = gamma (P - U) + beta T,  dgamma = colsum_m W (P - U),  dbeta = colsum_m W T,  db = T.
//
// CTA = 8 warps walking images; the NEXT image's tile is prefetched into registers while the current one is processed
// (the predecessors serialised load / convert / MMA per image and ran at 1.4 - 2.5 TB/s).
#include "kernels.h"

#ifndef TLFB_M
#define TLFB_M 2   // CTAs per SM the backward kernel is compiled for (A/B knob)
#endif

namespace {

constexpr int FC = 192;         // channels
constexpr int FXP = FC + 8;     // pitch (bf16) of [rows][C] tiles: 400 B rows, conflict-free for ldmatrix
constexpr int FSP = 24;         // pitch (bf16) of [N][16] tiles
constexpr int FM = 16;          // learned tokens
constexpr int FNT = 256, FNW = 8;
constexpr int FMAXN = 64;       // stream tokens per image (multiple of 16)

__device__ __forceinline__ uint32_t sa(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm4(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm2(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void ldsm2t(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// acc += (ah + al)(bh + bl) without the lo * lo term
__device__ __forceinline__ void mma3(float* acc, const uint32_t* ah, const uint32_t* al, uint32_t bh0, uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma16816(acc, ah, bh0, bh1);
  mma16816(acc, al, bh0, bh1);
  mma16816(acc, ah, bl0, bl1);
}
__device__ __forceinline__ uint32_t pack2(float x, float y) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void split2(float x, float y, uint32_t* hi, uint32_t* lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  const float2 hf = __bfloat1622float2(h);
  *hi = *reinterpret_cast<const uint32_t*>(&h);
  *lo = pack2(x - hf.x, y - hf.y);
}
__device__ __forceinline__ void store_split(bf16* hi, bf16* lo, int idx, float v) {
  const bf16 h = __float2bfloat16_rn(v);
  hi[idx] = h;
  lo[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
}
__device__ __forceinline__ void store_split4(bf16* hi, bf16* lo, int idx, float4 v) {
  uint32_t h0, l0, h1, l1;
  split2(v.x, v.y, &h0, &l0);
  split2(v.z, v.w, &h1, &l1);
  *reinterpret_cast<uint2*>(hi + idx) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(lo + idx) = make_uint2(l0, l1);
}
// value of element pair (hi + lo) at an even column
__device__ __forceinline__ float2 pair_val(const bf16* hi, const bf16* lo, int idx) {
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(hi + idx));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(lo + idx));
  return make_float2(a.x + b.x, a.y + b.y);
}
// A (16 x 16) from smem [m][k] (k contiguous)
__device__ __forceinline__ void ldA(uint32_t* a, const bf16* base, int pitch, int m0, int k0, int lane) {
  const int mat = lane >> 3, r = lane & 7;
  ldsm4(a, sa(base + (m0 + r + (mat & 1) * 8) * pitch + k0 + (mat >> 1) * 8));
}
// A (16 x 16) from smem [k][m] (m contiguous)
__device__ __forceinline__ void ldAt(uint32_t* a, const bf16* base, int pitch, int m0, int k0, int lane) {
  const int mat = lane >> 3, r = lane & 7;
  ldsm4t(a, sa(base + (k0 + r + (mat >> 1) * 8) * pitch + m0 + (mat & 1) * 8));
}
// B of ONE n8 tile from smem [n][k] (k contiguous)
__device__ __forceinline__ void ldB2(uint32_t* b, const bf16* base, int pitch, int n0, int k0, int lane) {
  const int mat = (lane >> 3) & 1, r = lane & 7;
  ldsm2(b, sa(base + (n0 + r) * pitch + k0 + mat * 8));
}
// B of two adjacent n8 tiles from smem [k][n] (n contiguous)
__device__ __forceinline__ void ldBt(uint32_t* b, const bf16* base, int pitch, int n0, int k0, int lane) {
  const int mat = lane >> 3, r = lane & 7;
  ldsm4t(b, sa(base + (k0 + r + (mat & 1) * 8) * pitch + n0 + (mat >> 1) * 8));
}
// B of ONE n8 tile from smem [k][n] (n contiguous)
__device__ __forceinline__ void ldBt2(uint32_t* b, const bf16* base, int pitch, int n0, int k0, int lane) {
  const int mat = (lane >> 3) & 1, r = lane & 7;
  ldsm2t(b, sa(base + (k0 + r + mat * 8) * pitch + n0));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float g_sum(float v) {      // over the 8 row groups of a fragment (lanes with equal t)
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}
__device__ __forceinline__ float sum4(float4 v) { return (v.x + v.y) + (v.z + v.w); }
__device__ __forceinline__ float sq4(float4 v, float m) {
  const float a = v.x - m, b = v.y - m, c = v.z - m, d = v.w - m;
  return (a * a + b * b) + (c * c + d * d);
}

// ---- the [N, C] fp32 stream tile of one image in registers: pass p holds rows p * 16 + 2 * warp (+1) as 3 float4 per lane
// (the 2 rows are 96 contiguous float4; lane l owns float4 l, l + 32, l + 64)
struct XRegs { float4 v[4][3]; };
__device__ __forceinline__ void xload(XRegs& r, const float* __restrict__ img, int N, int warp, int lane) {
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    if (p * 16 < N) {
      const float4* src = reinterpret_cast<const float4*>(img + (long)(p * 16 + 2 * warp) * FC);
#pragma unroll
      for (int u = 0; u < 3; ++u) r.v[p][u] = __ldg(src + lane + 32 * u);
    }
  }
}
// registers -> hi / lo bf16 tiles; STATS: also mean / rstd of every row (two-pass, in registers) -> shared (and global when st_g)
template <bool STATS>
__device__ __forceinline__ void xstore(const XRegs& r, bf16* X, bf16* XL, float* mean_s, float* rstd_s, float* st_g, int N, int warp,
                                       int lane, float eps) {
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    if (p * 16 < N) {
      const int r0 = p * 16 + 2 * warp;
      if (STATS) {
        const bool lo16 = lane < 16;
        float sA = sum4(r.v[p][0]) + (lo16 ? sum4(r.v[p][1]) : 0.f);
        float sB = sum4(r.v[p][2]) + (lo16 ? 0.f : sum4(r.v[p][1]));
        sA = warp_sum(sA); sB = warp_sum(sB);
        const float mA = sA * (1.f / FC), mB = sB * (1.f / FC);
        float qA = sq4(r.v[p][0], mA) + (lo16 ? sq4(r.v[p][1], mA) : 0.f);
        float qB = sq4(r.v[p][2], mB) + (lo16 ? 0.f : sq4(r.v[p][1], mB));
        qA = warp_sum(qA); qB = warp_sum(qB);
        if (lane == 0) {
          const float rA = rsqrtf(qA * (1.f / FC) + eps), rB = rsqrtf(qB * (1.f / FC) + eps);
          mean_s[r0] = mA; mean_s[r0 + 1] = mB; rstd_s[r0] = rA; rstd_s[r0 + 1] = rB;
          if (st_g) { *reinterpret_cast<float4*>(st_g + 2 * r0) = make_float4(mA, rA, mB, rB); }
        }
      }
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int j = lane + 32 * u;
        const int row = r0 + (j >= 48 ? 1 : 0), col = (j >= 48 ? j - 48 : j) * 4;
        store_split4(X, XL, row * FXP + col, r.v[p][u]);
      }
    }
  }
}
// the [16, C] fp32 tile (xc / dxc) in registers: 3 float4 per thread of a 256-thread CTA
struct DRegs { float4 v[3]; };
__device__ __forceinline__ void dload(DRegs& r, const float* __restrict__ img, int tid) {
#pragma unroll
  for (int u = 0; u < 3; ++u) r.v[u] = __ldg(reinterpret_cast<const float4*>(img) + tid + FNT * u);
}
__device__ __forceinline__ void dstore(const DRegs& r, bf16* D, bf16* DL, int tid) {
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    const int i = tid + FNT * u, row = i / 48, col = (i - row * 48) * 4;
    store_split4(D, DL, row * FXP + col, r.v[u]);
  }
}

struct Carve {
  uint8_t* p;
  __device__ explicit Carve(uint8_t* base) : p(base) {}
  template <typename T> __device__ T* take(size_t n) {
    T* r = reinterpret_cast<T*>(p);
    p += (n * sizeof(T) + 15) & ~(size_t)15;
    return r;
  }
};
constexpr size_t al16(size_t b) { return (b + 15) & ~(size_t)15; }

// =============================================================================================== TokenLearner forward
constexpr size_t TLF_FWD_SMEM = 2 * al16(FMAXN * FXP * 2) + 2 * al16(FM * FXP * 2) + 2 * al16(FMAXN * FSP * 2) + al16(FMAXN * FM * 4) +
                                2 * al16(16 * FM * 4) + 2 * al16(FMAXN * 4) + 2 * al16(FM * 4);

__global__ void __launch_bounds__(FNT, 2) tlf_fwd_kernel(const float* __restrict__ x, int B, int N, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, const float* __restrict__ W,
                                                         const float* __restrict__ bias, float eps, float* __restrict__ Sout,
                                                         float* __restrict__ Zout, float* __restrict__ xc) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  Carve cv(smraw);
  bf16* X = cv.take<bf16>(FMAXN * FXP); bf16* XL = cv.take<bf16>(FMAXN * FXP);
  bf16* Wg = cv.take<bf16>(FM * FXP); bf16* WgL = cv.take<bf16>(FM * FXP);
  bf16* S = cv.take<bf16>(FMAXN * FSP); bf16* SL = cv.take<bf16>(FMAXN * FSP);
  float* F = cv.take<float>(FMAXN * FM);
  float* Rmax = cv.take<float>(16 * FM); float* Rsum = cv.take<float>(16 * FM);
  float* mean_s = cv.take<float>(FMAXN); float* rstd_s = cv.take<float>(FMAXN);
  float* c0 = cv.take<float>(FM); float* c1 = cv.take<float>(FM);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int MT = N / 16;

  // ---- per-CTA constants: Wg = gamma (.) W as hi / lo, c1 = rowsum(Wg), c0 = W beta + b
  for (int i = tid; i < FM * FC; i += FNT) {
    const int m = i / FC, k = i - m * FC;
    store_split(Wg, WgL, m * FXP + k, W[i] * gamma[k]);
  }
  for (int m = warp; m < FM; m += FNW) {
    float a1 = 0.f, a0 = 0.f;
    for (int k = lane; k < FC; k += 32) { const float w = W[m * FC + k]; a1 = fmaf(w, gamma[k], a1); a0 = fmaf(w, beta[k], a0); }
    a1 = warp_sum(a1); a0 = warp_sum(a0);
    if (lane == 0) { c1[m] = a1; c0[m] = a0 + bias[m]; }
  }
  XRegs xr;
  if ((int)blockIdx.x < B) xload(xr, x + (long)blockIdx.x * N * FC, N, warp, lane);
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    xstore<true>(xr, X, XL, mean_s, rstd_s, nullptr, N, warp, lane, eps);
    if (b + (int)gridDim.x < B) xload(xr, x + (long)(b + gridDim.x) * N * FC, N, warp, lane);
    __syncthreads();
    // ---- gate logits [N, 16]: warp = (token tile, slot half)
    {
      const int mt = warp & 3, nt = warp >> 2;
      if (mt < MT) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int ks = 0; ks < FC / 16; ++ks) {
          uint32_t a[4], al[4], bh[2], bl[2];
          ldA(a, X, FXP, mt * 16, ks * 16, lane);
          ldA(al, XL, FXP, mt * 16, ks * 16, lane);
          ldB2(bh, Wg, FXP, nt * 8, ks * 16, lane);
          ldB2(bl, WgL, FXP, nt * 8, ks * 16, lane);
          mma16816(acc, a, bh[0], bh[1]);       // three independent chains (one accumulator = 36 dependent MMAs per image)
          mma16816(acc1, al, bh[0], bh[1]);
          mma16816(acc2, a, bl[0], bl[1]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] += acc1[j] + acc2[j];
        const int n0 = mt * 16 + g, n1 = n0 + 8, m = nt * 8 + 2 * t;
        const float r0 = rstd_s[n0], u0 = mean_s[n0], r1 = rstd_s[n1], u1 = mean_s[n1];
        // zc = LN-normalised projection without the constant c0: kept for backward, where sum_k g xhat = sum_m dlogits zc
        const float z00 = r0 * (acc[0] - u0 * c1[m]), z01 = r0 * (acc[1] - u0 * c1[m + 1]);
        const float z10 = r1 * (acc[2] - u1 * c1[m]), z11 = r1 * (acc[3] - u1 * c1[m + 1]);
        *reinterpret_cast<float2*>(Zout + ((long)b * N + n0) * FM + m) = make_float2(z00, z01);
        *reinterpret_cast<float2*>(Zout + ((long)b * N + n1) * FM + m) = make_float2(z10, z11);
        F[n0 * FM + m] = z00 + c0[m];
        F[n0 * FM + m + 1] = z01 + c0[m + 1];
        F[n1 * FM + m] = z10 + c0[m];
        F[n1 * FM + m + 1] = z11 + c0[m + 1];
      }
    }
    __syncthreads();
    // ---- softmax over the token axis, per slot: 16 threads per slot
    {
      const int col = tid & 15, part = tid >> 4;
      float mx = -INFINITY;
      for (int n = part; n < N; n += 16) mx = fmaxf(mx, F[n * FM + col]);
      Rmax[part * FM + col] = mx;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 16; ++k) mx = fmaxf(mx, Rmax[k * FM + col]);
      float z = 0.f;
      float e[FMAXN / 16];
#pragma unroll
      for (int q = 0; q < FMAXN / 16; ++q) {
        const int n = part + 16 * q;
        e[q] = n < N ? __expf(F[n * FM + col] - mx) : 0.f;
        z += e[q];
      }
      Rsum[part * FM + col] = z;
      __syncthreads();
      z = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) z += Rsum[k * FM + col];
      z = 1.f / z;
#pragma unroll
      for (int q = 0; q < FMAXN / 16; ++q) {
        const int n = part + 16 * q;
        if (n < N) {
          const float s = e[q] * z;
          Sout[((long)b * N + n) * FM + col] = s;
          store_split(S, SL, n * FSP + col, s);
        }
      }
    }
    __syncthreads();
    // ---- xc[16, C] = S^T x: warp = 3 channel tiles of 8
    {
      float acc[3][4];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      for (int ks = 0; ks < MT; ++ks) {
        uint32_t a[4], al[4];
        ldAt(a, S, FSP, 0, ks * 16, lane);
        ldAt(al, SL, FSP, 0, ks * 16, lane);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          uint32_t bh[2], bl[2];
          ldBt2(bh, X, FXP, (warp * 3 + i) * 8, ks * 16, lane);
          ldBt2(bl, XL, FXP, (warp * 3 + i) * 8, ks * 16, lane);
          mma3(acc[i], a, al, bh[0], bh[1], bl[0], bl[1]);
        }
      }
      float* o = xc + (long)b * FM * FC + 2 * t;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int c = (warp * 3 + i) * 8;
        *reinterpret_cast<float2*>(o + (long)g * FC + c) = make_float2(acc[i][0], acc[i][1]);
        *reinterpret_cast<float2*>(o + (long)(g + 8) * FC + c) = make_float2(acc[i][2], acc[i][3]);
      }
    }
    __syncthreads();
  }
}

// =============================================================================================== TokenLearner backward
constexpr size_t TLF_BWD_SMEM = 2 * al16(FMAXN * FXP * 2) + 4 * al16(FM * FXP * 2) + 6 * al16(FMAXN * FSP * 2) + 2 * al16(FMAXN * FM * 4) +
                                2 * al16(FMAXN * 4) + 4 * al16(FM * 4) + al16(2 * FMAXN * 2 * 4) + 2 * al16(FC * 4);

__global__ void __launch_bounds__(FNT, TLFB_M) tlf_bwd_kernel(const float* __restrict__ x, const float* __restrict__ Sin,
                                                         const float* __restrict__ Zin, const float* __restrict__ dxc, int B, int N,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         const float* __restrict__ W, float eps, float* __restrict__ dx,
                                                         float* __restrict__ dW, float* __restrict__ dbias,
                                                         float* __restrict__ dgamma, float* __restrict__ dbeta) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  Carve cv(smraw);
  bf16* X = cv.take<bf16>(FMAXN * FXP); bf16* XL = cv.take<bf16>(FMAXN * FXP);
  bf16* D = cv.take<bf16>(FM * FXP); bf16* DL = cv.take<bf16>(FM * FXP);        // dxc  [slot][channel]
  bf16* Wt = cv.take<bf16>(FM * FXP); bf16* WtL = cv.take<bf16>(FM * FXP);      // W    [slot][channel]
  bf16* S = cv.take<bf16>(FMAXN * FSP); bf16* SL = cv.take<bf16>(FMAXN * FSP);  // S    [token][slot]
  bf16* G = cv.take<bf16>(FMAXN * FSP); bf16* GL = cv.take<bf16>(FMAXN * FSP);  // dlogits
  bf16* H = cv.take<bf16>(FMAXN * FSP); bf16* HL = cv.take<bf16>(FMAXN * FSP);  // dlogits * rstd
  float* F = cv.take<float>(FMAXN * FM);                                        // S fp32
  float* Zs = cv.take<float>(FMAXN * FM);                                       // zc fp32 (forward's normalised projection)
  float* c1s = cv.take<float>(FM);                                              // rowsum(gamma (.) W)
  float* mean_s = cv.take<float>(FMAXN); float* rstd_s = cv.take<float>(FMAXN);
  float* R = cv.take<float>(FM); float* Tacc = cv.take<float>(FM); float* Uacc = cv.take<float>(FM);
  float* RS = cv.take<float>(2 * FMAXN * 2);                                    // [half][row][2] partial row sums
  float* gam = cv.take<float>(FC); float* bet = cv.take<float>(FC);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int MT = N / 16;

  for (int i = tid; i < FM * FC; i += FNT) {
    const int m = i / FC, k = i - m * FC;
    store_split(Wt, WtL, m * FXP + k, W[i]);
  }
  for (int k = tid; k < FC; k += FNT) { gam[k] = gamma[k]; bet[k] = beta[k]; }
  if (tid < FM) { Tacc[tid] = 0.f; Uacc[tid] = 0.f; }
  for (int m = warp; m < FM; m += FNW) {
    float a1 = 0.f;
    for (int k = lane; k < FC; k += 32) a1 = fmaf(W[m * FC + k], gamma[k], a1);
    a1 = warp_sum(a1);
    if (lane == 0) c1s[m] = a1;
  }
  float Pacc[3][4];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) Pacc[i][j] = 0.f;

  XRegs xr;
  DRegs dr;
  float4 sr = make_float4(0.f, 0.f, 0.f, 0.f), zr = make_float4(0.f, 0.f, 0.f, 0.f);
  // the next image's x tile is prefetched into registers (48 per thread) at the top of an iteration; its small tiles (dxc, S, Z: 20
  // more registers) are only pulled into L2 mid-iteration and loaded at the top of their own iteration, ahead of the LayerNorm pass
  // over the x registers -- with all 68 held for a whole iteration, ptxas spilled eight of them right behind their loads and the
  // spill stores waited out the full memory latency (6.5 % of the kernel's stall samples on one STL)
  auto prefetch_x = [&](int b) { xload(xr, x + (long)b * N * FC, N, warp, lane); };
  auto load_small = [&](int b) {
    dload(dr, dxc + (long)b * FM * FC, tid);
    if (tid * 4 < N * FM) {
      sr = __ldg(reinterpret_cast<const float4*>(Sin + (long)b * N * FM) + tid);
      zr = __ldg(reinterpret_cast<const float4*>(Zin + (long)b * N * FM) + tid);
    }
  };
  auto l2_small = [&](int b) {        // 128 B lines: dxc 96, S and Z N / 2 each
    const char* q = nullptr;
    if (tid < 96) q = reinterpret_cast<const char*>(dxc + (long)b * FM * FC) + tid * 128;
    else if (tid < 96 + N / 2) q = reinterpret_cast<const char*>(Sin + (long)b * N * FM) + (tid - 96) * 128;
    else if (tid >= 128 && tid < 128 + N / 2) q = reinterpret_cast<const char*>(Zin + (long)b * N * FM) + (tid - 128) * 128;
    if (q) asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
  };
  if ((int)blockIdx.x < B) prefetch_x(blockIdx.x);
  const int mt = warp & 3, hf = warp >> 2;      // (token tile, slot half | channel half)
  const bool act = mt < MT;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    load_small(b);
    xstore<true>(xr, X, XL, mean_s, rstd_s, nullptr, N, warp, lane, eps);
    dstore(dr, D, DL, tid);
    if (tid * 4 < N * FM) {
      const int n = tid >> 2, c = (tid & 3) * 4;
      *reinterpret_cast<float4*>(F + n * FM + c) = sr;
      *reinterpret_cast<float4*>(Zs + n * FM + c) = zr;
      store_split4(S, SL, n * FSP + c, sr);
    }
    if (tid < FM) R[tid] = 0.f;
    if (b + (int)gridDim.x < B) prefetch_x(b + gridDim.x);
    __syncthreads();
    // ---- dS[N, 16] = x dxc^T ; column sums of S * dS
    float dS[4] = {0.f, 0.f, 0.f, 0.f}, dS1[4] = {0.f, 0.f, 0.f, 0.f}, dS2[4] = {0.f, 0.f, 0.f, 0.f};
    const int n0 = mt * 16 + g, n1 = n0 + 8, m0 = hf * 8 + 2 * t;
    if (act) {
#pragma unroll 4
      for (int ks = 0; ks < FC / 16; ++ks) {
        uint32_t a[4], al[4], bh[2], bl[2];
        ldA(a, X, FXP, mt * 16, ks * 16, lane);
        ldA(al, XL, FXP, mt * 16, ks * 16, lane);
        ldB2(bh, D, FXP, hf * 8, ks * 16, lane);
        ldB2(bl, DL, FXP, hf * 8, ks * 16, lane);
        mma16816(dS, a, bh[0], bh[1]);          // three independent chains
        mma16816(dS1, al, bh[0], bh[1]);
        mma16816(dS2, a, bl[0], bl[1]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) dS[j] += dS1[j] + dS2[j];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float v = F[n0 * FM + m0 + j] * dS[j] + F[n1 * FM + m0 + j] * dS[2 + j];
        v = g_sum(v);
        if (g == 0) atomicAdd(R + m0 + j, v);
      }
    }
    __syncthreads();
    // ---- dlogits = S (dS - colsum); dl' = dlogits * rstd; T += colsum(dlogits); U += colsum(dl' * mean)
    // The LayerNorm backward's two row sums follow algebraically from dlogits (no pass over the [N, C] gradient needed):
    //   sum_k g = sum_m dlogits[n, m] c1[m],   sum_k g xhat = sum_m dlogits[n, m] zc[n, m]        (g = (dlogits W) gamma)
    if (act) {
      const float r0 = rstd_s[n0], r1 = rstd_s[n1], u0 = mean_s[n0], u1 = mean_s[n1];
      float s1a = 0.f, s2a = 0.f, s1b = 0.f, s2b = 0.f;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float cs = R[m0 + j];
        const float d0 = F[n0 * FM + m0 + j] * (dS[j] - cs), d1 = F[n1 * FM + m0 + j] * (dS[2 + j] - cs);
        store_split(G, GL, n0 * FSP + m0 + j, d0);
        store_split(G, GL, n1 * FSP + m0 + j, d1);
        store_split(H, HL, n0 * FSP + m0 + j, d0 * r0);
        store_split(H, HL, n1 * FSP + m0 + j, d1 * r1);
        const float cc = c1s[m0 + j];
        s1a = fmaf(d0, cc, s1a); s2a = fmaf(d0, Zs[n0 * FM + m0 + j], s2a);
        s1b = fmaf(d1, cc, s1b); s2b = fmaf(d1, Zs[n1 * FM + m0 + j], s2b);
        const float tv = g_sum(d0 + d1), uv = g_sum(d0 * r0 * u0 + d1 * r1 * u1);
        if (g == 0) { atomicAdd(Tacc + m0 + j, tv); atomicAdd(Uacc + m0 + j, uv); }
      }
      s1a = quad_sum(s1a); s2a = quad_sum(s2a); s1b = quad_sum(s1b); s2b = quad_sum(s2b);
      if (t == 0) {      // partial over this warp's 8 slots; the other slot half lands in RS[1 - hf]
        *reinterpret_cast<float2*>(RS + (hf * FMAXN + n0) * 2) = make_float2(s1a, s2a);
        *reinterpret_cast<float2*>(RS + (hf * FMAXN + n1) * 2) = make_float2(s1b, s2b);
      }
    }
    __syncthreads();
    if (b + (int)gridDim.x < B) l2_small(b + gridDim.x);
    // ---- P[16, C] += dl'^T x (persistent accumulators): warp = 3 channel tiles of 8
    for (int ks = 0; ks < MT; ++ks) {
      uint32_t a[4], al[4];
      ldAt(a, H, FSP, 0, ks * 16, lane);
      ldAt(al, HL, FSP, 0, ks * 16, lane);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        uint32_t bh[2], bl[2];
        ldBt2(bh, X, FXP, (warp * 3 + i) * 8, ks * 16, lane);
        ldBt2(bl, XL, FXP, (warp * 3 + i) * 8, ks * 16, lane);
        mma3(Pacc[i], a, al, bh[0], bh[1], bl[0], bl[1]);
      }
    }
    // ---- d_ln = dlogits W (K = 16), dxA = S dxc
    uint32_t ag[4], agl[4], as[4], asl[4];
    float r0 = 0.f, r1 = 0.f, u0 = 0.f, u1 = 0.f;
    if (act) {
      ldA(ag, G, FSP, mt * 16, 0, lane);
      ldA(agl, GL, FSP, mt * 16, 0, lane);
      ldA(as, S, FSP, mt * 16, 0, lane);
      ldA(asl, SL, FSP, mt * 16, 0, lane);
      r0 = rstd_s[n0]; r1 = rstd_s[n1]; u0 = mean_s[n0]; u1 = mean_s[n1];
    }
    // ---- dx = S dxc + rstd (g - mean(g) - xhat mean(g xhat))
    if (act) {
      const float2 pa0 = *reinterpret_cast<const float2*>(RS + n0 * 2), pa1 = *reinterpret_cast<const float2*>(RS + (FMAXN + n0) * 2);
      const float2 pb0 = *reinterpret_cast<const float2*>(RS + n1 * 2), pb1 = *reinterpret_cast<const float2*>(RS + (FMAXN + n1) * 2);
      const float m1a = (pa0.x + pa1.x) * (1.f / FC), m2a = (pa0.y + pa1.y) * (1.f / FC);
      const float m1b = (pb0.x + pb1.x) * (1.f / FC), m2b = (pb0.y + pb1.y) * (1.f / FC);
      float* oa = dx + ((long)b * N + n0) * FC;
      float* ob = dx + ((long)b * N + n1) * FC;
#pragma unroll
      for (int pr = 0; pr < 6; ++pr) {
        const int c0 = hf * 96 + pr * 16;
        uint32_t bw[4], bwl[4], bd[4], bdl[4];
        ldBt(bw, Wt, FXP, c0, 0, lane);
        ldBt(bwl, WtL, FXP, c0, 0, lane);
        ldBt(bd, D, FXP, c0, 0, lane);
        ldBt(bdl, DL, FXP, c0, 0, lane);
        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        float acx[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        mma3(acc[0], ag, agl, bw[0], bw[1], bwl[0], bwl[1]);
        mma3(acc[1], ag, agl, bw[2], bw[3], bwl[2], bwl[3]);
        mma3(acx[0], as, asl, bd[0], bd[1], bdl[0], bdl[1]);
        mma3(acx[1], as, asl, bd[2], bd[3], bdl[2], bdl[3]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = c0 + h * 8 + 2 * t;
          const float2 xa = pair_val(X, XL, n0 * FXP + c), xb = pair_val(X, XL, n1 * FXP + c);
          const float g0 = gam[c], g1 = gam[c + 1];
          float2 ra, rb;
          ra.x = acx[h][0] + r0 * (acc[h][0] * g0 - m1a - (xa.x - u0) * r0 * m2a);
          ra.y = acx[h][1] + r0 * (acc[h][1] * g1 - m1a - (xa.y - u0) * r0 * m2a);
          rb.x = acx[h][2] + r1 * (acc[h][2] * g0 - m1b - (xb.x - u1) * r1 * m2b);
          rb.y = acx[h][3] + r1 * (acc[h][3] * g1 - m1b - (xb.y - u1) * r1 * m2b);
          *reinterpret_cast<float2*>(oa + c) = ra;
          *reinterpret_cast<float2*>(ob + c) = rb;
        }
      }
    }
    __syncthreads();
  }
  // ---- flush: dW = gamma (P - U) + beta T ; db = T ; dgamma = colsum_m W (P - U) ; dbeta = colsum_m W T
  float* Pb = reinterpret_cast<float*>(X);    // [16][C] fp32 (12 KB) over the idle x tile
  {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = (warp * 3 + i) * 8 + 2 * t;
      *reinterpret_cast<float2*>(Pb + g * FC + c) = make_float2(Pacc[i][0], Pacc[i][1]);
      *reinterpret_cast<float2*>(Pb + (g + 8) * FC + c) = make_float2(Pacc[i][2], Pacc[i][3]);
    }
  }
  __syncthreads();
  for (int i = tid * 4; i < FM * FC; i += FNT * 4) {          // 16 B vector reductions: a quarter of the atomic operations
    const int m = i / FC, k = i - m * FC;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = gam[k + e] * (Pb[i + e] - Uacc[m]) + bet[k + e] * Tacc[m];
    red_add_v4(dW + i, v[0], v[1], v[2], v[3]);
  }
  for (int k = tid; k < FC; k += FNT) {
    float dg = 0.f, db = 0.f;
#pragma unroll
    for (int m = 0; m < FM; ++m) {
      const float w = W[m * FC + k];
      dg = fmaf(w, Pb[m * FC + k] - Uacc[m], dg);
      db = fmaf(w, Tacc[m], db);
    }
    atomicAdd(dgamma + k, dg);
    atomicAdd(dbeta + k, db);
  }
  if (tid < FM) atomicAdd(dbias + tid, Tacc[tid]);
}

// =============================================================================================== TokenUpMix forward
constexpr size_t UPF_FWD_SMEM = 2 * al16(FMAXN * FSP * 2) + 2 * al16(FM * FXP * 2) + al16(FMAXN * 4) + 2 * al16(FC * 4) + al16(2 * FMAXN * 2 * 4);

__global__ void __launch_bounds__(FNT, 2) upf_fwd_kernel(const float* __restrict__ xc, int B, int N, const float* __restrict__ W,
                                                         const float* __restrict__ bias, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float eps, float* __restrict__ out,
                                                         float* __restrict__ stats) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  Carve cv(smraw);
  bf16* Wu = cv.take<bf16>(FMAXN * FSP); bf16* WuL = cv.take<bf16>(FMAXN * FSP);   // [token][slot]
  bf16* D = cv.take<bf16>(FM * FXP); bf16* DL = cv.take<bf16>(FM * FXP);           // xc [slot][channel]
  float* bs = cv.take<float>(FMAXN);
  float* gam = cv.take<float>(FC); float* bet = cv.take<float>(FC);
  float* RS = cv.take<float>(2 * FMAXN * 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int MT = N / 16;
  for (int i = tid; i < N * FM; i += FNT) store_split(Wu, WuL, (i >> 4) * FSP + (i & 15), W[i]);
  for (int i = tid; i < N; i += FNT) bs[i] = bias[i];
  for (int k = tid; k < FC; k += FNT) { gam[k] = gamma[k]; bet[k] = beta[k]; }
  DRegs dr;
  if ((int)blockIdx.x < B) dload(dr, xc + (long)blockIdx.x * FM * FC, tid);
  const int mt = warp & 3, hf = warp >> 2;
  const bool act = mt < MT;
  const int n0 = mt * 16 + g, n1 = n0 + 8;
  __syncthreads();
  uint32_t a[4] = {0u, 0u, 0u, 0u}, al[4] = {0u, 0u, 0u, 0u};
  float b0 = 0.f, b1 = 0.f;
  if (act) {
    ldA(a, Wu, FSP, mt * 16, 0, lane);
    ldA(al, WuL, FSP, mt * 16, 0, lane);
    b0 = bs[n0]; b1 = bs[n1];
  }
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    dstore(dr, D, DL, tid);
    if (b + (int)gridDim.x < B) dload(dr, xc + (long)(b + gridDim.x) * FM * FC, tid);
    __syncthreads();
    float acc[6][2][4];
    float ma = 0.f, mb = 0.f;
    if (act) {
      float sa_ = 0.f, sb_ = 0.f;
#pragma unroll
      for (int pr = 0; pr < 6; ++pr) {
        const int c0 = hf * 96 + pr * 16;
        uint32_t bd[4], bdl[4];
        ldBt(bd, D, FXP, c0, 0, lane);
        ldBt(bdl, DL, FXP, c0, 0, lane);
#pragma unroll
        for (int h = 0; h < 2; ++h) { acc[pr][h][0] = b0; acc[pr][h][1] = b0; acc[pr][h][2] = b1; acc[pr][h][3] = b1; }
        mma3(acc[pr][0], a, al, bd[0], bd[1], bdl[0], bdl[1]);
        mma3(acc[pr][1], a, al, bd[2], bd[3], bdl[2], bdl[3]);
#pragma unroll
        for (int h = 0; h < 2; ++h) { sa_ += acc[pr][h][0] + acc[pr][h][1]; sb_ += acc[pr][h][2] + acc[pr][h][3]; }
      }
      ma = quad_sum(sa_) * (1.f / 96.f); mb = quad_sum(sb_) * (1.f / 96.f);      // mean over this warp's 96 channels
      float qa = 0.f, qb = 0.f;
#pragma unroll
      for (int pr = 0; pr < 6; ++pr)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float d0 = acc[pr][h][0] - ma, d1 = acc[pr][h][1] - ma, d2 = acc[pr][h][2] - mb, d3 = acc[pr][h][3] - mb;
          qa += d0 * d0 + d1 * d1; qb += d2 * d2 + d3 * d3;
        }
      qa = quad_sum(qa); qb = quad_sum(qb);
      if (t == 0) {
        *reinterpret_cast<float2*>(RS + (hf * FMAXN + n0) * 2) = make_float2(ma, qa);
        *reinterpret_cast<float2*>(RS + (hf * FMAXN + n1) * 2) = make_float2(mb, qb);
      }
    }
    __syncthreads();
    if (act) {
      // Chan's combination of the two half-row (mean, M2) pairs: n_a = n_b = 96
      const float2 oa = *reinterpret_cast<const float2*>(RS + ((1 - hf) * FMAXN + n0) * 2);
      const float2 ob = *reinterpret_cast<const float2*>(RS + ((1 - hf) * FMAXN + n1) * 2);
      const float2 wa = *reinterpret_cast<const float2*>(RS + (hf * FMAXN + n0) * 2);
      const float2 wb = *reinterpret_cast<const float2*>(RS + (hf * FMAXN + n1) * 2);
      const float mean_a = 0.5f * (wa.x + oa.x), mean_b = 0.5f * (wb.x + ob.x);
      const float da = wa.x - oa.x, db = wb.x - ob.x;
      const float ra = rsqrtf((wa.y + oa.y + da * da * 48.f) * (1.f / FC) + eps);
      const float rb = rsqrtf((wb.y + ob.y + db * db * 48.f) * (1.f / FC) + eps);
      if (hf == 0 && t == 0) {
        *reinterpret_cast<float2*>(stats + ((long)b * N + n0) * 2) = make_float2(mean_a, ra);
        *reinterpret_cast<float2*>(stats + ((long)b * N + n1) * 2) = make_float2(mean_b, rb);
      }
      float* pa = out + ((long)b * N + n0) * FC;
      float* pb = out + ((long)b * N + n1) * FC;
#pragma unroll
      for (int pr = 0; pr < 6; ++pr)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = hf * 96 + pr * 16 + h * 8 + 2 * t;
          const float g0 = gam[c], g1 = gam[c + 1], e0 = bet[c], e1 = bet[c + 1];
          *reinterpret_cast<float2*>(pa + c) = make_float2(fmaf((acc[pr][h][0] - mean_a) * ra, g0, e0), fmaf((acc[pr][h][1] - mean_a) * ra, g1, e1));
          *reinterpret_cast<float2*>(pb + c) = make_float2(fmaf((acc[pr][h][2] - mean_b) * rb, g0, e0), fmaf((acc[pr][h][3] - mean_b) * rb, g1, e1));
        }
    }
    __syncthreads();
  }
}

// =============================================================================================== TokenUpMix backward
constexpr size_t UPF_BWD_SMEM = 2 * al16(FMAXN * FSP * 2) + 2 * al16(FM * FXP * 2) + 2 * al16(FMAXN * FXP * 2) + al16(FMAXN * 4) + al16(FC * 4) +
                                al16(2 * FMAXN * 2 * 4) + al16(2 * FC * 4);

__global__ void __launch_bounds__(FNT, 1) upf_bwd_kernel(const float* __restrict__ xc, const float* __restrict__ dout,
                                                         const float* __restrict__ stats, int B, int N, const float* __restrict__ W,
                                                         const float* __restrict__ bias, const float* __restrict__ gamma,
                                                         float* __restrict__ dxc, float* __restrict__ dW, float* __restrict__ dgamma,
                                                         float* __restrict__ dbeta) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  Carve cv(smraw);
  bf16* Wu = cv.take<bf16>(FMAXN * FSP); bf16* WuL = cv.take<bf16>(FMAXN * FSP);   // [token][slot]
  bf16* D = cv.take<bf16>(FM * FXP); bf16* DL = cv.take<bf16>(FM * FXP);           // xc [slot][channel]
  bf16* G = cv.take<bf16>(FMAXN * FXP); bf16* GL = cv.take<bf16>(FMAXN * FXP);     // dout, then dup  [token][channel]
  float* bs = cv.take<float>(FMAXN);
  float* gam = cv.take<float>(FC);
  float* RS = cv.take<float>(2 * FMAXN * 2);
  float* CS = cv.take<float>(2 * FC);                                              // dgamma | dbeta of this CTA
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int MT = N / 16;
  for (int i = tid; i < N * FM; i += FNT) store_split(Wu, WuL, (i >> 4) * FSP + (i & 15), W[i]);
  for (int i = tid; i < N; i += FNT) bs[i] = bias[i];
  for (int k = tid; k < FC; k += FNT) { gam[k] = gamma[k]; CS[k] = 0.f; CS[FC + k] = 0.f; }
  const int mt = warp & 3, hf = warp >> 2;
  const bool act = mt < MT;
  const int n0 = mt * 16 + g, n1 = n0 + 8;
  XRegs xr;
  DRegs dr;
  float4 st = make_float4(0.f, 1.f, 0.f, 1.f);     // (mean, rstd) of rows n0, n1
  auto prefetch = [&](int b) {
    xload(xr, dout + (long)b * N * FC, N, warp, lane);
    dload(dr, xc + (long)b * FM * FC, tid);
    if (act) {
      const float2 s0 = __ldg(reinterpret_cast<const float2*>(stats + ((long)b * N + n0) * 2));
      const float2 s1 = __ldg(reinterpret_cast<const float2*>(stats + ((long)b * N + n1) * 2));
      st = make_float4(s0.x, s0.y, s1.x, s1.y);
    }
  };
  if ((int)blockIdx.x < B) prefetch(blockIdx.x);
  float dgam[6][2][2], dbet[6][2][2];
#pragma unroll
  for (int pr = 0; pr < 6; ++pr)
#pragma unroll
    for (int h = 0; h < 2; ++h) { dgam[pr][h][0] = dgam[pr][h][1] = 0.f; dbet[pr][h][0] = dbet[pr][h][1] = 0.f; }
  float aW[3][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};   // hi hi | lo hi | hi lo: three independent MMA chains
  __syncthreads();
  uint32_t a[4] = {0u, 0u, 0u, 0u}, al[4] = {0u, 0u, 0u, 0u};
  float b0 = 0.f, b1 = 0.f;
  if (act) {
    ldA(a, Wu, FSP, mt * 16, 0, lane);
    ldA(al, WuL, FSP, mt * 16, 0, lane);
    b0 = bs[n0]; b1 = bs[n1];
  }
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    xstore<false>(xr, G, GL, nullptr, nullptr, nullptr, N, warp, lane, 0.f);
    dstore(dr, D, DL, tid);
    const float ua = st.x, ra = st.y, ub = st.z, rb = st.w;
    if (b + (int)gridDim.x < B) prefetch(b + gridDim.x);
    __syncthreads();
    // ---- pass 1: recompute up, row sums of g = dout gamma and g xhat; column sums for dgamma / dbeta
    if (act) {
      float s1a = 0.f, s2a = 0.f, s1b = 0.f, s2b = 0.f;
#pragma unroll
      for (int pr = 0; pr < 6; ++pr) {
        const int c0 = hf * 96 + pr * 16;
        uint32_t bd[4], bdl[4];
        ldBt(bd, D, FXP, c0, 0, lane);
        ldBt(bdl, DL, FXP, c0, 0, lane);
        float up[2][4] = {{b0, b0, b1, b1}, {b0, b0, b1, b1}};
        mma3(up[0], a, al, bd[0], bd[1], bdl[0], bdl[1]);
        mma3(up[1], a, al, bd[2], bd[3], bdl[2], bdl[3]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = c0 + h * 8 + 2 * t;
          const float2 da = pair_val(G, GL, n0 * FXP + c), db = pair_val(G, GL, n1 * FXP + c);
          const float g0 = gam[c], g1 = gam[c + 1];
          const float xa0 = (up[h][0] - ua) * ra, xa1 = (up[h][1] - ua) * ra, xb0 = (up[h][2] - ub) * rb, xb1 = (up[h][3] - ub) * rb;
          s1a += da.x * g0 + da.y * g1; s2a += da.x * g0 * xa0 + da.y * g1 * xa1;
          s1b += db.x * g0 + db.y * g1; s2b += db.x * g0 * xb0 + db.y * g1 * xb1;
          dgam[pr][h][0] += da.x * xa0 + db.x * xb0; dgam[pr][h][1] += da.y * xa1 + db.y * xb1;
          dbet[pr][h][0] += da.x + db.x; dbet[pr][h][1] += da.y + db.y;
        }
      }
      s1a = quad_sum(s1a); s2a = quad_sum(s2a); s1b = quad_sum(s1b); s2b = quad_sum(s2b);
      if (t == 0) {
        *reinterpret_cast<float2*>(RS + (hf * FMAXN + n0) * 2) = make_float2(s1a, s2a);
        *reinterpret_cast<float2*>(RS + (hf * FMAXN + n1) * 2) = make_float2(s1b, s2b);
      }
    }
    __syncthreads();
    // ---- pass 2: dup = rstd (g - mean(g) - xhat mean(g xhat)), written over dout (same thread, same elements)
    if (act) {
      const float2 pa0 = *reinterpret_cast<const float2*>(RS + n0 * 2), pa1 = *reinterpret_cast<const float2*>(RS + (FMAXN + n0) * 2);
      const float2 pb0 = *reinterpret_cast<const float2*>(RS + n1 * 2), pb1 = *reinterpret_cast<const float2*>(RS + (FMAXN + n1) * 2);
      const float m1a = (pa0.x + pa1.x) * (1.f / FC), m2a = (pa0.y + pa1.y) * (1.f / FC);
      const float m1b = (pb0.x + pb1.x) * (1.f / FC), m2b = (pb0.y + pb1.y) * (1.f / FC);
#pragma unroll
      for (int pr = 0; pr < 6; ++pr) {
        const int c0 = hf * 96 + pr * 16;
        uint32_t bd[4], bdl[4];
        ldBt(bd, D, FXP, c0, 0, lane);
        ldBt(bdl, DL, FXP, c0, 0, lane);
        float up[2][4] = {{b0, b0, b1, b1}, {b0, b0, b1, b1}};
        mma3(up[0], a, al, bd[0], bd[1], bdl[0], bdl[1]);
        mma3(up[1], a, al, bd[2], bd[3], bdl[2], bdl[3]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = c0 + h * 8 + 2 * t;
          const float2 da = pair_val(G, GL, n0 * FXP + c), db = pair_val(G, GL, n1 * FXP + c);
          const float g0 = gam[c], g1 = gam[c + 1];
          const float xa0 = (up[h][0] - ua) * ra, xa1 = (up[h][1] - ua) * ra, xb0 = (up[h][2] - ub) * rb, xb1 = (up[h][3] - ub) * rb;
          const float ya0 = ra * (da.x * g0 - m1a - xa0 * m2a), ya1 = ra * (da.y * g1 - m1a - xa1 * m2a);
          const float yb0 = rb * (db.x * g0 - m1b - xb0 * m2b), yb1 = rb * (db.y * g1 - m1b - xb1 * m2b);
          uint32_t h0, l0, h1, l1;
          split2(ya0, ya1, &h0, &l0);
          split2(yb0, yb1, &h1, &l1);
          *reinterpret_cast<uint32_t*>(G + n0 * FXP + c) = h0; *reinterpret_cast<uint32_t*>(GL + n0 * FXP + c) = l0;
          *reinterpret_cast<uint32_t*>(G + n1 * FXP + c) = h1; *reinterpret_cast<uint32_t*>(GL + n1 * FXP + c) = l1;
        }
      }
    }
    __syncthreads();
    // ---- dxc[16, C] = W^T dup: warp = 3 channel tiles of 8
    {
      float acc[3][4];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      for (int ks = 0; ks < MT; ++ks) {
        uint32_t wa[4], wl[4];
        ldAt(wa, Wu, FSP, 0, ks * 16, lane);
        ldAt(wl, WuL, FSP, 0, ks * 16, lane);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          uint32_t bh[2], bl[2];
          ldBt2(bh, G, FXP, (warp * 3 + i) * 8, ks * 16, lane);
          ldBt2(bl, GL, FXP, (warp * 3 + i) * 8, ks * 16, lane);
          mma3(acc[i], wa, wl, bh[0], bh[1], bl[0], bl[1]);
        }
      }
      float* o = dxc + (long)b * FM * FC + 2 * t;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int c = (warp * 3 + i) * 8;
        *reinterpret_cast<float2*>(o + (long)g * FC + c) = make_float2(acc[i][0], acc[i][1]);
        *reinterpret_cast<float2*>(o + (long)(g + 8) * FC + c) = make_float2(acc[i][2], acc[i][3]);
      }
    }
    // ---- dW[N, 16] += dup xc^T (K = C): warp = (token tile, slot half), accumulated over the CTA's images
    if (act) {
#pragma unroll 4
      for (int ks = 0; ks < FC / 16; ++ks) {
        uint32_t ga[4], gl[4], bh[2], bl[2];
        ldA(ga, G, FXP, mt * 16, ks * 16, lane);
        ldA(gl, GL, FXP, mt * 16, ks * 16, lane);
        ldB2(bh, D, FXP, hf * 8, ks * 16, lane);
        ldB2(bl, DL, FXP, hf * 8, ks * 16, lane);
        mma16816(aW[0], ga, bh[0], bh[1]);      // (one accumulator made the 36 MMAs of an image one dependent chain)
        mma16816(aW[1], gl, bh[0], bh[1]);
        mma16816(aW[2], ga, bl[0], bl[1]);
      }
    }
    __syncthreads();
  }
  // ---- flush
  if (act) {
    const int m = hf * 8 + 2 * t;
    atomicAdd(dW + n0 * FM + m, aW[0][0] + aW[1][0] + aW[2][0]); atomicAdd(dW + n0 * FM + m + 1, aW[0][1] + aW[1][1] + aW[2][1]);
    atomicAdd(dW + n1 * FM + m, aW[0][2] + aW[1][2] + aW[2][2]); atomicAdd(dW + n1 * FM + m + 1, aW[0][3] + aW[1][3] + aW[2][3]);
#pragma unroll
    for (int pr = 0; pr < 6; ++pr)
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float vg = g_sum(dgam[pr][h][j]), vb = g_sum(dbet[pr][h][j]);
          if (g == 0) {
            const int c = hf * 96 + pr * 16 + h * 8 + 2 * t + j;
            atomicAdd(CS + c, vg);
            atomicAdd(CS + FC + c, vb);
          }
        }
  }
  __syncthreads();
  for (int k = tid; k < FC; k += FNT) { atomicAdd(dgamma + k, CS[k]); atomicAdd(dbeta + k, CS[FC + k]); }
}

template <typename K>
int opt_in(K kernel, size_t bytes) {
  QV_CHECK(bytes <= 227 * 1024, "kernel needs %zu B of shared memory (> 227 KB)", bytes);
  if (bytes > 48 * 1024) QV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}
int grid_for(int B, int per_sm) { return max(1, min(B, qv_num_sms() * per_sm)); }

}  // namespace

bool tokens_fused_ok(int M, int N, int C) { return M == FM && C == FC && N >= 16 && N <= FMAXN && N % 16 == 0; }

int tlf_fwd(cudaStream_t s, const float* x, int B, int N, const float* gamma, const float* beta, const float* W, const float* bias,
            float eps, float* S, float* Z, float* xc) {
  if (B <= 0) return 0;
  QV_TRY(opt_in(tlf_fwd_kernel, TLF_FWD_SMEM));
  qv_launch(tlf_fwd_kernel, grid_for(B, 2), FNT, TLF_FWD_SMEM, s, x, B, N, gamma, beta, W, bias, eps, S, Z, xc);
  QV_LAUNCH_CHECK();
  return 0;
}
int tlf_bwd(cudaStream_t s, const float* x, const float* S, const float* Z, const float* dxc, int B, int N, const float* gamma,
            const float* beta, const float* W, float eps, float* dx, float* dW, float* dbias, float* dgamma, float* dbeta) {
  QV_CHECK(((uintptr_t)dW & 15) == 0, "tlf_bwd: dW must be 16 B aligned (vector reductions)");
  if (B <= 0) return 0;
  QV_TRY(opt_in(tlf_bwd_kernel, TLF_BWD_SMEM));
  qv_launch(tlf_bwd_kernel, grid_for(B, 2), FNT, TLF_BWD_SMEM, s, x, S, Z, dxc, B, N, gamma, beta, W, eps, dx, dW, dbias, dgamma, dbeta);
  QV_LAUNCH_CHECK();
  return 0;
}
int upf_fwd(cudaStream_t s, const float* xc, int B, int N, const float* W, const float* bias, const float* gamma, const float* beta,
            float eps, float* out, float* stats) {
  if (B <= 0) return 0;
  QV_TRY(opt_in(upf_fwd_kernel, UPF_FWD_SMEM));
  qv_launch(upf_fwd_kernel, grid_for(B, 2), FNT, UPF_FWD_SMEM, s, xc, B, N, W, bias, gamma, beta, eps, out, stats);
  QV_LAUNCH_CHECK();
  return 0;
}
int upf_bwd(cudaStream_t s, const float* xc, const float* dout, const float* stats, int B, int N, const float* W, const float* bias,
            const float* gamma, float* dxc, float* dW, float* dgamma, float* dbeta) {
  if (B <= 0) return 0;
  QV_TRY(opt_in(upf_bwd_kernel, UPF_BWD_SMEM));
  qv_launch(upf_bwd_kernel, grid_for(B, 1), FNT, UPF_BWD_SMEM, s, xc, dout, stats, B, N, W, bias, gamma, dxc, dW, dgamma, dbeta);
  QV_LAUNCH_CHECK();
  return 0;
}
