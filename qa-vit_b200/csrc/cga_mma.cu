// Channel-group attention (H:559-595) on mma.sync.m16n8k16 for the 16-token blocks of HQAViT (bf16 runs):
// one WARP per image; the per-group q/k/v projections, the 4 heads of head_dim 4 (contraction padded to 16 by masking
// the other heads' columns to zero), softmax, PV and -- in backward -- all their transposes stay in registers and
// ~5 KB of warp-private shared memory.  The SIMT kernels in cga.cu remain the fp32 parity path and the path for
// other token counts.
#include "kernels.h"

namespace {

constexpr int NQ = 16, CG = 32, CPG = 16, NH = 4, KB = 16, NKEY = NQ + KB;   // 32 keys: own tokens then bank
constexpr int PW = 40;    // pitch of the stacked [48][32] projection weight
constexpr int PK = 24;    // pitch of 16-wide rows (q, k, v)
constexpr int PP = 40;    // pitch of 32-wide rows (P, dS)
constexpr int PD = 56;    // pitch of the 48-wide [dq | dk | dv] staging
constexpr int WARPS = 4;
#ifndef CGA_WB
#define CGA_WB 12
#endif
constexpr int WARPS_BWD = CGA_WB;   // backward: 17.5 KB of warp-private tiles + 168 registers per thread -- ONE 12-warp CTA fills an SM's shared
                                // memory (223 KB) and register file (63 K); 4-warp CTAs fitted only twice (8 warps per SM, ncu: 12 % occupancy)

__device__ __forceinline__ uint32_t sa(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm4(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float x, float y) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void ldA(uint32_t* a, const bf16* base, int pitch, int m0, int k0, int lane) {   // smem [m][k]
  const int mat = lane >> 3, r = lane & 7;
  ldsm4(a, sa(base + (m0 + r + (mat & 1) * 8) * pitch + k0 + (mat >> 1) * 8));
}
__device__ __forceinline__ void ldAt(uint32_t* a, const bf16* base, int pitch, int m0, int k0, int lane) {  // smem [k][m]
  const int mat = lane >> 3, r = lane & 7;
  ldsm4t(a, sa(base + (k0 + r + (mat >> 1) * 8) * pitch + m0 + (mat & 1) * 8));
}
__device__ __forceinline__ void ldB(uint32_t* b, const bf16* base, int pitch, int n0, int k0, int lane) {   // smem [n][k], 2 n-tiles
  const int mat = lane >> 3, r = lane & 7;
  ldsm4(b, sa(base + (n0 + r + (mat >> 1) * 8) * pitch + k0 + (mat & 1) * 8));
}
__device__ __forceinline__ void ldBt(uint32_t* b, const bf16* base, int pitch, int n0, int k0, int lane) {  // smem [k][n], 2 n-tiles
  const int mat = lane >> 3, r = lane & 7;
  ldsm4t(b, sa(base + (k0 + r + (mat & 1) * 8) * pitch + n0 + (mat >> 1) * 8));
}
__device__ __forceinline__ void stC(bf16* base, int pitch, int m0, int n0, const float* c, int lane) {
  const int g = lane >> 2, t = lane & 3;
  *reinterpret_cast<uint32_t*>(base + (m0 + g) * pitch + n0 + 2 * t) = pack2(c[0], c[1]);
  *reinterpret_cast<uint32_t*>(base + (m0 + g + 8) * pitch + n0 + 2 * t) = pack2(c[2], c[3]);
}

struct WS {  // per-warp bf16 regions
  static constexpr int X = 0, Q = X + NQ * 200, K = Q + NQ * PK, V = K + NKEY * PK, DO = V + NKEY * PK;
  static constexpr int P = DO + NQ * 104, DS = P + NQ * PP, DQ = DS + NQ * PP, END_FWD = DO, END_BWD = DQ + NQ * PD;
};

// A fragment (k = the 16 compressed channels of a group) with every head but `h` masked to zero
__device__ __forceinline__ void head_frag(uint32_t* a, const float (*x)[4], int h, int lane) {
  const bool keep = (((lane & 3) >> 1) == (h & 1));
  const int n = h >> 1;
  const uint32_t lo = keep ? pack2(x[n][0], x[n][1]) : 0u, hi = keep ? pack2(x[n][2], x[n][3]) : 0u;
  a[0] = n == 0 ? lo : 0u; a[1] = n == 0 ? hi : 0u; a[2] = n == 1 ? lo : 0u; a[3] = n == 1 ? hi : 0u;
}
// acc[n-tile of head h] (+)= src[n-tile of head h] on the 4 columns of head h only
__device__ __forceinline__ void head_keep(float (*acc)[4], const float (*src)[4], int h, int lane, bool add) {
  if ((((lane & 3) >> 1) == (h & 1))) {
    const int n = h >> 1;
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[n][e] = add ? acc[n][e] + src[n][e] : src[n][e];
  }
}

__device__ __forceinline__ void load_shared_consts(const CgaP& p, bf16* Wst, float* bias, bf16* Kb, bf16* Vb) {
  for (int i = threadIdx.x; i < 3 * CPG * CG; i += blockDim.x) {
    const int which = i / (CPG * CG), r = i % (CPG * CG);
    const float* src = which == 0 ? p.Wq : (which == 1 ? p.Wk : p.Wv);
    Wst[(which * CPG + r / CG) * PW + r % CG] = __float2bfloat16_rn(src[r]);
  }
  for (int i = threadIdx.x; i < 3 * CPG; i += blockDim.x) bias[i] = i < CPG ? p.bq[i] : (i < 2 * CPG ? p.bk[i - CPG] : p.bv[i - 2 * CPG]);
  for (int i = threadIdx.x; i < KB * CPG; i += blockDim.x) {
    Kb[(i / CPG) * PK + i % CPG] = __float2bfloat16_rn(p.kbp[i]);
    Vb[(i / CPG) * PK + i % CPG] = __float2bfloat16_rn(p.vbp[i]);
  }
}

// xn rows of image b (16 x 192 bf16) -> X tile; bank rows -> K/V rows 16..31
__device__ __forceinline__ void load_image(const CgaP& p, bf16* W, const bf16* Kb, const bf16* Vb, int b, int lane, int ncols) {
  const bf16* xn = static_cast<const bf16*>(p.xn);
  const int cpr = ncols / 8;                                   // 16 B chunks per row
  for (int c = lane; c < NQ * cpr; c += 32) {
    const int i = c / cpr, ch = c % cpr;
    *reinterpret_cast<uint4*>(W + WS::X + i * 200 + ch * 8) = *reinterpret_cast<const uint4*>(xn + ((long)b * NQ + i) * p.ldx + ch * 8);
  }
  for (int c = lane; c < KB * 2; c += 32) {
    const int i = c >> 1, ch = c & 1;
    *reinterpret_cast<uint4*>(W + WS::K + (NQ + i) * PK + ch * 8) = *reinterpret_cast<const uint4*>(Kb + i * PK + ch * 8);
    *reinterpret_cast<uint4*>(W + WS::V + (NQ + i) * PK + ch * 8) = *reinterpret_cast<const uint4*>(Vb + i * PK + ch * 8);
  }
}

// q | k | v = x_g W^T + b in C layout: acc[0..1] = q, [2..3] = k, [4..5] = v; k, v (and optionally q) -> smem rows 0..15
__device__ __forceinline__ void project(float (*acc)[4], bf16* W, const bf16* Wst, const float* bias, int grp, int lane, bool store_q) {
  const int t = lane & 3;
#pragma unroll
  for (int n = 0; n < 6; ++n) {
    acc[n][0] = acc[n][2] = bias[n * 8 + 2 * t];
    acc[n][1] = acc[n][3] = bias[n * 8 + 2 * t + 1];
  }
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    uint32_t a[4];
    ldA(a, W + WS::X, 200, 0, grp * CG + kk * 16, lane);
#pragma unroll
    for (int np = 0; np < 3; ++np) {
      uint32_t b[4];
      ldB(b, Wst, PW, np * 16, kk * 16, lane);
      mma16816(acc[2 * np], a, b[0], b[1]);
      mma16816(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
  if (store_q) { stC(W + WS::Q, PK, 0, 0, acc[0], lane); stC(W + WS::Q, PK, 0, 8, acc[1], lane); }
  stC(W + WS::K, PK, 0, 0, acc[2], lane); stC(W + WS::K, PK, 0, 8, acc[3], lane);
  stC(W + WS::V, PK, 0, 0, acc[4], lane); stC(W + WS::V, PK, 0, 8, acc[5], lane);
  __syncwarp();
}

// softmax(q_h K^T / 2) over the 32 keys, C layout s[4][4]
__device__ __forceinline__ void head_scores(float (*s)[4], const float (*q)[4], const bf16* Ks, int h, int lane) {
  uint32_t a[4];
  head_frag(a, q, h, lane);
#pragma unroll
  for (int n = 0; n < 4; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
  for (int np = 0; np < 2; ++np) {
    uint32_t b[4];
    ldB(b, Ks, PK, np * 16, 0, lane);
    mma16816(s[2 * np], a, b[0], b[1]);
    mma16816(s[2 * np + 1], a, b[2], b[3]);
  }
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    s[n][0] *= 0.5f; s[n][1] *= 0.5f; s[n][2] *= 0.5f; s[n][3] *= 0.5f;
    m0 = fmaxf(m0, fmaxf(s[n][0], s[n][1]));
    m1 = fmaxf(m1, fmaxf(s[n][2], s[n][3]));
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  float z0 = 0.f, z1 = 0.f;
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    s[n][0] = __expf(s[n][0] - m0); s[n][1] = __expf(s[n][1] - m0);
    s[n][2] = __expf(s[n][2] - m1); s[n][3] = __expf(s[n][3] - m1);
    z0 += s[n][0] + s[n][1];
    z1 += s[n][2] + s[n][3];
  }
  z0 += __shfl_xor_sync(0xffffffffu, z0, 1); z0 += __shfl_xor_sync(0xffffffffu, z0, 2);
  z1 += __shfl_xor_sync(0xffffffffu, z1, 1); z1 += __shfl_xor_sync(0xffffffffu, z1, 2);
  z0 = 1.f / z0; z1 = 1.f / z1;
#pragma unroll
  for (int n = 0; n < 4; ++n) { s[n][0] *= z0; s[n][1] *= z0; s[n][2] *= z1; s[n][3] *= z1; }
}

// o[2][4] = X[16 x 32 keys] (C-layout regs as A) * Bs (smem [key][dim])
__device__ __forceinline__ void keys_times(float (*o)[4], const float (*x)[4], const bf16* Bs, int lane) {
#pragma unroll
  for (int n = 0; n < 2; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    uint32_t a[4] = {pack2(x[2 * kk][0], x[2 * kk][1]), pack2(x[2 * kk][2], x[2 * kk][3]),
                     pack2(x[2 * kk + 1][0], x[2 * kk + 1][1]), pack2(x[2 * kk + 1][2], x[2 * kk + 1][3])};
    uint32_t b[4];
    ldBt(b, Bs, PK, 0, kk * 16, lane);
    mma16816(o[0], a, b[0], b[1]);
    mma16816(o[1], a, b[2], b[3]);
  }
}

__global__ void __launch_bounds__(WARPS * 32) cga_mma_fwd_kernel(CgaP p) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  float* bias = reinterpret_cast<float*>(smraw);                    // [48]
  bf16* Wst = reinterpret_cast<bf16*>(bias + 3 * CPG);              // [48][PW]
  bf16* Kb = Wst + 3 * CPG * PW;                                    // [16][PK]
  bf16* Vb = Kb + KB * PK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  bf16* W = Vb + KB * PK + warp * WS::END_FWD;
  load_shared_consts(p, Wst, bias, Kb, Vb);
  __syncthreads();
  bf16* out = static_cast<bf16*>(p.out);
  for (int b = blockIdx.x * WARPS + warp; b < p.B; b += gridDim.x * WARPS) {
    __syncwarp();
    load_image(p, W, Kb, Vb, b, lane, p.G * CG);
    __syncwarp();
    for (int grp = 0; grp < p.G; ++grp) {
      float acc[6][4], o[2][4];
      project(acc, W, Wst, bias, grp, lane, false);
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        float s[4][4], o2[2][4];
        head_scores(s, acc, W + WS::K, h, lane);
        if (p.drop.p > 0.f) {   // SDPA dropout_p on the probabilities
          const DropState ds = drop_state(p.drop);
          drop_apply_c<4>(s, drop_bits_c<4>(ds, (uint32_t)((b * p.G + grp) * NH + h), lane), ds.inv);
        }
        keys_times(o2, s, W + WS::V, lane);
        head_keep(o, o2, h, lane, false);
      }
      const long r0 = (long)b * NQ + g, r1 = r0 + 8;
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        *reinterpret_cast<uint32_t*>(out + r0 * p.ldo + grp * CPG + n * 8 + 2 * t) = pack2(o[n][0], o[n][1]);
        *reinterpret_cast<uint32_t*>(out + r1 * p.ldo + grp * CPG + n * 8 + 2 * t) = pack2(o[n][2], o[n][3]);
      }
      __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(WARPS_BWD * 32, 1) cga_mma_bwd_kernel(CgaP p) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  float* bias = reinterpret_cast<float*>(smraw);                    // [48]
  float* dkb = bias + 3 * CPG;                                      // [16][16] d(projected bank k), then v
  float* dvb = dkb + KB * CPG;
  bf16* Wst = reinterpret_cast<bf16*>(dvb + KB * CPG);
  bf16* Kb = Wst + 3 * CPG * PW;
  bf16* Vb = Kb + KB * PK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  bf16* W = Vb + KB * PK + warp * WS::END_BWD;
  load_shared_consts(p, Wst, bias, Kb, Vb);
  for (int i = threadIdx.x; i < 2 * KB * CPG; i += blockDim.x) dkb[i] = 0.f;
  __syncthreads();
  const bf16* dout = static_cast<const bf16*>(p.dout);
  float dWacc[3][4][4], dbacc[6][2];
  constexpr int WARPS = WARPS_BWD;
#pragma unroll
  for (int m = 0; m < 3; ++m)
#pragma unroll
    for (int n = 0; n < 4; ++n) dWacc[m][n][0] = dWacc[m][n][1] = dWacc[m][n][2] = dWacc[m][n][3] = 0.f;
#pragma unroll
  for (int n = 0; n < 6; ++n) dbacc[n][0] = dbacc[n][1] = 0.f;
  float dkbr[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dvbr[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};

  for (int b = blockIdx.x * WARPS + warp; b < p.B; b += gridDim.x * WARPS) {
    __syncwarp();
    load_image(p, W, Kb, Vb, b, lane, p.G * CG);
    for (int c = lane; c < NQ * (p.G * CPG / 8); c += 32) {          // dO tile [16][G*16]
      const int cpr = p.G * CPG / 8, i = c / cpr, ch = c % cpr;
      *reinterpret_cast<uint4*>(W + WS::DO + i * 104 + ch * 8) = *reinterpret_cast<const uint4*>(dout + ((long)b * NQ + i) * p.lddo + ch * 8);
    }
    __syncwarp();
    for (int grp = 0; grp < p.G; ++grp) {
      float acc[6][4];
      project(acc, W, Wst, bias, grp, lane, true);
      // dO of this group in C layout (for the masked A fragments): read back from the staged tile
      float dO[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(W + WS::DO + g * 104 + grp * CPG + n * 8 + 2 * t));
        const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(W + WS::DO + (g + 8) * 104 + grp * CPG + n * 8 + 2 * t));
        dO[n][0] = lo.x; dO[n][1] = lo.y; dO[n][2] = hi.x; dO[n][3] = hi.y;
      }
      float dq[2][4], dk[2][2][4], dv[2][2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) { dq[n][e] = 0.f; dk[0][n][e] = dk[1][n][e] = dv[0][n][e] = dv[1][n][e] = 0.f; }
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        float P[4][4], dS[4][4];
        head_scores(P, acc, W + WS::K, h, lane);
        // dP = dO_h V^T
        uint32_t a[4];
        head_frag(a, dO, h, lane);
#pragma unroll
        for (int n = 0; n < 4; ++n) dS[n][0] = dS[n][1] = dS[n][2] = dS[n][3] = 0.f;
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t bb[4];
          ldB(bb, W + WS::V, PK, np * 16, 0, lane);
          mma16816(dS[2 * np], a, bb[0], bb[1]);
          mma16816(dS[2 * np + 1], a, bb[2], bb[3]);
        }
        unsigned long long keep = ~0ull;
        float kinv = 1.f;
        if (p.drop.p > 0.f) {   // dP <- d(P_dropped) * keep; P_dropped (for dv) is formed after the softmax backward
          const DropState ds = drop_state(p.drop);
          keep = drop_bits_c<4>(ds, (uint32_t)((b * p.G + grp) * NH + h), lane);
          kinv = ds.inv;
          drop_apply_c<4>(dS, keep, kinv);
        }
        float r0 = 0.f, r1 = 0.f;
#pragma unroll
        for (int n = 0; n < 4; ++n) { r0 += dS[n][0] * P[n][0] + dS[n][1] * P[n][1]; r1 += dS[n][2] * P[n][2] + dS[n][3] * P[n][3]; }
        r0 += __shfl_xor_sync(0xffffffffu, r0, 1); r0 += __shfl_xor_sync(0xffffffffu, r0, 2);
        r1 += __shfl_xor_sync(0xffffffffu, r1, 1); r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          dS[n][0] = P[n][0] * (dS[n][0] - r0) * 0.5f; dS[n][1] = P[n][1] * (dS[n][1] - r0) * 0.5f;
          dS[n][2] = P[n][2] * (dS[n][2] - r1) * 0.5f; dS[n][3] = P[n][3] * (dS[n][3] - r1) * 0.5f;
        }
        if (p.drop.p > 0.f) drop_apply_c<4>(P, keep, kinv);
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          stC(W + WS::P, PP, 0, n * 8, P[n], lane);
          stC(W + WS::DS, PP, 0, n * 8, dS[n], lane);
        }
        // dq_h = dS K
        float t2[2][4];
        keys_times(t2, dS, W + WS::K, lane);
        head_keep(dq, t2, h, lane, true);
        __syncwarp();
        // dk_h = dS^T q, dv_h = P^T dO   (M = 32 keys, N = 16 channels, K = 16 queries)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          uint32_t aa[4], bq[4], bo[4];
          float ck[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, cv[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
          ldBt(bq, W + WS::Q, PK, 0, 0, lane);
          ldBt(bo, W + WS::DO, 104, grp * CPG, 0, lane);
          ldAt(aa, W + WS::DS, PP, mt * 16, 0, lane);
          mma16816(ck[0], aa, bq[0], bq[1]);
          mma16816(ck[1], aa, bq[2], bq[3]);
          ldAt(aa, W + WS::P, PP, mt * 16, 0, lane);
          mma16816(cv[0], aa, bo[0], bo[1]);
          mma16816(cv[1], aa, bo[2], bo[3]);
          head_keep(dk[mt], ck, h, lane, true);
          head_keep(dv[mt], cv, h, lane, true);
        }
        __syncwarp();
      }
      // bank rows (keys 16..31) -> d(projected bank): summed in registers over groups and images, flushed once per warp (shared-memory
      // fp32 atomics are compare-and-swap loops: 16 per lane and group with 12 warps on the same 512 addresses)
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) { dkbr[n][e] += dk[1][n][e]; dvbr[n][e] += dv[1][n][e]; }
      // [dq | dk | dv] (16 x 48): bias grads, staging for dW, A operand for dx
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        dbacc[n][0] += dq[n][0] + dq[n][2]; dbacc[n][1] += dq[n][1] + dq[n][3];
        dbacc[2 + n][0] += dk[0][n][0] + dk[0][n][2]; dbacc[2 + n][1] += dk[0][n][1] + dk[0][n][3];
        dbacc[4 + n][0] += dv[0][n][0] + dv[0][n][2]; dbacc[4 + n][1] += dv[0][n][1] + dv[0][n][3];
        stC(W + WS::DQ, PD, 0, n * 8, dq[n], lane);
        stC(W + WS::DQ, PD, 0, 16 + n * 8, dk[0][n], lane);
        stC(W + WS::DQ, PD, 0, 32 + n * 8, dv[0][n], lane);
      }
      __syncwarp();
      // dx_g[16 x 32] = [dq|dk|dv] Wst : A from smem [n][o], B = Wst [k = o][n = c]
      {
        float dx[4][4];
#pragma unroll
        for (int n = 0; n < 4; ++n) dx[n][0] = dx[n][1] = dx[n][2] = dx[n][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
          uint32_t aa[4];
          ldA(aa, W + WS::DQ, PD, 0, kk * 16, lane);
#pragma unroll
          for (int np = 0; np < 2; ++np) {
            uint32_t bb[4];
            ldBt(bb, Wst, PW, np * 16, kk * 16, lane);
            mma16816(dx[2 * np], aa, bb[0], bb[1]);
            mma16816(dx[2 * np + 1], aa, bb[2], bb[3]);
          }
        }
        float* d0 = p.dxn + ((long)b * NQ + g) * p.lddx + grp * CG + 2 * t;
        float* d1 = d0 + 8L * p.lddx;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          float2 u = *reinterpret_cast<float2*>(d0 + n * 8), v = *reinterpret_cast<float2*>(d1 + n * 8);
          u.x += dx[n][0]; u.y += dx[n][1]; v.x += dx[n][2]; v.y += dx[n][3];
          *reinterpret_cast<float2*>(d0 + n * 8) = u;
          *reinterpret_cast<float2*>(d1 + n * 8) = v;
        }
      }
      // dW[o, c] += sum_n dqkv[n, o] x_g[n, c] : A = dqkv^T (smem [k = n][m = o]), B = X [k = n][n = c]
#pragma unroll
      for (int mt = 0; mt < 3; ++mt) {
        uint32_t aa[4];
        ldAt(aa, W + WS::DQ, PD, mt * 16, 0, lane);
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t bb[4];
          ldBt(bb, W + WS::X, 200, grp * CG + np * 16, 0, lane);
          mma16816(dWacc[mt][2 * np], aa, bb[0], bb[1]);
          mma16816(dWacc[mt][2 * np + 1], aa, bb[2], bb[3]);
        }
      }
      __syncwarp();
    }
  }
  // ---- flush the per-warp accumulators
#pragma unroll
  for (int n = 0; n < 2; ++n) {
    atomicAdd(dkb + g * CPG + n * 8 + 2 * t, dkbr[n][0]); atomicAdd(dkb + g * CPG + n * 8 + 2 * t + 1, dkbr[n][1]);
    atomicAdd(dkb + (g + 8) * CPG + n * 8 + 2 * t, dkbr[n][2]); atomicAdd(dkb + (g + 8) * CPG + n * 8 + 2 * t + 1, dkbr[n][3]);
    atomicAdd(dvb + g * CPG + n * 8 + 2 * t, dvbr[n][0]); atomicAdd(dvb + g * CPG + n * 8 + 2 * t + 1, dvbr[n][1]);
    atomicAdd(dvb + (g + 8) * CPG + n * 8 + 2 * t, dvbr[n][2]); atomicAdd(dvb + (g + 8) * CPG + n * 8 + 2 * t + 1, dvbr[n][3]);
  }
#pragma unroll
  for (int mt = 0; mt < 3; ++mt) {
    float* dst = mt == 0 ? p.dWq : (mt == 1 ? p.dWk : p.dWv);
#pragma unroll
    for (int n = 0; n < 4; ++n) red_frag_v4(dst + g * CG + n * 8, dst + (g + 8) * CG + n * 8, dWacc[mt][n], t);
  }
#pragma unroll
  for (int n = 0; n < 6; ++n) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float v = dbacc[n][e];
      v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (g == 0) {
        float* dst = n < 2 ? p.dbq : (n < 4 ? p.dbk : p.dbv);
        atomicAdd(dst + (n & 1) * 8 + 2 * t + e, v);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < KB * CPG; i += blockDim.x) { atomicAdd(p.dkbp + i, dkb[i]); atomicAdd(p.dvbp + i, dvb[i]); }
}

}  // namespace

bool cga_mma_ok(const CgaP& p) {
  return p.Nt == NQ && p.cg == CG && p.cpg == CPG && p.H == NH && p.kb == KB && p.G * CG <= 192 && p.ldx % 8 == 0 &&
         p.ldo % 8 == 0 && (p.G * CG) % 8 == 0;
}

int cga_mma_fwd(cudaStream_t s, const CgaP& p) {
  if (p.B <= 0) return 0;
  const size_t smem = 3 * CPG * 4 + (size_t)(3 * CPG * PW + 2 * KB * PK + WARPS * WS::END_FWD) * 2;
  if (smem > 48 * 1024) QV_CUDA(cudaFuncSetAttribute(cga_mma_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = min(cdiv(p.B, WARPS), qv_num_sms() * 4);
  qv_launch(cga_mma_fwd_kernel, grid, WARPS * 32, smem, s, p);
  QV_LAUNCH_CHECK();
  return 0;
}

int cga_mma_bwd(cudaStream_t s, const CgaP& p) {
  if (p.B <= 0) return 0;
  QV_CHECK(p.lddo % 8 == 0 && p.lddx % 2 == 0 && (((uintptr_t)p.dWq | (uintptr_t)p.dWk | (uintptr_t)p.dWv) & 15) == 0,
           "cga_mma_bwd: unaligned gradient buffers (the dW flush uses 16 B vector reductions)");
  const size_t smem = (3 * CPG + 2 * KB * CPG) * 4 + (size_t)(3 * CPG * PW + 2 * KB * PK + WARPS_BWD * WS::END_BWD) * 2;
  QV_CUDA(cudaFuncSetAttribute(cga_mma_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = min(cdiv(p.B, WARPS_BWD), qv_num_sms() * (WARPS_BWD >= 12 ? 1 : 2));
  qv_launch(cga_mma_bwd_kernel, grid, WARPS_BWD * 32, smem, s, p);
  QV_LAUNCH_CHECK();
  return 0;
}
