// Channel-group attention (H:559-595) on mma.sync.m16n8k16 for blocks with more than 16 tokens (QAViTv2: 64 tokens per
// image, HQAViT-TinyImageNet: 64 learned tokens), bf16 runs.  Same arithmetic as cga_mma.cu, different decomposition:
// one CTA per image, warp w owns the 16-query tile w; the keys / values of ALL tokens of the image (plus the 16 projected
// bank rows) are shared through shared memory, and in backward the dK / dV contributions of the four query tiles meet
// in an fp32 shared accumulator.  Before this kernel these shapes ran the SIMT kernels of cga.cu, which were 65 % of the
// QAViTv2 training step (profiles/r1_launches_step_qavitv2_b1184.summary.txt).
#include "kernels.h"

namespace {

constexpr int CG = 32, CPG = 16, NH = 4, KB = 16;
constexpr int PW = 40;    // pitch of the stacked [48][32] projection weight
constexpr int PK = 24;    // pitch of 16-wide rows (q, k, v)
constexpr int PD = 56;    // pitch of the 48-wide [dq | dk | dv] staging
constexpr int PX = 200;   // pitch of the staged xn rows (<= 192 channels)
constexpr int PO = 104;   // pitch of the staged dO rows (<= 96 channels)
constexpr int WARPS = 4;

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

// T = query tiles per image (Nt = 16 T); keys = Nt own tokens then the 16 bank rows
template <int T>
struct Lay {
  static constexpr int NT = 16 * T, NKEY = NT + KB, NT8 = NKEY / 8, KS = NKEY / 16, PP = NKEY + 8;
  // CTA-shared bf16 regions
  static constexpr int X = 0, K = X + NT * PX, V = K + NKEY * PK, DO = V + NKEY * PK, SH_FWD = DO, SH_BWD = DO + NT * PO;
  // per-warp bf16 regions
  static constexpr int Q = 0, P = Q + 16 * PK, DS = P + 16 * PP, DQ = DS + 16 * PP, W_FWD = 0, W_BWD = DQ + 16 * PD;
};

__device__ __forceinline__ void load_consts(const CgaP& p, bf16* Wst, float* bias, bf16* Ks, bf16* Vs, int nt) {
  for (int i = threadIdx.x; i < 3 * CPG * CG; i += blockDim.x) {
    const int which = i / (CPG * CG), r = i % (CPG * CG);
    const float* src = which == 0 ? p.Wq : (which == 1 ? p.Wk : p.Wv);
    Wst[(which * CPG + r / CG) * PW + r % CG] = __float2bfloat16_rn(src[r]);
  }
  for (int i = threadIdx.x; i < 3 * CPG; i += blockDim.x) bias[i] = i < CPG ? p.bq[i] : (i < 2 * CPG ? p.bk[i - CPG] : p.bv[i - 2 * CPG]);
  for (int i = threadIdx.x; i < KB * CPG; i += blockDim.x) {      // bank rows sit behind the image's own keys, for good
    Ks[(nt + i / CPG) * PK + i % CPG] = __float2bfloat16_rn(p.kbp[i]);
    Vs[(nt + i / CPG) * PK + i % CPG] = __float2bfloat16_rn(p.vbp[i]);
  }
}

// q | k | v = x_g W^T + b for this warp's 16 rows (C layout: acc[0..1] = q, [2..3] = k, [4..5] = v); k, v -> shared rows
__device__ __forceinline__ void project(float (*acc)[4], const bf16* Xs, bf16* Ks, bf16* Vs, bf16* Qw, const bf16* Wst, const float* bias,
                                        int grp, int row0, int lane) {
  const int t = lane & 3;
#pragma unroll
  for (int n = 0; n < 6; ++n) {
    acc[n][0] = acc[n][2] = bias[n * 8 + 2 * t];
    acc[n][1] = acc[n][3] = bias[n * 8 + 2 * t + 1];
  }
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    uint32_t a[4];
    ldA(a, Xs, PX, row0, grp * CG + kk * 16, lane);
#pragma unroll
    for (int np = 0; np < 3; ++np) {
      uint32_t b[4];
      ldB(b, Wst, PW, np * 16, kk * 16, lane);
      mma16816(acc[2 * np], a, b[0], b[1]);
      mma16816(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
  if (Qw) { stC(Qw, PK, 0, 0, acc[0], lane); stC(Qw, PK, 0, 8, acc[1], lane); }
  stC(Ks, PK, row0, 0, acc[2], lane); stC(Ks, PK, row0, 8, acc[3], lane);
  stC(Vs, PK, row0, 0, acc[4], lane); stC(Vs, PK, row0, 8, acc[5], lane);
}

// softmax(q_h K^T / 2) over all keys, C layout s[NT8][4]
template <int NT8>
__device__ __forceinline__ void head_scores(float (*s)[4], const float (*q)[4], const bf16* Ks, int h, int lane) {
  uint32_t a[4];
  head_frag(a, q, h, lane);
#pragma unroll
  for (int n = 0; n < NT8; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
  for (int np = 0; np < NT8 / 2; ++np) {
    uint32_t b[4];
    ldB(b, Ks, PK, np * 16, 0, lane);
    mma16816(s[2 * np], a, b[0], b[1]);
    mma16816(s[2 * np + 1], a, b[2], b[3]);
  }
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int n = 0; n < NT8; ++n) {
    s[n][0] *= 0.5f; s[n][1] *= 0.5f; s[n][2] *= 0.5f; s[n][3] *= 0.5f;
    m0 = fmaxf(m0, fmaxf(s[n][0], s[n][1]));
    m1 = fmaxf(m1, fmaxf(s[n][2], s[n][3]));
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  float z0 = 0.f, z1 = 0.f;
#pragma unroll
  for (int n = 0; n < NT8; ++n) {
    s[n][0] = __expf(s[n][0] - m0); s[n][1] = __expf(s[n][1] - m0);
    s[n][2] = __expf(s[n][2] - m1); s[n][3] = __expf(s[n][3] - m1);
    z0 += s[n][0] + s[n][1];
    z1 += s[n][2] + s[n][3];
  }
  z0 += __shfl_xor_sync(0xffffffffu, z0, 1); z0 += __shfl_xor_sync(0xffffffffu, z0, 2);
  z1 += __shfl_xor_sync(0xffffffffu, z1, 1); z1 += __shfl_xor_sync(0xffffffffu, z1, 2);
  z0 = 1.f / z0; z1 = 1.f / z1;
#pragma unroll
  for (int n = 0; n < NT8; ++n) { s[n][0] *= z0; s[n][1] *= z0; s[n][2] *= z1; s[n][3] *= z1; }
}

// o[2][4] = X[16 x NKEY] (C-layout regs as A) * Bs (smem [key][dim])
template <int KS>
__device__ __forceinline__ void keys_times(float (*o)[4], const float (*x)[4], const bf16* Bs, int lane) {
#pragma unroll
  for (int n = 0; n < 2; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < KS; ++kk) {
    uint32_t a[4] = {pack2(x[2 * kk][0], x[2 * kk][1]), pack2(x[2 * kk][2], x[2 * kk][3]),
                     pack2(x[2 * kk + 1][0], x[2 * kk + 1][1]), pack2(x[2 * kk + 1][2], x[2 * kk + 1][3])};
    uint32_t b[4];
    ldBt(b, Bs, PK, 0, kk * 16, lane);
    mma16816(o[0], a, b[0], b[1]);
    mma16816(o[1], a, b[2], b[3]);
  }
}

template <int T>
__device__ __forceinline__ void load_rows(bf16* dst, int pitch, const bf16* src, long row0, int ld, int ncols) {
  const int cpr = ncols / 8;
  for (int c = threadIdx.x; c < 16 * T * cpr; c += blockDim.x) {
    const int i = c / cpr, ch = c % cpr;
    *reinterpret_cast<uint4*>(dst + i * pitch + ch * 8) = *reinterpret_cast<const uint4*>(src + (row0 + i) * ld + ch * 8);
  }
}

template <int T>
__global__ void __launch_bounds__(WARPS * 32) cga64_fwd_kernel(CgaP p) {
  QV_PDL_ENTRY();
  static_assert(T <= WARPS, "one query tile per warp");
  using L = Lay<T>;
  extern __shared__ __align__(16) uint8_t smraw[];
  float* bias = reinterpret_cast<float*>(smraw);                    // [48]
  bf16* Wst = reinterpret_cast<bf16*>(bias + 3 * CPG);              // [48][PW]
  bf16* SH = Wst + 3 * CPG * PW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  load_consts(p, Wst, bias, SH + L::K, SH + L::V, L::NT);
  bf16* out = static_cast<bf16*>(p.out);
  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    __syncthreads();
    load_rows<T>(SH + L::X, PX, static_cast<const bf16*>(p.xn), (long)b * L::NT, p.ldx, p.G * CG);
    for (int grp = 0; grp < p.G; ++grp) {
      __syncthreads();                                              // X staged / previous group's K, V consumed
      const int tile = warp;
      float acc[6][4];
      if (tile < T) project(acc, SH + L::X, SH + L::K, SH + L::V, nullptr, Wst, bias, grp, tile * 16, lane);
      __syncthreads();                                              // all keys / values of the image are in place
      if (tile < T) {
        float o[2][4];
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          float s[L::NT8][4], o2[2][4];
          head_scores<L::NT8>(s, acc, SH + L::K, h, lane);
          if (p.drop.p > 0.f) {   // SDPA dropout_p on the probabilities
            const DropState ds = drop_state(p.drop);
            drop_apply_c<L::NT8>(s, drop_bits_c<L::NT8>(ds, (uint32_t)(((b * T + tile) * p.G + grp) * NH + h), lane), ds.inv);
          }
          keys_times<L::KS>(o2, s, SH + L::V, lane);
          head_keep(o, o2, h, lane, false);
        }
        const long r0 = (long)b * L::NT + tile * 16 + g, r1 = r0 + 8;
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          *reinterpret_cast<uint32_t*>(out + r0 * p.ldo + grp * CPG + n * 8 + 2 * t) = pack2(o[n][0], o[n][1]);
          *reinterpret_cast<uint32_t*>(out + r1 * p.ldo + grp * CPG + n * 8 + 2 * t) = pack2(o[n][2], o[n][3]);
        }
      }
    }
  }
}

template <int T>
__global__ void __launch_bounds__(WARPS * 32) cga64_bwd_kernel(CgaP p) {
  QV_PDL_ENTRY();
  static_assert(T <= WARPS, "one query tile per warp");
  using L = Lay<T>;
  extern __shared__ __align__(16) uint8_t smraw[];
  float* bias = reinterpret_cast<float*>(smraw);                    // [48]
  float* dkb = bias + 3 * CPG;                                      // [16][16] d(projected bank k), then v
  float* dvb = dkb + KB * CPG;
  float* dKs = dvb + KB * CPG;                                      // [NKEY][16] fp32 accumulators of the current group
  float* dVs = dKs + L::NKEY * CPG;
  bf16* Wst = reinterpret_cast<bf16*>(dVs + L::NKEY * CPG);
  bf16* SH = Wst + 3 * CPG * PW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  bf16* W = SH + L::SH_BWD + warp * L::W_BWD;
  const bool active = warp < T;
  const int row0 = warp * 16;
  load_consts(p, Wst, bias, SH + L::K, SH + L::V, L::NT);
  for (int i = threadIdx.x; i < 2 * KB * CPG; i += blockDim.x) dkb[i] = 0.f;
  float dWacc[3][4][4], dbacc[6][2];
#pragma unroll
  for (int m = 0; m < 3; ++m)
#pragma unroll
    for (int n = 0; n < 4; ++n) dWacc[m][n][0] = dWacc[m][n][1] = dWacc[m][n][2] = dWacc[m][n][3] = 0.f;
#pragma unroll
  for (int n = 0; n < 6; ++n) dbacc[n][0] = dbacc[n][1] = 0.f;

  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    __syncthreads();
    load_rows<T>(SH + L::X, PX, static_cast<const bf16*>(p.xn), (long)b * L::NT, p.ldx, p.G * CG);
    load_rows<T>(SH + L::DO, PO, static_cast<const bf16*>(p.dout), (long)b * L::NT, p.lddo, p.G * CPG);
    for (int grp = 0; grp < p.G; ++grp) {
      __syncthreads();                                              // tiles staged / previous group fully consumed
      for (int i = threadIdx.x; i < 2 * L::NKEY * CPG; i += blockDim.x) dKs[i] = 0.f;
      float acc[6][4];
      if (active) project(acc, SH + L::X, SH + L::K, SH + L::V, W + L::Q, Wst, bias, grp, row0, lane);
      __syncthreads();
      float dq[2][4];
      float dk[L::KS][2][4], dv[L::KS][2][4];
      if (active) {
        float dO[2][4];
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(SH + L::DO + (row0 + g) * PO + grp * CPG + n * 8 + 2 * t));
          const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(SH + L::DO + (row0 + g + 8) * PO + grp * CPG + n * 8 + 2 * t));
          dO[n][0] = lo.x; dO[n][1] = lo.y; dO[n][2] = hi.x; dO[n][3] = hi.y;
        }
#pragma unroll
        for (int n = 0; n < 2; ++n)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            dq[n][e] = 0.f;
#pragma unroll
            for (int m = 0; m < L::KS; ++m) dk[m][n][e] = dv[m][n][e] = 0.f;
          }
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          float P[L::NT8][4], dS[L::NT8][4];
          head_scores<L::NT8>(P, acc, SH + L::K, h, lane);
          uint32_t a[4];
          head_frag(a, dO, h, lane);                                // dP = dO_h V^T
#pragma unroll
          for (int n = 0; n < L::NT8; ++n) dS[n][0] = dS[n][1] = dS[n][2] = dS[n][3] = 0.f;
#pragma unroll
          for (int np = 0; np < L::NT8 / 2; ++np) {
            uint32_t bb[4];
            ldB(bb, SH + L::V, PK, np * 16, 0, lane);
            mma16816(dS[2 * np], a, bb[0], bb[1]);
            mma16816(dS[2 * np + 1], a, bb[2], bb[3]);
          }
          unsigned long long keep = ~0ull;
          float kinv = 1.f;
          if (p.drop.p > 0.f) {   // dP <- d(P_dropped) * keep; P_dropped (for dv) is formed after the softmax backward
            const DropState ds = drop_state(p.drop);
            keep = drop_bits_c<L::NT8>(ds, (uint32_t)(((b * T + warp) * p.G + grp) * NH + h), lane);
            kinv = ds.inv;
            drop_apply_c<L::NT8>(dS, keep, kinv);
          }
          float r0 = 0.f, r1 = 0.f;
#pragma unroll
          for (int n = 0; n < L::NT8; ++n) { r0 += dS[n][0] * P[n][0] + dS[n][1] * P[n][1]; r1 += dS[n][2] * P[n][2] + dS[n][3] * P[n][3]; }
          r0 += __shfl_xor_sync(0xffffffffu, r0, 1); r0 += __shfl_xor_sync(0xffffffffu, r0, 2);
          r1 += __shfl_xor_sync(0xffffffffu, r1, 1); r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
#pragma unroll
          for (int n = 0; n < L::NT8; ++n) {
            dS[n][0] = P[n][0] * (dS[n][0] - r0) * 0.5f; dS[n][1] = P[n][1] * (dS[n][1] - r0) * 0.5f;
            dS[n][2] = P[n][2] * (dS[n][2] - r1) * 0.5f; dS[n][3] = P[n][3] * (dS[n][3] - r1) * 0.5f;
          }
          if (p.drop.p > 0.f) drop_apply_c<L::NT8>(P, keep, kinv);
#pragma unroll
          for (int n = 0; n < L::NT8; ++n) {
            stC(W + L::P, L::PP, 0, n * 8, P[n], lane);
            stC(W + L::DS, L::PP, 0, n * 8, dS[n], lane);
          }
          float t2[2][4];
          keys_times<L::KS>(t2, dS, SH + L::K, lane);               // dq_h = dS K
          head_keep(dq, t2, h, lane, true);
          __syncwarp();
          // dk_h = dS^T q, dv_h = P^T dO   (M = keys, N = 16 channels, K = this warp's 16 queries)
          uint32_t bq[4], bo[4];
          ldBt(bq, W + L::Q, PK, 0, 0, lane);
          ldBt(bo, SH + L::DO, PO, grp * CPG, row0, lane);
#pragma unroll
          for (int mt = 0; mt < L::KS; ++mt) {
            uint32_t aa[4];
            float ck[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, cv[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
            ldAt(aa, W + L::DS, L::PP, mt * 16, 0, lane);
            mma16816(ck[0], aa, bq[0], bq[1]);
            mma16816(ck[1], aa, bq[2], bq[3]);
            ldAt(aa, W + L::P, L::PP, mt * 16, 0, lane);
            mma16816(cv[0], aa, bo[0], bo[1]);
            mma16816(cv[1], aa, bo[2], bo[3]);
            head_keep(dk[mt], ck, h, lane, true);
            head_keep(dv[mt], cv, h, lane, true);
          }
          __syncwarp();
        }
      }
      // this query tile's contribution to every key of the image -> shared fp32 accumulators, one warp after the other with plain
      // read-modify-writes.  (fp32 atomicAdd on shared memory compiles to a compare-and-swap loop -- ATOMS.CAST.SPIN: 80 of them per
      // lane and group, with all four warps on the same addresses.)  The fixed order also makes the sums reproducible.
      for (int w = 0; w < WARPS; ++w) {
        if (warp == w && active) {
#pragma unroll
          for (int mt = 0; mt < L::KS; ++mt)
#pragma unroll
            for (int n = 0; n < 2; ++n) {
              const int c = n * 8 + 2 * t, ra = (mt * 16 + g) * CPG, rb = (mt * 16 + g + 8) * CPG;
              float2* ka = reinterpret_cast<float2*>(dKs + ra + c); float2* kb2 = reinterpret_cast<float2*>(dKs + rb + c);
              float2* va = reinterpret_cast<float2*>(dVs + ra + c); float2* vb2 = reinterpret_cast<float2*>(dVs + rb + c);
              float2 x;
              x = *ka; x.x += dk[mt][n][0]; x.y += dk[mt][n][1]; *ka = x;
              x = *kb2; x.x += dk[mt][n][2]; x.y += dk[mt][n][3]; *kb2 = x;
              x = *va; x.x += dv[mt][n][0]; x.y += dv[mt][n][1]; *va = x;
              x = *vb2; x.x += dv[mt][n][2]; x.y += dv[mt][n][3]; *vb2 = x;
            }
        }
        __syncthreads();                                            // after the last pass: dK / dV of the group are complete
      }
      for (int i = threadIdx.x; i < KB * CPG; i += blockDim.x) {   // bank rows -> d(projected bank)
        dkb[i] += dKs[L::NT * CPG + i];
        dvb[i] += dVs[L::NT * CPG + i];
      }
      if (active) {
        float dk0[2][4], dv0[2][4];                                 // this warp's own rows, back in C layout
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          const int c = n * 8 + 2 * t, ra = (row0 + g) * CPG, rb = (row0 + g + 8) * CPG;
          dk0[n][0] = dKs[ra + c]; dk0[n][1] = dKs[ra + c + 1]; dk0[n][2] = dKs[rb + c]; dk0[n][3] = dKs[rb + c + 1];
          dv0[n][0] = dVs[ra + c]; dv0[n][1] = dVs[ra + c + 1]; dv0[n][2] = dVs[rb + c]; dv0[n][3] = dVs[rb + c + 1];
        }
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          dbacc[n][0] += dq[n][0] + dq[n][2]; dbacc[n][1] += dq[n][1] + dq[n][3];
          dbacc[2 + n][0] += dk0[n][0] + dk0[n][2]; dbacc[2 + n][1] += dk0[n][1] + dk0[n][3];
          dbacc[4 + n][0] += dv0[n][0] + dv0[n][2]; dbacc[4 + n][1] += dv0[n][1] + dv0[n][3];
          stC(W + L::DQ, PD, 0, n * 8, dq[n], lane);
          stC(W + L::DQ, PD, 0, 16 + n * 8, dk0[n], lane);
          stC(W + L::DQ, PD, 0, 32 + n * 8, dv0[n], lane);
        }
        __syncwarp();
        // dx_g[16 x 32] = [dq|dk|dv] Wst
        {
          float dx[4][4];
#pragma unroll
          for (int n = 0; n < 4; ++n) dx[n][0] = dx[n][1] = dx[n][2] = dx[n][3] = 0.f;
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) {
            uint32_t aa[4];
            ldA(aa, W + L::DQ, PD, 0, kk * 16, lane);
#pragma unroll
            for (int np = 0; np < 2; ++np) {
              uint32_t bb[4];
              ldBt(bb, Wst, PW, np * 16, kk * 16, lane);
              mma16816(dx[2 * np], aa, bb[0], bb[1]);
              mma16816(dx[2 * np + 1], aa, bb[2], bb[3]);
            }
          }
          float* d0 = p.dxn + ((long)b * L::NT + row0 + g) * p.lddx + grp * CG + 2 * t;
          float* d1 = d0 + 8L * p.lddx;
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            float2 u = *reinterpret_cast<float2*>(d0 + n * 8), v = *reinterpret_cast<float2*>(d1 + n * 8);
            u.x += dx[n][0]; u.y += dx[n][1]; v.x += dx[n][2]; v.y += dx[n][3];
            *reinterpret_cast<float2*>(d0 + n * 8) = u;
            *reinterpret_cast<float2*>(d1 + n * 8) = v;
          }
        }
        // dW[o, c] += sum_n dqkv[n, o] x_g[n, c]
#pragma unroll
        for (int mt = 0; mt < 3; ++mt) {
          uint32_t aa[4];
          ldAt(aa, W + L::DQ, PD, mt * 16, 0, lane);
#pragma unroll
          for (int np = 0; np < 2; ++np) {
            uint32_t bb[4];
            ldBt(bb, SH + L::X, PX, grp * CG + np * 16, row0, lane);
            mma16816(dWacc[mt][2 * np], aa, bb[0], bb[1]);
            mma16816(dWacc[mt][2 * np + 1], aa, bb[2], bb[3]);
          }
        }
      }
    }
  }
  // ---- flush the per-warp accumulators
  if (active) {
#pragma unroll
    for (int mt = 0; mt < 3; ++mt) {
      float* dst = mt == 0 ? p.dWq : (mt == 1 ? p.dWk : p.dWv);
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        atomicAdd(dst + g * CG + n * 8 + 2 * t, dWacc[mt][n][0]); atomicAdd(dst + g * CG + n * 8 + 2 * t + 1, dWacc[mt][n][1]);
        atomicAdd(dst + (g + 8) * CG + n * 8 + 2 * t, dWacc[mt][n][2]); atomicAdd(dst + (g + 8) * CG + n * 8 + 2 * t + 1, dWacc[mt][n][3]);
      }
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
  }
  __syncthreads();
  for (int i = threadIdx.x; i < KB * CPG; i += blockDim.x) { atomicAdd(p.dkbp + i, dkb[i]); atomicAdd(p.dvbp + i, dvb[i]); }
}

template <int T>
size_t smem_fwd() {
  using L = Lay<T>;
  return 3 * CPG * 4 + (size_t)(3 * CPG * PW + L::SH_FWD) * 2;
}
template <int T>
size_t smem_bwd() {
  using L = Lay<T>;
  return (3 * CPG + 2 * KB * CPG + 2 * L::NKEY * CPG) * 4 + (size_t)(3 * CPG * PW + L::SH_BWD + WARPS * L::W_BWD) * 2;
}

}  // namespace

bool cga_mma64_ok(const CgaP& p) {
  return p.Nt == 64 && p.cg == CG && p.cpg == CPG && p.H == NH && p.kb == KB && p.G * CG <= 192 && p.ldx % 8 == 0 &&
         p.ldo % 8 == 0 && (p.G * CG) % 8 == 0 && p.G * CPG <= 96;
}

int cga_mma64_fwd(cudaStream_t s, const CgaP& p) {
  if (p.B <= 0) return 0;
  const size_t smem = smem_fwd<4>();
  QV_CUDA(cudaFuncSetAttribute(cga64_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int occ = max(1, min(6, (int)(200 * 1024 / (smem + 1024))));
  const int grid = min(p.B, qv_num_sms() * occ);
  qv_launch(cga64_fwd_kernel<4>, grid, WARPS * 32, smem, s, p);
  QV_LAUNCH_CHECK();
  return 0;
}

int cga_mma64_bwd(cudaStream_t s, const CgaP& p) {
  if (p.B <= 0) return 0;
  const size_t smem = smem_bwd<4>();
  QV_CHECK(smem <= 200 * 1024, "cga64 backward needs %zu B of shared memory", smem);
  QV_CUDA(cudaFuncSetAttribute(cga64_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int occ = max(1, min(4, (int)(200 * 1024 / (smem + 1024))));
  const int grid = min(p.B, qv_num_sms() * occ);
  qv_launch(cga64_bwd_kernel<4>, grid, WARPS * 32, smem, s, p);
  QV_LAUNCH_CHECK();
  return 0;
}
