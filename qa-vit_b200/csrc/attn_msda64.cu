// Multi-scale dilated attention (H:496-532) on mma.sync for blocks with more than 16 query tokens per image (QAViTv2:
// 64 tokens, HQAViT-TinyImageNet: 64 learned tokens), bf16 runs.  The 16-query kernel of attn_mma.cu keeps the Linformer
// contraction inside the attention kernel; here an image has several query tiles that share the compressed keys, so
// the contraction is hoisted:
//   lin_pre   : K' = E_k^T Ks, V' = E_v^T Vs per image          [B, 32, 2D] bf16           (tiny batched product, SIMT)
//   attention : one warp per (image, head); Kf = [K'_h ; bank_k_h] (48 keys) staged once, then a loop over the image's
//               16-query tiles: S = Q Kf^T, softmax, O = P Vf on m16n8k16; backward accumulates dK' / dV' of all tiles
//               in warp-private fp32 shared memory (no atomics), bank rows in registers across tasks
//   lin_post  : dKs = E_k dK', dVs = E_v dV' (-> d_kv), dE_k += Ks dK'^T, dE_v += Vs dV'^T
// Before this file these shapes ran attn_fwd/bwd_kernel (SIMT), 57 % of the QAViTv2 step after CGA moved to tensor cores.
#include "kernels.h"

namespace {

constexpr int HD = 48, NQ = 16, KLIN = 32, KB = 16, NKV = KLIN + KB;
constexpr int PT = 56;    // pitch (bf16) of 48-wide rows
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
__device__ __forceinline__ void ldA(uint32_t* a, const bf16* base, int pitch, int m0, int k0, int lane) {
  const int mat = lane >> 3, r = lane & 7;
  ldsm4(a, sa(base + (m0 + r + (mat & 1) * 8) * pitch + k0 + (mat >> 1) * 8));
}
__device__ __forceinline__ void ldAt(uint32_t* a, const bf16* base, int pitch, int m0, int k0, int lane) {
  const int mat = lane >> 3, r = lane & 7;
  ldsm4t(a, sa(base + (k0 + r + (mat >> 1) * 8) * pitch + m0 + (mat & 1) * 8));
}
__device__ __forceinline__ void ldB(uint32_t* b, const bf16* base, int pitch, int n0, int k0, int lane) {
  const int mat = lane >> 3, r = lane & 7;
  ldsm4(b, sa(base + (n0 + r + (mat >> 1) * 8) * pitch + k0 + (mat & 1) * 8));
}
__device__ __forceinline__ void ldBt(uint32_t* b, const bf16* base, int pitch, int n0, int k0, int lane) {
  const int mat = lane >> 3, r = lane & 7;
  ldsm4t(b, sa(base + (k0 + r + (mat & 1) * 8) * pitch + n0 + (mat >> 1) * 8));
}
__device__ __forceinline__ void stC(bf16* base, int pitch, int m0, int n0, const float* c, int lane) {
  const int g = lane >> 2, t = lane & 3;
  *reinterpret_cast<uint32_t*>(base + (m0 + g) * pitch + n0 + 2 * t) = pack2(c[0], c[1]);
  *reinterpret_cast<uint32_t*>(base + (m0 + g + 8) * pitch + n0 + 2 * t) = pack2(c[2], c[3]);
}

// ------------------------------------------------------------------------------------------------ Linformer pre / post
// pre[b, j, c] = sum_l E[l, j] kv[b * NM + l, c]   (E = E_k for c < D, E_v for c >= D; l < L)
__global__ void __launch_bounds__(256) lin_pre_kernel(const bf16* __restrict__ kv, int ldkv, int kcol, int vcol, int NM, int L, int D,
                                                      const float* __restrict__ Ek, const float* __restrict__ Ev, int B,
                                                      bf16* __restrict__ pre) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) float sm[];
  float* sE = sm;                        // [2][L][32]
  float* sX = sE + 2 * L * KLIN;         // [L][2D]
  for (int i = threadIdx.x; i < 2 * L * KLIN; i += blockDim.x) sE[i] = i < L * KLIN ? Ek[i] : Ev[i - L * KLIN];
  const int D2 = 2 * D;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < L * D; i += blockDim.x) {            // pairs of channels
      const int l = i / D, c2 = (i % D) * 2;                           // c2 in [0, 2D)
      const int col = c2 < D ? kcol + c2 : vcol + (c2 - D);
      const float2 v = ld2(kv + ((long)b * NM + l) * ldkv + col);
      *reinterpret_cast<float2*>(sX + l * D2 + c2) = v;
    }
    __syncthreads();
    // thread block of 4 (j) x 2 (c) outputs
    for (int i = threadIdx.x; i < (KLIN / 4) * D; i += blockDim.x) {
      const int jq = i / D, c2 = (i % D) * 2;
      const float* E = sE + (c2 < D ? 0 : L * KLIN) + jq * 4;
      float a[4][2] = {};
      for (int l = 0; l < L; ++l) {
        const float4 e = *reinterpret_cast<const float4*>(E + l * KLIN);
        const float2 xv = *reinterpret_cast<const float2*>(sX + l * D2 + c2);
        const float x0 = xv.x, x1 = xv.y;
        a[0][0] = fmaf(e.x, x0, a[0][0]); a[0][1] = fmaf(e.x, x1, a[0][1]);
        a[1][0] = fmaf(e.y, x0, a[1][0]); a[1][1] = fmaf(e.y, x1, a[1][1]);
        a[2][0] = fmaf(e.z, x0, a[2][0]); a[2][1] = fmaf(e.z, x1, a[2][1]);
        a[3][0] = fmaf(e.w, x0, a[3][0]); a[3][1] = fmaf(e.w, x1, a[3][1]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint32_t*>(pre + ((long)b * KLIN + jq * 4 + q) * D2 + c2) = pack2(a[q][0], a[q][1]);
    }
  }
}

// dkv[b * NM + l, c] = sum_j E[l, j] dpre[b, j, c];  dE[l, j] += sum_{b, c in half} kv[b * NM + l, c] dpre[b, j, c]
__global__ void __launch_bounds__(256) lin_post_kernel(const bf16* __restrict__ kv, int ldkv, int kcol, int vcol, int NM, int L, int D,
                                                       const float* __restrict__ Ek, const float* __restrict__ Ev, int B,
                                                       const float* __restrict__ dpre, bf16* __restrict__ dkv, int lddkv, int dkcol,
                                                       int dvcol, float* __restrict__ dEk, float* __restrict__ dEv) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) float sm[];
  const int D2 = 2 * D;
  float* sE = sm;                        // [2][L][32]
  float* sX = sE + 2 * L * KLIN;         // [L][2D]
  float* sG = sX + L * D2;               // [32][2D + 4]: in the dE product a warp's lanes read 32 different rows j at the same
  const int GP = D2 + 4;                 // column, so the pitch must not be a multiple of 32 banks (2D = 384 was a 32-way conflict)
  for (int i = threadIdx.x; i < 2 * L * KLIN; i += blockDim.x) sE[i] = i < L * KLIN ? Ek[i] : Ev[i - L * KLIN];
  // dE outputs owned by this thread: (which, l, j) = idx, idx + 256, ...  (2 * L * 32 of them)
  constexpr int MAXO = 32;               // 2 * 128 * 32 / 256
  float accE[MAXO];
#pragma unroll
  for (int q = 0; q < MAXO; ++q) accE[q] = 0.f;
  const int nout = 2 * L * KLIN;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < L * D; i += blockDim.x) {
      const int l = i / D, c2 = (i % D) * 2;
      const int col = c2 < D ? kcol + c2 : vcol + (c2 - D);
      const float2 v = ld2(kv + ((long)b * NM + l) * ldkv + col);
      *reinterpret_cast<float2*>(sX + l * D2 + c2) = v;
    }
    for (int i = threadIdx.x; i < KLIN * D2 / 4; i += blockDim.x) {
      const int j = (i * 4) / D2, c = (i * 4) % D2;
      *reinterpret_cast<float4*>(sG + j * GP + c) = *reinterpret_cast<const float4*>(dpre + (long)b * KLIN * D2 + i * 4);
    }
    __syncthreads();
    // dKs / dVs
    for (int i = threadIdx.x; i < L * D; i += blockDim.x) {
      const int l = i / D, c2 = (i % D) * 2;
      const float* E = sE + (c2 < D ? 0 : L * KLIN) + l * KLIN;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
      for (int j = 0; j < KLIN; ++j) {
        const float2 gv = *reinterpret_cast<const float2*>(sG + j * GP + c2);
        a0 = fmaf(E[j], gv.x, a0); a1 = fmaf(E[j], gv.y, a1);
      }
      const int col = c2 < D ? dkcol + c2 : dvcol + (c2 - D);
      *reinterpret_cast<uint32_t*>(dkv + ((long)b * NM + l) * lddkv + col) = pack2(a0, a1);
    }
    // dE partials
#pragma unroll
    for (int q = 0; q < MAXO; ++q) {
      const int o = threadIdx.x + q * 256;
      if (o < nout) {
        const int which = o / (L * KLIN), r = o % (L * KLIN), l = r / KLIN, j = r % KLIN;
        const float* xr = sX + l * D2 + which * D;
        const float* gr = sG + j * GP + which * D;
        float a = 0.f;
        for (int c = 0; c < D; c += 4) {
          const float4 x = *reinterpret_cast<const float4*>(xr + c), gq = *reinterpret_cast<const float4*>(gr + c);
          a = fmaf(x.x, gq.x, fmaf(x.y, gq.y, fmaf(x.z, gq.z, fmaf(x.w, gq.w, a))));
        }
        accE[q] += a;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < MAXO; ++q) {
    const int o = threadIdx.x + q * 256;
    if (o < nout) atomicAdd((o < L * KLIN ? dEk : dEv - L * KLIN) + o, accE[q]);
  }
}

// ------------------------------------------------------------------------------------------------ tensor-core pre / post
// Same contractions on mma.sync for L <= 48 pooled rows (every shipped 64-token configuration has L = 40): the SIMT
// kernels above were shared-memory bound (lin_post 17 % of the QAViTv2 step).  One CTA (4 warps) per image, grid-stride;
// warp w owns columns [96 w, 96 w + 96) of the [K | V] = 2 D = 384 wide matrices, i.e. warps 0-1 the K half, 2-3 the V half.
constexpr int LPAD = 48;              // pooled rows padded to a multiple of 16 (zero rows)
constexpr int PX = 2 * WARPS * HD + 8;   // pitch (bf16) of the 384-wide staged matrices: 784 B rows, 16 B aligned, 4-bank skew
constexpr int PEL = 40;               // pitch of the staged E matrices [LPAD][32]
constexpr int WCOLS = 96;             // columns per warp

struct LinSmem {
  static constexpr int E = 0, X = E + 2 * LPAD * PEL, G = X + LPAD * PX, END_PRE = G, END_POST = G + KLIN * PX;
};

__device__ __forceinline__ void lin_stage_E(bf16* sE, const float* Ek, const float* Ev, int L) {
  for (int i = threadIdx.x; i < 2 * LPAD * KLIN; i += blockDim.x) {
    const int which = i / (LPAD * KLIN), r = i % (LPAD * KLIN), l = r / KLIN, j = r % KLIN;
    sE[which * LPAD * PEL + l * PEL + j] = __float2bfloat16_rn(l < L ? (which ? Ev : Ek)[l * KLIN + j] : 0.f);
  }
}
// Ks | Vs rows of image b -> [LPAD][PX] (rows >= L zero)
__device__ __forceinline__ void lin_stage_X(bf16* sX, const bf16* kv, int ldkv, int kcol, int vcol, int NM, int L, int D, int b) {
  const int cpr = 2 * D / 8;
  for (int i = threadIdx.x; i < LPAD * cpr; i += blockDim.x) {
    const int l = i / cpr, c = (i % cpr) * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (l < L) v = *reinterpret_cast<const uint4*>(kv + ((long)b * NM + l) * ldkv + (c < D ? kcol + c : vcol + (c - D)));
    *reinterpret_cast<uint4*>(sX + l * PX + c) = v;
  }
}

// pre[b, j, c] = sum_l E[l, j] X[l, c]
__global__ void __launch_bounds__(WARPS * 32) lin_pre_mma_kernel(const bf16* __restrict__ kv, int ldkv, int kcol, int vcol, int NM, int L,
                                                                 const float* __restrict__ Ek, const float* __restrict__ Ev, int B,
                                                                 bf16* __restrict__ pre) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  bf16* sm = reinterpret_cast<bf16*>(smraw);
  bf16 *sE = sm + LinSmem::E, *sX = sm + LinSmem::X;
  constexpr int D = WARPS * HD, D2 = 2 * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int c0 = warp * WCOLS;
  const bf16* E = sE + (c0 >= D ? LPAD * PEL : 0);
  lin_stage_E(sE, Ek, Ev, L);
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    lin_stage_X(sX, kv, ldkv, kcol, vcol, NM, L, D, b);
    __syncthreads();
#pragma unroll
    for (int mt = 0; mt < KLIN / 16; ++mt) {
      float acc[WCOLS / 8][4];
#pragma unroll
      for (int n = 0; n < WCOLS / 8; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < LPAD / 16; ++kk) {
        uint32_t a[4];
        ldAt(a, E, PEL, mt * 16, kk * 16, lane);                       // A(m = j, k = l) = E[l][j]
#pragma unroll
        for (int np = 0; np < WCOLS / 16; ++np) {
          uint32_t bb[4];
          ldBt(bb, sX, PX, c0 + np * 16, kk * 16, lane);               // B(k = l, n = c) = X[l][c]
          mma16816(acc[2 * np], a, bb[0], bb[1]);
          mma16816(acc[2 * np + 1], a, bb[2], bb[3]);
        }
      }
      bf16* dst = pre + ((long)b * KLIN + mt * 16) * D2 + c0;
#pragma unroll
      for (int n = 0; n < WCOLS / 8; ++n) {
        *reinterpret_cast<uint32_t*>(dst + (long)g * D2 + n * 8 + 2 * t) = pack2(acc[n][0], acc[n][1]);
        *reinterpret_cast<uint32_t*>(dst + (long)(g + 8) * D2 + n * 8 + 2 * t) = pack2(acc[n][2], acc[n][3]);
      }
    }
  }
}

// dkv[b * NM + l, c] = sum_j E[l, j] dpre[b, j, c];  dE[l, j] += sum_{b, c in half} X[l, c] dpre[b, j, c]
__global__ void __launch_bounds__(WARPS * 32) lin_post_mma_kernel(const bf16* __restrict__ kv, int ldkv, int kcol, int vcol, int NM, int L,
                                                                  const float* __restrict__ Ek, const float* __restrict__ Ev, int B,
                                                                  const float* __restrict__ dpre, bf16* __restrict__ dkv, int lddkv,
                                                                  int dkcol, int dvcol, float* __restrict__ dEk, float* __restrict__ dEv) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  bf16* sm = reinterpret_cast<bf16*>(smraw);
  bf16 *sE = sm + LinSmem::E, *sX = sm + LinSmem::X, *sG = sm + LinSmem::G;
  constexpr int D = WARPS * HD, D2 = 2 * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int c0 = warp * WCOLS, which = c0 >= D ? 1 : 0;
  const bf16* E = sE + which * LPAD * PEL;
  lin_stage_E(sE, Ek, Ev, L);
  float dEacc[LPAD / 16][KLIN / 8][4];   // this warp's share of dE_k (warps 0-1) or dE_v (warps 2-3): its 96 channels, all images
#pragma unroll
  for (int m = 0; m < LPAD / 16; ++m)
#pragma unroll
    for (int n = 0; n < KLIN / 8; ++n) dEacc[m][n][0] = dEacc[m][n][1] = dEacc[m][n][2] = dEacc[m][n][3] = 0.f;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    lin_stage_X(sX, kv, ldkv, kcol, vcol, NM, L, D, b);
    for (int i = threadIdx.x; i < KLIN * D2 / 4; i += blockDim.x) {     // dpre fp32 -> bf16 [32][PX]
      const int j = (i * 4) / D2, c = (i * 4) % D2;
      const float4 v = *reinterpret_cast<const float4*>(dpre + (long)b * KLIN * D2 + i * 4);
      *reinterpret_cast<uint2*>(sG + j * PX + c) = make_uint2(pack2(v.x, v.y), pack2(v.z, v.w));
    }
    __syncthreads();
    // ---- dKs / dVs: M = l (LPAD), N = this warp's 96 columns, K = j (32)
#pragma unroll
    for (int mt = 0; mt < LPAD / 16; ++mt) {
      float acc[WCOLS / 8][4];
#pragma unroll
      for (int n = 0; n < WCOLS / 8; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < KLIN / 16; ++kk) {
        uint32_t a[4];
        ldA(a, E, PEL, mt * 16, kk * 16, lane);                        // A(m = l, k = j) = E[l][j]
#pragma unroll
        for (int np = 0; np < WCOLS / 16; ++np) {
          uint32_t bb[4];
          ldBt(bb, sG, PX, c0 + np * 16, kk * 16, lane);               // B(k = j, n = c) = dpre[j][c]
          mma16816(acc[2 * np], a, bb[0], bb[1]);
          mma16816(acc[2 * np + 1], a, bb[2], bb[3]);
        }
      }
      const int l0 = mt * 16 + g, l1 = l0 + 8;
      const int col = (which ? dvcol + (c0 - D) : dkcol + c0);
      bf16* d0 = dkv + ((long)b * NM + l0) * lddkv + col;
      bf16* d1 = dkv + ((long)b * NM + l1) * lddkv + col;
#pragma unroll
      for (int n = 0; n < WCOLS / 8; ++n) {
        if (l0 < L) *reinterpret_cast<uint32_t*>(d0 + n * 8 + 2 * t) = pack2(acc[n][0], acc[n][1]);
        if (l1 < L) *reinterpret_cast<uint32_t*>(d1 + n * 8 + 2 * t) = pack2(acc[n][2], acc[n][3]);
      }
    }
    // ---- dE partial over this warp's 96 channels: M = l, N = j (32), K = c (96)
#pragma unroll
    for (int kk = 0; kk < WCOLS / 16; ++kk) {
      uint32_t bb[2][4];
      ldB(bb[0], sG, PX, 0, c0 + kk * 16, lane);                        // B(n = j, k = c) = dpre[j][c]
      ldB(bb[1], sG, PX, 16, c0 + kk * 16, lane);
#pragma unroll
      for (int mt = 0; mt < LPAD / 16; ++mt) {
        uint32_t a[4];
        ldA(a, sX, PX, mt * 16, c0 + kk * 16, lane);                    // A(m = l, k = c) = X[l][c]
        mma16816(dEacc[mt][0], a, bb[0][0], bb[0][1]);
        mma16816(dEacc[mt][1], a, bb[0][2], bb[0][3]);
        mma16816(dEacc[mt][2], a, bb[1][0], bb[1][1]);
        mma16816(dEacc[mt][3], a, bb[1][2], bb[1][3]);
      }
    }
  }
  float* dE = which ? dEv : dEk;
#pragma unroll
  for (int mt = 0; mt < LPAD / 16; ++mt)
#pragma unroll
    for (int n = 0; n < KLIN / 8; ++n) {
      const int l0 = mt * 16 + g, l1 = l0 + 8, j = n * 8 + 2 * t;
      if (l0 < L) { atomicAdd(dE + l0 * KLIN + j, dEacc[mt][n][0]); atomicAdd(dE + l0 * KLIN + j + 1, dEacc[mt][n][1]); }
      if (l1 < L) { atomicAdd(dE + l1 * KLIN + j, dEacc[mt][n][2]); atomicAdd(dE + l1 * KLIN + j + 1, dEacc[mt][n][3]); }
    }
}

// ------------------------------------------------------------------------------------------------ attention
struct WS {   // per-warp bf16 regions
  static constexpr int Q = 0, DO = Q + NQ * PT, KF = DO + NQ * PT, VF = KF + NKV * PT, P = VF + NKV * PT, DS = P + NQ * PT;
  static constexpr int END_FWD = P, END_BWD = DS + NQ * PT;
};

// 16 consecutive global rows x 48 bf16 -> smem [16][PT]
__device__ __forceinline__ void load16(bf16* dst, const bf16* src, long ld, int lane) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int c = lane + 32 * k, i = c / 6, ch = c % 6;
    *reinterpret_cast<uint4*>(dst + i * PT + ch * 8) = *reinterpret_cast<const uint4*>(src + (long)i * ld + ch * 8);
  }
}
// 16 x 48 fp32 (row pitch D) -> bf16 smem rows
__device__ __forceinline__ void load_bank(bf16* dst, const float* src, int D, int col, int lane) {
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int c = lane + 32 * k, i = c / 12, ch = c % 12;
    const float4 v = *reinterpret_cast<const float4*>(src + i * D + col + ch * 4);
    *reinterpret_cast<uint2*>(dst + i * PT + ch * 4) = make_uint2(pack2(v.x, v.y), pack2(v.z, v.w));
  }
}
__device__ __forceinline__ void stage_keys(const AttnP& p, const bf16* pre, bf16* W, int img, int h, int lane) {
  const int D = p.H * HD;
  const bf16* kp = pre + (long)img * KLIN * 2 * D + h * HD;
  load16(W + WS::KF, kp, 2 * D, lane);
  load16(W + WS::KF + 16 * PT, kp + 16L * 2 * D, 2 * D, lane);
  load16(W + WS::VF, kp + D, 2 * D, lane);
  load16(W + WS::VF + 16 * PT, kp + D + 16L * 2 * D, 2 * D, lane);
  load_bank(W + WS::KF + KLIN * PT, p.bank_k, D, h * HD, lane);
  load_bank(W + WS::VF + KLIN * PT, p.bank_v, D, h * HD, lane);
}

__device__ __forceinline__ void scores_softmax(float (*s)[4], const bf16* Qs, const bf16* Kf, float scale, int lane) {
  constexpr int NT = NKV / 8;
#pragma unroll
  for (int n = 0; n < NT; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk) {
    uint32_t a[4];
    ldA(a, Qs, PT, 0, kk * 16, lane);
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      uint32_t b[4];
      ldB(b, Kf, PT, np * 16, kk * 16, lane);
      mma16816(s[2 * np], a, b[0], b[1]);
      mma16816(s[2 * np + 1], a, b[2], b[3]);
    }
  }
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    s[n][0] *= scale; s[n][1] *= scale; s[n][2] *= scale; s[n][3] *= scale;
    m0 = fmaxf(m0, fmaxf(s[n][0], s[n][1]));
    m1 = fmaxf(m1, fmaxf(s[n][2], s[n][3]));
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  float z0 = 0.f, z1 = 0.f;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    s[n][0] = __expf(s[n][0] - m0); s[n][1] = __expf(s[n][1] - m0);
    s[n][2] = __expf(s[n][2] - m1); s[n][3] = __expf(s[n][3] - m1);
    z0 += s[n][0] + s[n][1];
    z1 += s[n][2] + s[n][3];
  }
  z0 += __shfl_xor_sync(0xffffffffu, z0, 1); z0 += __shfl_xor_sync(0xffffffffu, z0, 2);
  z1 += __shfl_xor_sync(0xffffffffu, z1, 1); z1 += __shfl_xor_sync(0xffffffffu, z1, 2);
  z0 = 1.f / z0; z1 = 1.f / z1;
#pragma unroll
  for (int n = 0; n < NT; ++n) { s[n][0] *= z0; s[n][1] *= z0; s[n][2] *= z1; s[n][3] *= z1; }
}
// out[16 x 48] = X[16 x 48 keys] (C-layout registers as A) * Bs (smem [key][dim])
__device__ __forceinline__ void regA_times_Bt(float (*o)[4], const float (*x)[4], const bf16* Bs, int lane) {
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < NKV / 16; ++kk) {
    uint32_t a[4] = {pack2(x[2 * kk][0], x[2 * kk][1]), pack2(x[2 * kk][2], x[2 * kk][3]),
                     pack2(x[2 * kk + 1][0], x[2 * kk + 1][1]), pack2(x[2 * kk + 1][2], x[2 * kk + 1][3])};
#pragma unroll
    for (int np = 0; np < HD / 16; ++np) {
      uint32_t b[4];
      ldBt(b, Bs, PT, np * 16, kk * 16, lane);
      mma16816(o[2 * np], a, b[0], b[1]);
      mma16816(o[2 * np + 1], a, b[2], b[3]);
    }
  }
}
__device__ __forceinline__ void store16(bf16* dst, long ld, const float (*o)[4], int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) {
    *reinterpret_cast<uint32_t*>(dst + (long)g * ld + n * 8 + 2 * t) = pack2(o[n][0], o[n][1]);
    *reinterpret_cast<uint32_t*>(dst + (long)(g + 8) * ld + n * 8 + 2 * t) = pack2(o[n][2], o[n][3]);
  }
}

__global__ void __launch_bounds__(WARPS * 32) msda64_fwd_kernel(AttnP p, const bf16* __restrict__ pre, int ntask) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  bf16* W = reinterpret_cast<bf16*>(smraw) + warp * WS::END_FWD;
  const float scale = rsqrtf((float)HD);
  const bf16* q = static_cast<const bf16*>(p.q);
  bf16* out = static_cast<bf16*>(p.out);
  const int tiles = p.Nt / NQ;
  for (int task = blockIdx.x * WARPS + warp; task < ntask; task += gridDim.x * WARPS) {
    const int img = task / p.H, h = task % p.H;
    __syncwarp();
    stage_keys(p, pre, W, img, h, lane);
    for (int qt = 0; qt < tiles; ++qt) {
      const long row0 = (long)img * p.Nt + qt * NQ;
      __syncwarp();
      load16(W + WS::Q, q + row0 * p.ldq + p.qcol + h * HD, p.ldq, lane);
      __syncwarp();
      float s[NKV / 8][4], o[HD / 8][4];
      scores_softmax(s, W + WS::Q, W + WS::KF, scale, lane);
      if (p.drop.p > 0.f) {   // SDPA dropout_p on the probabilities
        const DropState ds = drop_state(p.drop);
        drop_apply_c<NKV / 8>(s, drop_bits_c<NKV / 8>(ds, (uint32_t)(task * tiles + qt), lane), ds.inv);
      }
      regA_times_Bt(o, s, W + WS::VF, lane);
      store16(out + row0 * p.ldo + h * HD, p.ldo, o, lane);
    }
  }
}

// c[3][6][4] = A^T (smem [k = 16 queries][m = 48 keys]) * B (smem [k = 16 queries][n = 48])
__device__ __forceinline__ void keysT_times(float (*c)[6][4], const bf16* As, const bf16* Bs, int lane) {
#pragma unroll
  for (int mt = 0; mt < NKV / 16; ++mt) {
    uint32_t a[4];
    ldAt(a, As, PT, mt * 16, 0, lane);
#pragma unroll
    for (int np = 0; np < 3; ++np) {
      uint32_t b[4];
      ldBt(b, Bs, PT, np * 16, 0, lane);
#pragma unroll
      for (int e = 0; e < 4; ++e) c[mt][2 * np][e] = c[mt][2 * np + 1][e] = 0.f;
      mma16816(c[mt][2 * np], a, b[0], b[1]);
      mma16816(c[mt][2 * np + 1], a, b[2], b[3]);
    }
  }
}
// Linformer rows of c (keys 0..31) += into the warp-private fp32 accumulator [32][48]; bank rows into registers
__device__ __forceinline__ void accumulate(float* acc, float (*dbank)[4], const float (*c)[6][4], int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int n = 0; n < 6; ++n) {
      float2* a0 = reinterpret_cast<float2*>(acc + (mt * 16 + g) * HD + n * 8 + 2 * t);
      float2* a1 = reinterpret_cast<float2*>(acc + (mt * 16 + g + 8) * HD + n * 8 + 2 * t);
      float2 u = *a0, v = *a1;
      u.x += c[mt][n][0]; u.y += c[mt][n][1]; v.x += c[mt][n][2]; v.y += c[mt][n][3];
      *a0 = u; *a1 = v;
    }
#pragma unroll
  for (int n = 0; n < 6; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) dbank[n][e] += c[2][n][e];
}

__global__ void __launch_bounds__(WARPS * 32) msda64_bwd_kernel(AttnP p, const bf16* __restrict__ pre, float* __restrict__ dpre, int ntask) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  const int D = p.H * HD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  float* accK = reinterpret_cast<float*>(smraw) + warp * 2 * KLIN * HD;     // [32][48] fp32, then accV
  float* accV = accK + KLIN * HD;
  bf16* W = reinterpret_cast<bf16*>(reinterpret_cast<float*>(smraw) + WARPS * 2 * KLIN * HD) + warp * WS::END_BWD;
  const float scale = rsqrtf((float)HD);
  const bf16* q = static_cast<const bf16*>(p.q);
  const bf16* dout = static_cast<const bf16*>(p.dout);
  bf16* dq = static_cast<bf16*>(p.dq);
  const int tiles = p.Nt / NQ;
  constexpr int NT = NKV / 8;
  // H == WARPS and the task stride is a multiple of WARPS: this warp always serves head `warp`
  float dbk[6][4], dbv[6][4];
#pragma unroll
  for (int n = 0; n < 6; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) dbk[n][e] = dbv[n][e] = 0.f;
  const int h = warp;
  for (int task = blockIdx.x * WARPS + warp; task < ntask; task += gridDim.x * WARPS) {
    const int img = task / p.H;
    __syncwarp();
    stage_keys(p, pre, W, img, h, lane);
    for (int i = lane; i < 2 * KLIN * HD; i += 32) accK[i] = 0.f;
    for (int qt = 0; qt < tiles; ++qt) {
      const long row0 = (long)img * p.Nt + qt * NQ;
      __syncwarp();
      load16(W + WS::Q, q + row0 * p.ldq + p.qcol + h * HD, p.ldq, lane);
      load16(W + WS::DO, dout + row0 * p.lddo + h * HD, p.lddo, lane);
      __syncwarp();
      float P[NT][4], dS[NT][4];
      scores_softmax(P, W + WS::Q, W + WS::KF, scale, lane);
#pragma unroll
      for (int n = 0; n < NT; ++n) dS[n][0] = dS[n][1] = dS[n][2] = dS[n][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {                         // dP = dO Vf^T
        uint32_t a[4];
        ldA(a, W + WS::DO, PT, 0, kk * 16, lane);
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) {
          uint32_t b[4];
          ldB(b, W + WS::VF, PT, np * 16, kk * 16, lane);
          mma16816(dS[2 * np], a, b[0], b[1]);
          mma16816(dS[2 * np + 1], a, b[2], b[3]);
        }
      }
      unsigned long long keep = ~0ull;
      float kinv = 1.f;
      if (p.drop.p > 0.f) {   // dP <- d(P_dropped) * keep; P_dropped (for dVf) is formed after the softmax backward
        const DropState ds = drop_state(p.drop);
        keep = drop_bits_c<NT>(ds, (uint32_t)(task * tiles + qt), lane);
        kinv = ds.inv;
        drop_apply_c<NT>(dS, keep, kinv);
      }
      float r0 = 0.f, r1 = 0.f;
#pragma unroll
      for (int n = 0; n < NT; ++n) { r0 += dS[n][0] * P[n][0] + dS[n][1] * P[n][1]; r1 += dS[n][2] * P[n][2] + dS[n][3] * P[n][3]; }
      r0 += __shfl_xor_sync(0xffffffffu, r0, 1); r0 += __shfl_xor_sync(0xffffffffu, r0, 2);
      r1 += __shfl_xor_sync(0xffffffffu, r1, 1); r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        dS[n][0] = P[n][0] * (dS[n][0] - r0) * scale; dS[n][1] = P[n][1] * (dS[n][1] - r0) * scale;
        dS[n][2] = P[n][2] * (dS[n][2] - r1) * scale; dS[n][3] = P[n][3] * (dS[n][3] - r1) * scale;
      }
      if (p.drop.p > 0.f) drop_apply_c<NT>(P, keep, kinv);
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        stC(W + WS::P, PT, 0, n * 8, P[n], lane);
        stC(W + WS::DS, PT, 0, n * 8, dS[n], lane);
      }
      {
        float o[HD / 8][4];
        regA_times_Bt(o, dS, W + WS::KF, lane);                      // dQ = dS Kf
        store16(dq + row0 * p.lddq + p.dqcol + h * HD, p.lddq, o, lane);
      }
      __syncwarp();
      {
        float c[NKV / 16][6][4];
        keysT_times(c, W + WS::P, W + WS::DO, lane);                 // dVf += P^T dO
        accumulate(accV, dbv, c, lane);
        keysT_times(c, W + WS::DS, W + WS::Q, lane);                 // dKf += dS^T Q
        accumulate(accK, dbk, c, lane);
      }
    }
    __syncwarp();
    // dK' / dV' of this (image, head): [32][48] fp32 -> dpre[img, j, (K | V) + h * 48 ..]
    float* dst = dpre + (long)img * KLIN * 2 * D + h * HD;
    for (int i = lane; i < KLIN * (HD / 4); i += 32) {
      const int j = i / (HD / 4), c4 = (i % (HD / 4)) * 4;
      *reinterpret_cast<float4*>(dst + (long)j * 2 * D + c4) = *reinterpret_cast<const float4*>(accK + j * HD + c4);
      *reinterpret_cast<float4*>(dst + (long)j * 2 * D + D + c4) = *reinterpret_cast<const float4*>(accV + j * HD + c4);
    }
  }
#pragma unroll
  for (int n = 0; n < 6; ++n) {
    float* k0 = p.dbank_k + g * D + h * HD + n * 8 + 2 * t;
    float* v0 = p.dbank_v + g * D + h * HD + n * 8 + 2 * t;
    atomicAdd(k0, dbk[n][0]); atomicAdd(k0 + 1, dbk[n][1]); atomicAdd(k0 + 8 * D, dbk[n][2]); atomicAdd(k0 + 8 * D + 1, dbk[n][3]);
    atomicAdd(v0, dbv[n][0]); atomicAdd(v0 + 1, dbv[n][1]); atomicAdd(v0 + 8 * D, dbv[n][2]); atomicAdd(v0 + 8 * D + 1, dbv[n][3]);
  }
}

size_t lin_smem(int L, int D, bool post) { return ((size_t)2 * L * KLIN + (size_t)L * 2 * D + (post ? (size_t)KLIN * (2 * D + 4) : 0)) * sizeof(float); }

}  // namespace

bool attn_msda64_ok(const AttnP& p) {
  return p.mode == 1 && p.hd == HD && p.H == WARPS && p.kb == KB && p.klin == KLIN && p.Nt > NQ && p.Nt % NQ == 0 &&
         p.L >= 1 && p.L <= 128 && p.NM <= 128 && p.L == min(p.NM, 128) && p.ldq % 8 == 0 && p.qcol % 8 == 0 && p.ldo % 8 == 0 &&
         p.ldkv % 2 == 0 && p.kcol % 2 == 0 && p.vcol % 2 == 0 && (p.H * HD) % 4 == 0;
}
size_t attn_msda64_scratch_bytes(int B, int D) { return (size_t)B * KLIN * 2 * D * (2 + 4) + 512; }

static bool lin_mma_ok(const AttnP& p) {
  return p.L <= LPAD && p.ldkv % 8 == 0 && p.kcol % 8 == 0 && p.vcol % 8 == 0 && p.lddkv % 2 == 0 && p.dkcol % 2 == 0 && p.dvcol % 2 == 0;
}

static int run_pre(cudaStream_t s, const AttnP& p, bf16* pre) {
  const int D = p.H * HD;
  if (lin_mma_ok(p)) {
    const size_t smem = (size_t)LinSmem::END_PRE * sizeof(bf16);
    QV_CUDA(cudaFuncSetAttribute(lin_pre_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    qv_launch(lin_pre_mma_kernel, max(1, min(p.B, qv_num_sms() * 4)), WARPS * 32, smem, s, (const bf16*)p.kv, p.ldkv, p.kcol, p.vcol, p.NM, p.L,
                                                                                 p.Ek, p.Ev, p.B, pre);
    QV_LAUNCH_CHECK();
    return 0;
  }
  const size_t smem = lin_smem(p.L, D, false);
  QV_CHECK(smem <= 200 * 1024, "msda64: Linformer staging needs %zu B of shared memory", smem);
  QV_CUDA(cudaFuncSetAttribute(lin_pre_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int occ = max(1, min(4, (int)(200 * 1024 / (smem + 1024))));
  qv_launch(lin_pre_kernel, max(1, min(p.B, qv_num_sms() * occ)), 256, smem, s, (const bf16*)p.kv, p.ldkv, p.kcol, p.vcol, p.NM, p.L, D, p.Ek, p.Ev,
                                                                        p.B, pre);
  QV_LAUNCH_CHECK();
  return 0;
}

int attn_msda64_fwd(cudaStream_t s, const AttnP& p, void* scratch) {
  if (p.B <= 0) return 0;
  QV_CHECK(scratch, "msda64: scratch missing");
  bf16* pre = static_cast<bf16*>(scratch);
  QV_TRY(run_pre(s, p, pre));
  const int ntask = p.B * p.H;
  const size_t smem = (size_t)WARPS * WS::END_FWD * sizeof(bf16);
  QV_CUDA(cudaFuncSetAttribute(msda64_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int occ = max(1, min(4, (int)(200 * 1024 / (smem + 1024))));
  qv_launch(msda64_fwd_kernel, min(cdiv(ntask, WARPS), qv_num_sms() * occ), WARPS * 32, smem, s, p, pre, ntask);
  QV_LAUNCH_CHECK();
  return 0;
}

int attn_msda64_bwd(cudaStream_t s, const AttnP& p, void* scratch) {
  if (p.B <= 0) return 0;
  QV_CHECK(scratch, "msda64: scratch missing");
  const int D = p.H * HD;
  bf16* pre = static_cast<bf16*>(scratch);
  float* dpre = reinterpret_cast<float*>(static_cast<uint8_t*>(scratch) + (((size_t)p.B * KLIN * 2 * D * 2 + 255) & ~(size_t)255));
  QV_TRY(run_pre(s, p, pre));                                       // recomputed: cheaper than keeping it across the block
  const int ntask = p.B * p.H;
  const size_t smem = (size_t)WARPS * 2 * KLIN * HD * sizeof(float) + (size_t)WARPS * WS::END_BWD * sizeof(bf16);
  QV_CUDA(cudaFuncSetAttribute(msda64_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int occ = max(1, min(2, (int)(200 * 1024 / (smem + 1024))));
  qv_launch(msda64_bwd_kernel, min(cdiv(ntask, WARPS), qv_num_sms() * occ), WARPS * 32, smem, s, p, pre, dpre, ntask);
  QV_LAUNCH_CHECK();
  if (lin_mma_ok(p)) {
    const size_t smem2 = (size_t)LinSmem::END_POST * sizeof(bf16);
    QV_CUDA(cudaFuncSetAttribute(lin_post_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    qv_launch(lin_post_mma_kernel, max(1, min(p.B, qv_num_sms() * 3)), WARPS * 32, smem2, s, 
        (const bf16*)p.kv, p.ldkv, p.kcol, p.vcol, p.NM, p.L, p.Ek, p.Ev, p.B, dpre, (bf16*)p.dkv, p.lddkv, p.dkcol, p.dvcol, p.dEk, p.dEv);
    QV_LAUNCH_CHECK();
    return 0;
  }
  const size_t smem2 = lin_smem(p.L, D, true);
  QV_CHECK(smem2 <= 200 * 1024, "msda64: Linformer backward staging needs %zu B of shared memory", smem2);
  QV_CUDA(cudaFuncSetAttribute(lin_post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  const int occ2 = max(1, min(2, (int)(200 * 1024 / (smem2 + 1024))));
  qv_launch(lin_post_kernel, max(1, min(p.B, qv_num_sms() * occ2)), 256, smem2, s, (const bf16*)p.kv, p.ldkv, p.kcol, p.vcol, p.NM, p.L, D, p.Ek, p.Ev,
                                                                           p.B, dpre, (bf16*)p.dkv, p.lddkv, p.dkcol, p.dvcol, p.dEk, p.dEv);
  QV_LAUNCH_CHECK();
  return 0;
}
