// Tensor-core attention for the 16-query problems of the quad block (bf16 runs): one WARP per (window, head) task,
// mma.sync.m16n8k16 (bf16 x bf16 -> fp32) fed by ldmatrix from warp-private shared memory.
//
// Why mma.sync and not tcgen05 here: a task is Q[16 x 48] against 48 (or 16) keys -- one m16 tile.  A 128-row UMMA
// tile would have to stack 8 unrelated tasks block-diagonally (8x wasted MMA) and round-trip S/P through TMEM;
// m16n8k16 matches the problem exactly and keeps S, P, dS in registers (SURVEY.md 7.2 hard part 3).  The SIMT kernels
// in attn.cu stay as the fp32 parity path and for shapes outside this file's (nq = 16, L <= 16, k = 32, bank = 16).
//
//   forward : K' = E_k^T Ks, V' = E_v^T Vs (Linformer) -> Kf = [K'; bank_k] -> S = Q Kf^T -> softmax -> O = P Vf
//   backward: recompute P; dP = dO Vf^T; dS = P (dP - rowsum(P dP)) / sqrt(hd); dQ = dS Kf; dVf = P^T dO;
//             dKf = dS^T Q; bank rows of dKf/dVf -> d bank; Linformer rows -> dKs = E_k dK', dE_k += Ks dK'^T
// Batch reductions (dE, d bank) accumulate in CTA shared memory (fp32 atomics) and are flushed once per CTA.
#include "kernels.h"

namespace {

constexpr int HD = 48, NQ = 16, LP = 16, KLIN = 32, KB = 16;
constexpr int PT = 56;    // pitch (bf16 elements) of 48-wide rows: 112 B -> conflict-free 16 B row accesses
constexpr int PE = 40;    // pitch of the 32-wide Linformer matrices
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

// A fragment (16 x 16) from smem stored [m][k] (k contiguous)
__device__ __forceinline__ void ldA(uint32_t* a, const bf16* base, int pitch, int m0, int k0, int lane) {
  const int mat = lane >> 3, r = lane & 7;
  ldsm4(a, sa(base + (m0 + r + (mat & 1) * 8) * pitch + k0 + (mat >> 1) * 8));
}
// A fragment from smem stored [k][m] (m contiguous): A(m, k) = S[k][m]
__device__ __forceinline__ void ldAt(uint32_t* a, const bf16* base, int pitch, int m0, int k0, int lane) {
  const int mat = lane >> 3, r = lane & 7;
  ldsm4t(a, sa(base + (k0 + r + (mat >> 1) * 8) * pitch + m0 + (mat & 1) * 8));
}
// B fragments of TWO adjacent n8 tiles (b[0..1] = tile n0, b[2..3] = tile n0 + 8) from smem stored [n][k]
__device__ __forceinline__ void ldB(uint32_t* b, const bf16* base, int pitch, int n0, int k0, int lane) {
  const int mat = lane >> 3, r = lane & 7;
  ldsm4(b, sa(base + (n0 + r + (mat >> 1) * 8) * pitch + k0 + (mat & 1) * 8));
}
// same from smem stored [k][n] (n contiguous)
__device__ __forceinline__ void ldBt(uint32_t* b, const bf16* base, int pitch, int n0, int k0, int lane) {
  const int mat = lane >> 3, r = lane & 7;
  ldsm4t(b, sa(base + (k0 + r + (mat & 1) * 8) * pitch + n0 + (mat >> 1) * 8));
}
// C fragment (rows g / g+8, cols 2t / 2t+1 of an m16n8 tile) -> bf16 smem [m][n]
__device__ __forceinline__ void stC(bf16* base, int pitch, int m0, int n0, const float* c, int lane) {
  const int g = lane >> 2, t = lane & 3;
  *reinterpret_cast<uint32_t*>(base + (m0 + g) * pitch + n0 + 2 * t) = pack2(c[0], c[1]);
  *reinterpret_cast<uint32_t*>(base + (m0 + g + 8) * pitch + n0 + 2 * t) = pack2(c[2], c[3]);
}

struct WarpSmem {   // per-warp regions (bf16 element offsets from the warp base)
  static constexpr int Q = 0, DO = Q + NQ * PT, KS = DO + NQ * PT, VS = KS + LP * PT, KF = VS + LP * PT;
  static constexpr int VF = KF + (KLIN + KB) * PT, P = VF + (KLIN + KB) * PT, DS = P + NQ * PT, END = DS + NQ * PT;
};

__device__ __forceinline__ int q_row(const AttnP& p, int w, int i) {
  if (p.mode == 0) {
    const int nws = p.side / p.ws, nW = nws * nws;
    const int b = w / nW, wi = w % nW;
    return b * p.Nt + ((wi / nws) * p.ws + i / p.ws) * p.side + (wi % nws) * p.ws + i % p.ws;
  }
  return w * p.Nt + i;
}

// 16 rows x 48 bf16 from global (row index from `rowsrc` lane registers) -> smem [16][PT]; rows >= nvalid zeroed
__device__ __forceinline__ void load_rows(bf16* dst, const bf16* src, long ld, int col, int my_row, int nvalid, int lane) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int c = lane + 32 * k, i = c / 6, ch = c % 6;
    const int row = __shfl_sync(0xffffffffu, my_row, i);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (i < nvalid) v = *reinterpret_cast<const uint4*>(src + (long)row * ld + col + ch * 8);
    *reinterpret_cast<uint4*>(dst + i * PT + ch * 8) = v;
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

// Kf/Vf rows [0, 32) = E^T Xs  (M = 32 (j), N = 48 (d), K = 16 (l))
__device__ __forceinline__ void linformer_fwd(bf16* Xf, const bf16* E, const bf16* Xs, int lane) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    uint32_t a[4];
    ldAt(a, E, PE, mt * 16, 0, lane);
#pragma unroll
    for (int np = 0; np < 3; ++np) {
      uint32_t b[4];
      ldBt(b, Xs, PT, np * 16, 0, lane);
      float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
      mma16816(c0, a, b[0], b[1]);
      mma16816(c1, a, b[2], b[3]);
      stC(Xf, PT, mt * 16, np * 16, c0, lane);
      stC(Xf, PT, mt * 16, np * 16 + 8, c1, lane);
    }
  }
}

// S = Q Kf^T -> softmax probabilities in C-fragment layout s[NT][4] (NT = NKV / 8 key tiles)
template <int NKV>
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

// out[16 x 48] = X[16 x NKV] (C-layout registers, used as A) * Bs (smem [k][n], n contiguous)
template <int NKV>
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

// C-layout [16 x 48] registers -> global bf16 rows
__device__ __forceinline__ void store_rows(bf16* dst, long ld, int col, int my_row, const float (*o)[4], int lane) {
  const int g = lane >> 2, t = lane & 3;
  const int r0 = __shfl_sync(0xffffffffu, my_row, g), r1 = __shfl_sync(0xffffffffu, my_row, g + 8);
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) {
    *reinterpret_cast<uint32_t*>(dst + (long)r0 * ld + col + n * 8 + 2 * t) = pack2(o[n][0], o[n][1]);
    *reinterpret_cast<uint32_t*>(dst + (long)r1 * ld + col + n * 8 + 2 * t) = pack2(o[n][2], o[n][3]);
  }
}

template <int NKV, bool LINF>
__device__ __forceinline__ void stage_task(const AttnP& p, bf16* W, const bf16* Es, int w, int h, int lane, int& my_q,
                                           int& my_kv) {
  const int D = p.H * HD;
  my_q = q_row(p, w, lane & 15);
  const bf16* q = static_cast<const bf16*>(p.q);
  load_rows(W + WarpSmem::Q, q, p.ldq, p.qcol + h * HD, my_q, NQ, lane);
  if (LINF) {
    my_kv = (p.mode == 0) ? my_q : w * p.NM + min(lane & 15, p.NM - 1);
    const bf16* kv = static_cast<const bf16*>(p.kv);
    load_rows(W + WarpSmem::KS, kv, p.ldkv, p.kcol + h * HD, my_kv, p.L, lane);
    load_rows(W + WarpSmem::VS, kv, p.ldkv, p.vcol + h * HD, my_kv, p.L, lane);
    load_bank(W + WarpSmem::KF + KLIN * PT, p.bank_k, D, h * HD, lane);
    load_bank(W + WarpSmem::VF + KLIN * PT, p.bank_v, D, h * HD, lane);
    __syncwarp();
    linformer_fwd(W + WarpSmem::KF, Es, W + WarpSmem::KS, lane);
    linformer_fwd(W + WarpSmem::VF, Es + LP * PE, W + WarpSmem::VS, lane);
  } else {
    load_bank(W + WarpSmem::KF, p.kc, D, h * HD, lane);
    load_bank(W + WarpSmem::VF, p.vc, D, h * HD, lane);
  }
  __syncwarp();
}

// E_k / E_v fp32 [L][32] -> bf16 smem [LP][PE] (rows >= L zero): once per CTA
__device__ __forceinline__ void load_E(bf16* Es, const AttnP& p) {
  for (int idx = threadIdx.x; idx < 2 * LP * KLIN; idx += blockDim.x) {
    const int which = idx / (LP * KLIN), r = idx % (LP * KLIN), l = r / KLIN, j = r % KLIN;
    const float* src = which ? p.Ev : p.Ek;
    Es[which * LP * PE + l * PE + j] = __float2bfloat16_rn(l < p.L ? src[l * KLIN + j] : 0.f);
  }
}

template <int NKV, bool LINF>
__global__ void __launch_bounds__(WARPS * 32) attn_mma_fwd_kernel(AttnP p, int ntask) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  bf16* Es = reinterpret_cast<bf16*>(smraw);                       // [2][LP][PE]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  bf16* W = Es + 2 * LP * PE + warp * WarpSmem::END;
  if (LINF) load_E(Es, p);
  __syncthreads();
  const float scale = rsqrtf((float)HD);
  bf16* out = static_cast<bf16*>(p.out);
  for (int task = blockIdx.x * WARPS + warp; task < ntask; task += gridDim.x * WARPS) {
    const int w = task / p.H, h = task % p.H;
    int my_q, my_kv;
    stage_task<NKV, LINF>(p, W, Es, w, h, lane, my_q, my_kv);
    float s[NKV / 8][4], o[HD / 8][4];
    scores_softmax<NKV>(s, W + WarpSmem::Q, W + WarpSmem::KF, scale, lane);
    if (p.drop.p > 0.f) {   // SDPA dropout_p on the probabilities
      const DropState ds = drop_state(p.drop);
      drop_apply_c<NKV / 8>(s, drop_bits_c<NKV / 8>(ds, (uint32_t)task, lane), ds.inv);
    }
    regA_times_Bt<NKV>(o, s, W + WarpSmem::VF, lane);
    store_rows(out, p.ldo, h * HD, my_q, o, lane);
    __syncwarp();
  }
}

// C[MT*16 x 48] = A^T(smem [k = 16][m]) * B(smem [k = 16][n = 48]) ; c[MT][6][4]
template <int MT>
__device__ __forceinline__ void At_times_Bt_k16(float (*c)[6][4], const bf16* As, const bf16* Bs, int lane) {
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
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

// Consume dXf (X = K or V): bank rows -> the warp's register accumulator (a warp always serves the same head);
// Linformer rows -> dXs = E dX' (global) and dE += Xs dX'^T (accumulated in the mma C registers across tasks).
template <int NKV, bool LINF>
__device__ __forceinline__ void consume_dXf(const AttnP& p, float (*c)[6][4], bf16* stage /*[32][PT]*/, const bf16* E,
                                            const bf16* Xs, float (*dE_acc)[4], float (*dbank_acc)[4], int h, int my_kv,
                                            int dcol, int lane) {
  const int g = lane >> 2, t = lane & 3;
  constexpr int MT = NKV / 16;
#pragma unroll
  for (int n = 0; n < 6; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) dbank_acc[n][e] += c[MT - 1][n][e];
  if (!LINF) return;
  // dX' (32 x 48) -> bf16 staging [j][d]
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int n = 0; n < 6; ++n) stC(stage, PT, mt * 16, n * 8, c[mt][n], lane);
  __syncwarp();
  // dXs[l, d] = sum_j E[l, j] dX'[j, d] : A = E [m = l][k = j], B = dX' [k = j][n = d]
  {
    float o[6][4];
#pragma unroll
    for (int n = 0; n < 6; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      uint32_t a[4];
      ldA(a, E, PE, 0, kk * 16, lane);
#pragma unroll
      for (int np = 0; np < 3; ++np) {
        uint32_t b[4];
        ldBt(b, stage, PT, np * 16, kk * 16, lane);
        mma16816(o[2 * np], a, b[0], b[1]);
        mma16816(o[2 * np + 1], a, b[2], b[3]);
      }
    }
    bf16* dkv = static_cast<bf16*>(p.dkv);
    const int r0 = __shfl_sync(0xffffffffu, my_kv, g), r1 = __shfl_sync(0xffffffffu, my_kv, g + 8);
#pragma unroll
    for (int n = 0; n < 6; ++n) {
      if (g < p.L) *reinterpret_cast<uint32_t*>(dkv + (long)r0 * p.lddkv + dcol + h * HD + n * 8 + 2 * t) = pack2(o[n][0], o[n][1]);
      if (g + 8 < p.L) *reinterpret_cast<uint32_t*>(dkv + (long)r1 * p.lddkv + dcol + h * HD + n * 8 + 2 * t) = pack2(o[n][2], o[n][3]);
    }
  }
  // dE[l, j] += sum_d Xs[l, d] dX'[j, d] : A = Xs [m = l][k = d], B = dX' [n = j][k = d]
#pragma unroll
  for (int kk = 0; kk < 3; ++kk) {
    uint32_t a[4];
    ldA(a, Xs, PT, 0, kk * 16, lane);
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t b[4];
      ldB(b, stage, PT, np * 16, kk * 16, lane);
      mma16816(dE_acc[2 * np], a, b[0], b[1]);
      mma16816(dE_acc[2 * np + 1], a, b[2], b[3]);
    }
  }
  __syncwarp();
}

template <int NKV, bool LINF>
__global__ void __launch_bounds__(WARPS * 32) attn_mma_bwd_kernel(AttnP p, int ntask) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  const int D = p.H * HD;
  bf16* Es = reinterpret_cast<bf16*>(smraw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  bf16* W = Es + 2 * LP * PE + warp * WarpSmem::END;
  if (LINF) load_E(Es, p);
  __syncthreads();
  const float scale = rsqrtf((float)HD);
  const bf16* dout = static_cast<const bf16*>(p.dout);
  bf16* dq = static_cast<bf16*>(p.dq);
  constexpr int NT = NKV / 8;
  // batch reductions live in registers: H == WARPS and the task stride is a multiple of WARPS, so this warp's head is
  // fixed (h == warp) and its bank-row gradients are one [16 x 48] C tile per K and V; dE_k / dE_v are [16 x 32] tiles.
  float dbk[6][4], dbv[6][4], dEk[4][4], dEv[4][4];
#pragma unroll
  for (int n = 0; n < 6; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) dbk[n][e] = dbv[n][e] = 0.f;
#pragma unroll
  for (int n = 0; n < 4; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) dEk[n][e] = dEv[n][e] = 0.f;
  const int h = warp;
  for (int task = blockIdx.x * WARPS + warp; task < ntask; task += gridDim.x * WARPS) {
    const int w = task / p.H;
    int my_q, my_kv = 0;
    stage_task<NKV, LINF>(p, W, Es, w, h, lane, my_q, my_kv);
    load_rows(W + WarpSmem::DO, dout, p.lddo, h * HD, my_q, NQ, lane);
    __syncwarp();
    float P[NT][4], dS[NT][4];
    scores_softmax<NKV>(P, W + WarpSmem::Q, W + WarpSmem::KF, scale, lane);
    // dP = dO Vf^T
#pragma unroll
    for (int n = 0; n < NT; ++n) dS[n][0] = dS[n][1] = dS[n][2] = dS[n][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
      uint32_t a[4];
      ldA(a, W + WarpSmem::DO, PT, 0, kk * 16, lane);
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t b[4];
        ldB(b, W + WarpSmem::VF, PT, np * 16, kk * 16, lane);
        mma16816(dS[2 * np], a, b[0], b[1]);
        mma16816(dS[2 * np + 1], a, b[2], b[3]);
      }
    }
    unsigned long long keep = ~0ull;
    float kinv = 1.f;
    if (p.drop.p > 0.f) {   // dP <- d(P_dropped) * keep; P_dropped (for dVf) is formed after the softmax backward
      const DropState ds = drop_state(p.drop);
      keep = drop_bits_c<NT>(ds, (uint32_t)task, lane);
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
      stC(W + WarpSmem::P, PT, 0, n * 8, P[n], lane);
      stC(W + WarpSmem::DS, PT, 0, n * 8, dS[n], lane);
    }
    // dQ = dS Kf
    {
      float o[HD / 8][4];
      regA_times_Bt<NKV>(o, dS, W + WarpSmem::KF, lane);
      store_rows(dq, p.lddq, p.dqcol + h * HD, my_q, o, lane);
    }
    __syncwarp();
    // dVf = P^T dO ; dKf = dS^T Q  (M = NKV keys, N = 48, K = 16 queries); Kf / Vf rows [0, 32) are free from here on
    {
      float c[NKV / 16][6][4];
      At_times_Bt_k16<NKV / 16>(c, W + WarpSmem::P, W + WarpSmem::DO, lane);
      consume_dXf<NKV, LINF>(p, c, W + WarpSmem::VF, Es + LP * PE, W + WarpSmem::VS, dEv, dbv, h, my_kv, p.dvcol, lane);
      At_times_Bt_k16<NKV / 16>(c, W + WarpSmem::DS, W + WarpSmem::Q, lane);
      consume_dXf<NKV, LINF>(p, c, W + WarpSmem::KF, Es, W + WarpSmem::KS, dEk, dbk, h, my_kv, p.dkcol, lane);
    }
    __syncwarp();
  }
  // ---- flush the per-warp register accumulators (one atomic per element per warp)
#pragma unroll
  for (int n = 0; n < 6; ++n) {
    float* k0 = p.dbank_k + g * D + h * HD + n * 8 + 2 * t;
    float* v0 = p.dbank_v + g * D + h * HD + n * 8 + 2 * t;
    atomicAdd(k0, dbk[n][0]); atomicAdd(k0 + 1, dbk[n][1]); atomicAdd(k0 + 8 * D, dbk[n][2]); atomicAdd(k0 + 8 * D + 1, dbk[n][3]);
    atomicAdd(v0, dbv[n][0]); atomicAdd(v0 + 1, dbv[n][1]); atomicAdd(v0 + 8 * D, dbv[n][2]); atomicAdd(v0 + 8 * D + 1, dbv[n][3]);
  }
  if (LINF) {
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int j = n * 8 + 2 * t;
      if (g < p.L) { atomicAdd(p.dEk + g * KLIN + j, dEk[n][0]); atomicAdd(p.dEk + g * KLIN + j + 1, dEk[n][1]);
                     atomicAdd(p.dEv + g * KLIN + j, dEv[n][0]); atomicAdd(p.dEv + g * KLIN + j + 1, dEv[n][1]); }
      if (g + 8 < p.L) { atomicAdd(p.dEk + (g + 8) * KLIN + j, dEk[n][2]); atomicAdd(p.dEk + (g + 8) * KLIN + j + 1, dEk[n][3]);
                         atomicAdd(p.dEv + (g + 8) * KLIN + j, dEv[n][2]); atomicAdd(p.dEv + (g + 8) * KLIN + j + 1, dEv[n][3]); }
    }
  }
}

size_t smem_bytes(const AttnP& p, bool bwd) {
  (void)p; (void)bwd;
  return (size_t)(2 * LP * PE + WARPS * WarpSmem::END) * sizeof(bf16);
}

template <typename K>
int launch(K kernel, cudaStream_t s, const AttnP& p, bool bwd) {
  const int nwin = (p.mode == 0) ? p.B * (p.side / p.ws) * (p.side / p.ws) : p.B;
  const int ntask = nwin * p.H;
  if (ntask <= 0) return 0;
  const size_t smem = smem_bytes(p, bwd);
  if (smem > 48 * 1024) QV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int occ = max(1, min(4, (int)(220 * 1024 / (smem + 1024))));
  const int grid = min(cdiv(ntask, WARPS), qv_num_sms() * occ);
  qv_launch(kernel, grid, WARPS * 32, smem, s, p, ntask);
  QV_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// shapes this file is instantiated for
bool attn_mma_ok(const AttnP& p) {
  const int nq = (p.mode == 0) ? p.ws * p.ws : p.Nt;
  if (p.hd != HD || nq != NQ || p.kb != KB || p.H != WARPS) return false;   // H == WARPS: a warp serves one head
  if (p.ldq % 8 || p.qcol % 8 || p.ldo % 8) return false;
  if (p.mode != 2) {
    if (p.klin != KLIN || p.L > LP || p.L < 1 || p.ldkv % 8 || p.kcol % 8 || p.vcol % 8) return false;
    if (p.mode == 1 && p.NM > LP) return false;   // rows beyond L would need explicit zero gradients
  }
  return true;
}

int attn_mma_fwd(cudaStream_t s, const AttnP& p) {
  if (p.mode == 2) return launch(attn_mma_fwd_kernel<KB, false>, s, p, false);
  return launch(attn_mma_fwd_kernel<KLIN + KB, true>, s, p, false);
}
int attn_mma_bwd(cudaStream_t s, const AttnP& p) {
  if (p.mode == 2) return launch(attn_mma_bwd_kernel<KB, false>, s, p, true);
  return launch(attn_mma_bwd_kernel<KLIN + KB, true>, s, p, true);
}
