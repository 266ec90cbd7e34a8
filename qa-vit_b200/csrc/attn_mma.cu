// Tensor-core attention for the 16-query problems of the quad block (bf16 runs): one WARP per (window, head) task,
// mma.sync.m16n8k16 (bf16 x bf16 -> fp32), every intermediate kept in registers.
//
// Why mma.sync and not tcgen05 here: a task is Q[16 x 48] against 48 (or 16) keys -- one m16 tile.  A 128-row UMMA
// tile would have to stack 8 unrelated tasks block-diagonally (8x wasted MMA) and round-trip S/P through TMEM;
// m16n8k16 matches the problem exactly and keeps S, P, dS in registers (SURVEY.md 7.2 hard part 3).  The SIMT kernels
// in attn.cu stay as the fp32 parity path and for shapes outside this file's (nq = 16, L <= 16, k = 32, bank = 16).
//
// The Linformer projections K' = E_k^T Ks, V' = E_v^T Vs (H:332-352) are never materialised.  With T = Q Ks^T [16 x 16]:
//   forward : S = [T E_k | Q Bk^T] / sqrt(hd) -> P = softmax(S) (+ dropout) -> U = P_lin E_v^T -> O = U Vs + P_bank Bv
//   backward: W = dO Vs^T; dP = [W E_v | dO Bv^T]; dS = P (dP - rowsum(P dP)) / sqrt(hd); X = dS_lin E_k^T;
//             dQ = X Ks + dS_bank Bk;  dKs = X^T Q;  dVs = U^T dO;  dE_k += T^T dS_lin;  dE_v += W^T P_lin;
//             dBk += dS_bank^T Q;  dBv += P_bank^T dO
// (Bk / Bv = the bank snapshot of this head, or the projected bank K / V of the cross branch, where the Linformer half
// is absent.)  An mma C fragment of a [16 x 16] product IS the A fragment of the next product (same thread <-> element
// map after packing to bf16), and its transpose is four movmatrix instructions, so the chain never touches shared memory:
// 84 MMAs per backward task against 180 for the version that formed K', V', dK', dV' in shared memory, no intra-task
// __syncwarp, and 7 KB instead of 21 KB of shared memory per warp.  The task's inputs (Q, Ks, Vs, dO: 16 x 48 bf16 each)
// are fetched with cp.async into a double buffer one task ahead.
// Batch reductions (dE, d bank) accumulate in mma C registers across a warp's tasks and are flushed once per warp.
#include "kernels.h"

#ifndef ATTB_M
#define ATTB_M 3   // CTAs per SM the backward kernel is compiled for (A/B knob)
#endif

namespace {

constexpr int HD = 48, NQ = 16, LP = 16, KLIN = 32, KB = 16;
constexpr int PT = 56;    // pitch (bf16 elements) of 48-wide rows: 112 B -> conflict-free 16 B row accesses
constexpr int PE = 40;    // pitch of the 32-wide Linformer matrices
constexpr int WARPS = 4;
constexpr int MAT = NQ * PT;   // one staged [16 x 48] matrix

__device__ __forceinline__ uint32_t sa(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm4(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ uint32_t movt(uint32_t a) {   // transpose of an 8 x 8 bf16 tile held one row-pair per thread
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
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
__device__ __forceinline__ void cp16(bf16* dst, const bf16* src, bool valid) {   // 16 B global -> shared, zero-filled if !valid
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa(dst)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// A fragment (16 x 16) from smem stored [m][k] (k contiguous)
__device__ __forceinline__ void ldA(uint32_t* a, const bf16* base, int pitch, int k0, int lane) {
  const int mat = lane >> 3, r = lane & 7;
  ldsm4(a, sa(base + (r + (mat & 1) * 8) * pitch + k0 + (mat >> 1) * 8));
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
// two adjacent C tiles (16 x 16, fp32) -> the A fragment of the same matrix in bf16
__device__ __forceinline__ void packA(uint32_t* a, const float* c0, const float* c1) {
  a[0] = pack2(c0[0], c0[1]); a[1] = pack2(c0[2], c0[3]); a[2] = pack2(c1[0], c1[1]); a[3] = pack2(c1[2], c1[3]);
}
// A fragment of M^T from the A fragment of M (16 x 16)
__device__ __forceinline__ void transA(uint32_t* t, const uint32_t* a) {
  t[0] = movt(a[0]); t[1] = movt(a[2]); t[2] = movt(a[1]); t[3] = movt(a[3]);
}
// C[16 x 16] (two n8 tiles) = A[16 x 48] * B^T, B staged [n = 16][k = 48]
__device__ __forceinline__ void mm_k48_n16(float (*c)[4], const uint32_t (*a)[4], const bf16* Bs, int lane) {
#pragma unroll
  for (int e = 0; e < 4; ++e) c[0][e] = c[1][e] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 3; ++kk) {
    uint32_t b[4];
    ldB(b, Bs, PT, 0, kk * 16, lane);
    mma16816(c[0], a[kk], b[0], b[1]);
    mma16816(c[1], a[kk], b[2], b[3]);
  }
}
// o[16 x 48] += A[16 x 16] * B, B staged [k = 16][n = 48]
__device__ __forceinline__ void mm_k16_n48(float (*o)[4], const uint32_t* a, const bf16* Bs, int lane) {
#pragma unroll
  for (int np = 0; np < 3; ++np) {
    uint32_t b[4];
    ldBt(b, Bs, PT, np * 16, 0, lane);
    mma16816(o[2 * np], a, b[0], b[1]);
    mma16816(o[2 * np + 1], a, b[2], b[3]);
  }
}
// o[16 x 48] += A[16 x 16] * B with B already in fragment registers bf[3][4] (ldBt of a [k = 16][n = 48] matrix)
__device__ __forceinline__ void mm_k16_n48_r(float (*o)[4], const uint32_t* a, const uint32_t (*bf)[4]) {
#pragma unroll
  for (int np = 0; np < 3; ++np) {
    mma16816(o[2 * np], a, bf[np][0], bf[np][1]);
    mma16816(o[2 * np + 1], a, bf[np][2], bf[np][3]);
  }
}
// c[16 x 32] (4 tiles, overwritten) = A[16 x 16] * E, E staged [k = l][n = j] (j contiguous)
__device__ __forceinline__ void mm_times_E(float (*c)[4], const uint32_t* a, const bf16* E, int lane) {
#pragma unroll
  for (int np = 0; np < 2; ++np) {
    uint32_t b[4];
    ldBt(b, E, PE, np * 16, 0, lane);
#pragma unroll
    for (int e = 0; e < 4; ++e) c[2 * np][e] = c[2 * np + 1][e] = 0.f;
    mma16816(c[2 * np], a, b[0], b[1]);
    mma16816(c[2 * np + 1], a, b[2], b[3]);
  }
}
// c[16 x 16] (2 tiles) = X[16 x 32] * E^T, X given as its two A fragments, E staged [n = l][k = j]
__device__ __forceinline__ void mm_times_Et(float (*c)[4], const uint32_t (*xa)[4], const bf16* E, int lane) {
#pragma unroll
  for (int e = 0; e < 4; ++e) c[0][e] = c[1][e] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    uint32_t b[4];
    ldB(b, E, PE, 0, kk * 16, lane);
    mma16816(c[0], xa[kk], b[0], b[1]);
    mma16816(c[1], xa[kk], b[2], b[3]);
  }
}
// acc[16 x 32] += M^T Y: mt = A fragment of M^T, Y [16 x 32] given as its two A fragments ya[2][4] (k = rows of Y)
__device__ __forceinline__ void acc_Mt_Y(float (*acc)[4], const uint32_t* mt, const uint32_t (*ya)[4]) {
#pragma unroll
  for (int n = 0; n < 4; ++n)   // B tile n (columns 8 n ..): rows 0-7 sit in ya[n / 2][2 (n & 1)], rows 8-15 in the next register
    mma16816(acc[n], mt, movt(ya[n >> 1][(n & 1) * 2]), movt(ya[n >> 1][(n & 1) * 2 + 1]));
}

__device__ __forceinline__ int q_row(const AttnP& p, int w, int i) {
  if (p.mode == 0 && p.side != p.ws) {
    const int nws = p.side / p.ws, nW = nws * nws;
    const int b = w / nW, wi = w % nW;
    return b * p.Nt + ((wi / nws) * p.ws + i / p.ws) * p.side + (wi % nws) * p.ws + i % p.ws;
  }
  return w * p.Nt + i;
}

// 16 x 48 fp32 (row pitch D) -> bf16 smem rows (once per warp: the bank tile of the warp's head)
__device__ __forceinline__ void load_bank(bf16* dst, const float* src, int D, int col, int lane) {
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int c = lane + 32 * k, i = c / 12, ch = c % 12;
    const float4 v = *reinterpret_cast<const float4*>(src + i * D + col + ch * 4);
    *reinterpret_cast<uint2*>(dst + i * PT + ch * 4) = make_uint2(pack2(v.x, v.y), pack2(v.z, v.w));
  }
}
// E_k / E_v fp32 [L][32] -> bf16 smem [LP][PE] (rows >= L zero): once per CTA
__device__ __forceinline__ void load_E(bf16* Es, const AttnP& p) {
  for (int idx = threadIdx.x; idx < 2 * LP * KLIN; idx += blockDim.x) {
    const int which = idx / (LP * KLIN), r = idx % (LP * KLIN), l = r / KLIN, j = r % KLIN;
    const float* src = which ? p.Ev : p.Ek;
    Es[which * LP * PE + l * PE + j] = __float2bfloat16_rn(l < p.L ? src[l * KLIN + j] : 0.f);
  }
}

// cp.async of one task's inputs into a staging buffer: [Q | Ks | Vs | dO], each [16][PT]; key / value rows >= L zero-filled
template <bool LINF, bool BWD>
__device__ __forceinline__ void prefetch_task(const AttnP& p, bf16* buf, int w, int h, int lane) {
  const bf16* q = static_cast<const bf16*>(p.q);
  const bf16* kv = static_cast<const bf16*>(p.kv);
  const bf16* dout = static_cast<const bf16*>(p.dout);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int c = lane + 32 * k, i = c / 6, ch = c % 6;
    const int rq = q_row(p, w, i);
    cp16(buf + i * PT + ch * 8, q + (long)rq * p.ldq + p.qcol + h * HD + ch * 8, true);
    if (BWD) cp16(buf + 3 * MAT + i * PT + ch * 8, dout + (long)rq * p.lddo + h * HD + ch * 8, true);
    if (LINF) {
      const int rk = (p.mode == 0) ? rq : w * p.NM + min(i, p.NM - 1);
      const bool ok = i < p.L;
      cp16(buf + MAT + i * PT + ch * 8, kv + (long)rk * p.ldkv + p.kcol + h * HD + ch * 8, ok);
      cp16(buf + 2 * MAT + i * PT + ch * 8, kv + (long)rk * p.ldkv + p.vcol + h * HD + ch * 8, ok);
    }
  }
  cp_commit();
}

// C-layout [16 x 48] registers -> global bf16 rows r0 (fragment rows g) / r1 (rows g + 8); a row < 0 is skipped
__device__ __forceinline__ void store_rows(bf16* dst, long ld, int col, int r0, int r1, const float (*o)[4], int t) {
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) {
    if (r0 >= 0) *reinterpret_cast<uint32_t*>(dst + (long)r0 * ld + col + n * 8 + 2 * t) = pack2(o[n][0], o[n][1]);
    if (r1 >= 0) *reinterpret_cast<uint32_t*>(dst + (long)r1 * ld + col + n * 8 + 2 * t) = pack2(o[n][2], o[n][3]);
  }
}

// softmax over the NT key tiles of a C-layout score tile (rows g and g + 8)
template <int NT>
__device__ __forceinline__ void softmax_c(float (*s)[4], float scale) {
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

// scores -> probabilities P[NT][4] (tiles [0, 4) = Linformer keys, last two = bank keys); ta = bf16 A fragment of T = Q Ks^T
template <bool LINF>
__device__ __forceinline__ void probabilities(float (*P)[4], uint32_t* ta, const uint32_t (*aq)[4], const bf16* Ks, const bf16* Bk,
                                              const bf16* Ek, float scale, int lane) {
  constexpr int LT = LINF ? 4 : 0;
  mm_k48_n16(P + LT, aq, Bk, lane);
  if (LINF) {
    float T[2][4];
    mm_k48_n16(T, aq, Ks, lane);
    packA(ta, T[0], T[1]);
    mm_times_E(P, ta, Ek, lane);
  }
  softmax_c<LT + 2>(P, scale);
}

template <bool LINF>
__global__ void __launch_bounds__(WARPS * 32, 3) attn_mma_fwd_kernel(AttnP p, int ntask) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  constexpr int NT = LINF ? 6 : 2, LT = LINF ? 4 : 0, NM = LINF ? 3 : 1;   // NM staged matrices per task: Q [, Ks, Vs]
  bf16* Es = reinterpret_cast<bf16*>(smraw);                               // [2][LP][PE]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  bf16* Bk = Es + 2 * LP * PE + warp * 2 * MAT;                            // this warp's head: bank K | bank V
  bf16* Bv = Bk + MAT;
  bf16* stage = Es + 2 * LP * PE + WARPS * 2 * MAT + warp * 2 * NM * MAT;  // [2 buffers][NM][16][PT]
  const int D = p.H * HD, h = warp;                                        // H == WARPS and the task stride is a multiple of it
  if (LINF) load_E(Es, p);
  load_bank(Bk, LINF ? p.bank_k : p.kc, D, h * HD, lane);
  load_bank(Bv, LINF ? p.bank_v : p.vc, D, h * HD, lane);
  __syncthreads();
  const float scale = rsqrtf((float)HD);
  bf16* out = static_cast<bf16*>(p.out);
  const int stride = gridDim.x * WARPS;
  int task = blockIdx.x * WARPS + warp, cur = 0;
  if (task < ntask) prefetch_task<LINF, false>(p, stage, task / p.H, h, lane);
  for (; task < ntask; task += stride, cur ^= 1) {
    const int w = task / p.H;
    bf16* buf = stage + cur * NM * MAT;
    if (task + stride < ntask) {
      prefetch_task<LINF, false>(p, stage + (cur ^ 1) * NM * MAT, (task + stride) / p.H, h, lane);
      cp_wait<1>();
    } else {
      cp_wait<0>();
    }
    __syncwarp();
    uint32_t aq[3][4];
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) ldA(aq[kk], buf, PT, kk * 16, lane);
    float P[NT][4];
    uint32_t ta[4];
    probabilities<LINF>(P, ta, aq, buf + MAT, Bk, Es, scale, lane);
    if (p.drop.p > 0.f) {   // SDPA dropout_p on the probabilities
      const DropState ds = drop_state(p.drop);
      drop_apply_c<NT>(P, drop_bits_c<NT>(ds, (uint32_t)task, lane), ds.inv);
    }
    float o[6][4];
#pragma unroll
    for (int n = 0; n < 6; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    uint32_t pb[4];
    packA(pb, P[LT], P[LT + 1]);
    mm_k16_n48(o, pb, Bv, lane);
    if (LINF) {
      uint32_t pa[2][4], ua[4];
      float U[2][4];
      packA(pa[0], P[0], P[1]);
      packA(pa[1], P[2], P[3]);
      mm_times_Et(U, pa, Es + LP * PE, lane);
      packA(ua, U[0], U[1]);
      mm_k16_n48(o, ua, buf + 2 * MAT, lane);
    }
    store_rows(out, p.ldo, h * HD, q_row(p, w, g), q_row(p, w, g + 8), o, t);
    __syncwarp();   // every lane is done with `buf` before the next iteration's prefetch overwrites it
  }
}

template <bool LINF>
__global__ void __launch_bounds__(WARPS * 32, ATTB_M) attn_mma_bwd_kernel(AttnP p, int ntask) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  constexpr int NT = LINF ? 6 : 2, LT = LINF ? 4 : 0;
  constexpr int NM = 4;                                                    // Q, Ks, Vs, dO (cross: Ks / Vs slots unused)
  bf16* Es = reinterpret_cast<bf16*>(smraw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  bf16* Bk = Es + 2 * LP * PE + warp * 2 * MAT;
  bf16* Bv = Bk + MAT;
  bf16* stage = Es + 2 * LP * PE + WARPS * 2 * MAT + warp * 2 * NM * MAT;
  // backward: a CTA serves ONE head (its warps take different windows), so the batch-reduction tiles of its 4 warps are summed in
  // shared memory before they go to global memory -- a quarter of the reductions of a warp-per-head mapping
  const int D = p.H * HD, h = blockIdx.x % p.H;
  if (LINF) load_E(Es, p);
  load_bank(Bk, LINF ? p.bank_k : p.kc, D, h * HD, lane);
  load_bank(Bv, LINF ? p.bank_v : p.vc, D, h * HD, lane);
  __syncthreads();
  const bf16 *Ek = Es, *Ev = Es + LP * PE;
  const float scale = rsqrtf((float)HD);
  bf16* dq = static_cast<bf16*>(p.dq);
  bf16* dkv = static_cast<bf16*>(p.dkv);
  // batch reductions live in registers: this warp's head is fixed, so its bank-row gradients are one [16 x 48] C tile per
  // K and V; dE_k / dE_v are [16 x 32] tiles
  float dbk[6][4], dbv[6][4], dEk[4][4], dEv[4][4];
#pragma unroll
  for (int n = 0; n < 6; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) dbk[n][e] = dbv[n][e] = 0.f;
#pragma unroll
  for (int n = 0; n < 4; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) dEk[n][e] = dEv[n][e] = 0.f;
  const int nwin = ntask / p.H, wstride = (gridDim.x / p.H) * WARPS;      // the grid is a multiple of H (launch_bwd)
  int w = (blockIdx.x / p.H) * WARPS + warp, cur = 0;
  if (w < nwin) prefetch_task<LINF, true>(p, stage, w, h, lane);
  for (; w < nwin; w += wstride, cur ^= 1) {
    const int task = w * p.H + h;                                          // the forward kernel's task id (dropout element ids)
    bf16* buf = stage + cur * NM * MAT;
    const bf16 *Qs = buf, *Ks = buf + MAT, *Vs = buf + 2 * MAT, *DOs = buf + 3 * MAT;
    if (w + wstride < nwin) {
      prefetch_task<LINF, true>(p, stage + (cur ^ 1) * NM * MAT, w + wstride, h, lane);
      cp_wait<1>();
    } else {
      cp_wait<0>();
    }
    __syncwarp();
    float P[NT][4], dS[NT][4];
    uint32_t ta[4], wa[4];
    {
      uint32_t aq[3][4];
#pragma unroll
      for (int kk = 0; kk < 3; ++kk) ldA(aq[kk], Qs, PT, kk * 16, lane);
      probabilities<LINF>(P, ta, aq, Ks, Bk, Ek, scale, lane);
    }
    {   // dP = [W E_v | dO Bv^T],  W = dO Vs^T
      uint32_t ado[3][4];
#pragma unroll
      for (int kk = 0; kk < 3; ++kk) ldA(ado[kk], DOs, PT, kk * 16, lane);
      mm_k48_n16(dS + LT, ado, Bv, lane);
      if (LINF) {
        float Wm[2][4];
        mm_k48_n16(Wm, ado, Vs, lane);
        packA(wa, Wm[0], Wm[1]);
        mm_times_E(dS, wa, Ev, lane);
      }
    }
    unsigned long long keep = ~0ull;
    float kinv = 1.f;
    if (p.drop.p > 0.f) {   // dP <- d(P_dropped) * keep; P_dropped (for dV) is formed after the softmax backward
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
    // from here on P and dS are only MMA operands: keep their bf16 A fragments (k-step kk = key tiles 2 kk, 2 kk + 1)
    uint32_t pP[NT / 2][4], pS[NT / 2][4];
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
      packA(pP[kk], P[2 * kk], P[2 * kk + 1]);
      packA(pS[kk], dS[2 * kk], dS[2 * kk + 1]);
    }
    const int rq0 = q_row(p, w, g), rq1 = q_row(p, w, g + 8);
    uint32_t xa[4];
    {   // dQ = X Ks + dS_bank Bk,  X = dS_lin E_k^T
      float o[6][4];
#pragma unroll
      for (int n = 0; n < 6; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
      mm_k16_n48(o, pS[LT / 2], Bk, lane);
      if (LINF) {
        float X[2][4];
        mm_times_Et(X, pS, Ek, lane);
        packA(xa, X[0], X[1]);
        mm_k16_n48(o, xa, Ks, lane);
      }
      store_rows(dq, p.lddq, p.dqcol + h * HD, rq0, rq1, o, t);
    }
    uint32_t bq[3][4], bdo[3][4], tr[4];   // Q and dO as [k = query][n = channel] B operands
#pragma unroll
    for (int np = 0; np < 3; ++np) {
      ldBt(bq[np], Qs, PT, np * 16, 0, lane);
      ldBt(bdo[np], DOs, PT, np * 16, 0, lane);
    }
    transA(tr, pS[LT / 2]);     // d bank_k += dS_bank^T Q
    mm_k16_n48_r(dbk, tr, bq);
    transA(tr, pP[LT / 2]);     // d bank_v += P_bank^T dO
    mm_k16_n48_r(dbv, tr, bdo);
    if (LINF) {
      int rk0, rk1;
      if (p.mode == 0) { rk0 = rq0; rk1 = rq1; }
      else { rk0 = w * p.NM + min(g, p.NM - 1); rk1 = w * p.NM + min(g + 8, p.NM - 1); }
      if (g >= p.L) rk0 = -1;
      if (g + 8 >= p.L) rk1 = -1;
      float o[6][4];
      // dKs = X^T Q
#pragma unroll
      for (int n = 0; n < 6; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
      transA(tr, xa);
      mm_k16_n48_r(o, tr, bq);
      store_rows(dkv, p.lddkv, p.dkcol + h * HD, rk0, rk1, o, t);
      // dVs = U^T dO,  U = P_lin E_v^T (dropped probabilities)
      float U[2][4];
      uint32_t ua[4];
      mm_times_Et(U, pP, Ev, lane);
      packA(ua, U[0], U[1]);
#pragma unroll
      for (int n = 0; n < 6; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
      transA(tr, ua);
      mm_k16_n48_r(o, tr, bdo);
      store_rows(dkv, p.lddkv, p.dvcol + h * HD, rk0, rk1, o, t);
      // dE_k += T^T dS_lin,  dE_v += W^T P_lin
      transA(tr, ta);
      acc_Mt_Y(dEk, tr, pS);
      transA(tr, wa);
      acc_Mt_Y(dEv, tr, pP);
    }
    __syncwarp();
  }
  // ---- flush: the 4 warps' register tiles are summed in shared memory (the staging buffers are free now), then one 16 B vector
  // reduction per 4 elements and CTA.  (Scalar atomics per warp were a third of the kernel: 444 x 4 warps x 2560 elements onto 7 k
  // addresses; the cost follows the number of reduction operations, not bytes.)
  __syncthreads();
  float* acc = reinterpret_cast<float*>(Es + 2 * LP * PE + WARPS * 2 * MAT);   // [dbk 16x48 | dbv 16x48 | dEk 16x32 | dEv 16x32]
  constexpr int NB = NQ * HD, NE = LINF ? LP * KLIN : 0, NACC = 2 * NB + 2 * NE;
  for (int i = threadIdx.x; i < NACC; i += WARPS * 32) acc[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int n = 0; n < 6; ++n) {
    float* k0 = acc + g * HD + n * 8 + 2 * t;
    atomicAdd(k0, dbk[n][0]); atomicAdd(k0 + 1, dbk[n][1]); atomicAdd(k0 + 8 * HD, dbk[n][2]); atomicAdd(k0 + 8 * HD + 1, dbk[n][3]);
    float* v0 = k0 + NB;
    atomicAdd(v0, dbv[n][0]); atomicAdd(v0 + 1, dbv[n][1]); atomicAdd(v0 + 8 * HD, dbv[n][2]); atomicAdd(v0 + 8 * HD + 1, dbv[n][3]);
  }
  if (LINF) {
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      float* k0 = acc + 2 * NB + g * KLIN + n * 8 + 2 * t;
      atomicAdd(k0, dEk[n][0]); atomicAdd(k0 + 1, dEk[n][1]); atomicAdd(k0 + 8 * KLIN, dEk[n][2]); atomicAdd(k0 + 8 * KLIN + 1, dEk[n][3]);
      float* v0 = k0 + NE;
      atomicAdd(v0, dEv[n][0]); atomicAdd(v0 + 1, dEv[n][1]); atomicAdd(v0 + 8 * KLIN, dEv[n][2]); atomicAdd(v0 + 8 * KLIN + 1, dEv[n][3]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NACC / 4; i += WARPS * 32) {
    const float4 v = reinterpret_cast<const float4*>(acc)[i];
    const int e = i * 4;
    float* dst;
    if (e < 2 * NB) {
      const int r = (e % NB) / HD, c = e % HD;
      dst = (e < NB ? p.dbank_k : p.dbank_v) + r * D + h * HD + c;
    } else {
      constexpr int NE1 = NE ? NE : 1;                       // (no Linformer tiles: branch not reached)
      const int f = e - 2 * NB, r = (f % NE1) / KLIN, c = f % KLIN;
      if (r >= p.L) continue;
      dst = (f < NE ? p.dEk : p.dEv) + r * KLIN + c;
    }
    red_add_v4(dst, v.x, v.y, v.z, v.w);
  }
}

template <typename K>
int launch(K kernel, cudaStream_t s, const AttnP& p, int staged) {
  const int nwin = (p.mode == 0) ? p.B * (p.side / p.ws) * (p.side / p.ws) : p.B;
  const int ntask = nwin * p.H;
  if (ntask <= 0) return 0;
  const size_t smem = (size_t)(2 * LP * PE + WARPS * 2 * MAT + WARPS * 2 * staged * MAT) * sizeof(bf16);
  if (smem > 48 * 1024) QV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = min(cdiv(ntask, WARPS), qv_num_sms() * 3);
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
  if (p.mode == 2) return launch(attn_mma_fwd_kernel<false>, s, p, 1);
  return launch(attn_mma_fwd_kernel<true>, s, p, 3);
}
template <typename K>
int launch_bwd(K kernel, cudaStream_t s, const AttnP& p) {   // as launch(), with the grid a multiple of H: a CTA serves one head
  const int nwin = (p.mode == 0) ? p.B * (p.side / p.ws) * (p.side / p.ws) : p.B;
  const int ntask = nwin * p.H;
  if (ntask <= 0) return 0;
  const size_t smem = (size_t)(2 * LP * PE + WARPS * 2 * MAT + WARPS * 2 * 4 * MAT) * sizeof(bf16);
  if (smem > 48 * 1024) QV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int per_head = max(1, min(cdiv(nwin, WARPS), qv_num_sms() * 3 / p.H));
  qv_launch(kernel, per_head * p.H, WARPS * 32, smem, s, p, ntask);
  QV_LAUNCH_CHECK();
  return 0;
}

int attn_mma_bwd(cudaStream_t s, const AttnP& p) {
  auto al16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };   // the flush uses 16 B vector reductions
  QV_CHECK(al16(p.dbank_k) && al16(p.dbank_v) && (p.mode == 2 || (al16(p.dEk) && al16(p.dEv))),
           "attn_mma_bwd: bank / Linformer gradient buffers must be 16 B aligned");
  if (p.mode == 2) return launch_bwd(attn_mma_bwd_kernel<false>, s, p);
  return launch_bwd(attn_mma_bwd_kernel<true>, s, p);
}
