// TokenLearner / TokenUpMix (H:971-1031) on tensor cores for bf16 runs with 16 learned tokens.
//
// Per image these are four skinny products with one dimension = 16 (the learned tokens):
//   TokenLearner fwd : xc[16, C] = S^T[16, N] x[N, C]                     S = softmax over tokens of the gate logits
//   TokenLearner bwd : dS[N, 16] = x[N, C] dxc^T[C, 16];  dx[N, C] = S[N, 16] dxc[16, C]
//   TokenUpMix  fwd  : up[N, C] = W[N, 16] xc[16, C] + bias
//   TokenUpMix  bwd  : dxc[16, C] = W^T[16, N] dup[N, C];  dW[N, 16] += dup[N, C] xc^T[C, 16];  dbias += rowsum(dup)
// The reference runs them as torch.bmm / nn.Linear under autocast, i.e. with bf16 operands and fp32 accumulation
// (SURVEY appendix C): exactly mma.sync.m16n8k16.  One 128-thread CTA walks images; the fp32 stream tile is converted
// to bf16 once into shared memory and read with ldmatrix; outputs leave the accumulator fragments as fp32 float2
// stores.  A 16-wide problem has no use for a 128-row tcgen05 tile (8 images would have to be stacked block-
// diagonally), which is why this is warp-level MMA.  HBM-bound in the ideal: one pass over x / dx / up / dup.
// The FORWARD kernels keep their operands as bf16 PAIRS (hi + lo, v = hi + lo to ~16 mantissa bits) and compute every
// product as three MMAs (hi*hi + lo*hi + hi*lo): the token stream is the fp32 residual stream, and rounding it to 8
// mantissa bits in the forward pass moved the whole-model bf16 logits from 0.9e-2 to 1.4-1.9e-2 of the fp32 run (the
// north-star gate is 1e-2).  The backward kernels use plain bf16 operands, like the reference's autocast bmm / Linear.
// fp32 runs and other shapes use the SIMT kernels of tokens.cu / misc.cu.
#include "kernels.h"

namespace {

constexpr int M16 = 16;
constexpr int SP = 24;      // pitch (bf16) of the [N][16] matrices: 48 B rows, conflict-free for ldmatrix
constexpr int NWARP = 4;

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
// B of two adjacent n8 tiles from smem [n][k] (k contiguous)
__device__ __forceinline__ void ldB(uint32_t* b, const bf16* base, int pitch, int n0, int k0, int lane) {
  const int mat = lane >> 3, r = lane & 7;
  ldsm4(b, sa(base + (n0 + r + (mat >> 1) * 8) * pitch + k0 + (mat & 1) * 8));
}
// same from smem [k][n] (n contiguous)
__device__ __forceinline__ void ldBt(uint32_t* b, const bf16* base, int pitch, int n0, int k0, int lane) {
  const int mat = lane >> 3, r = lane & 7;
  ldsm4t(b, sa(base + (k0 + r + (mat & 1) * 8) * pitch + n0 + (mat >> 1) * 8));
}

__device__ __forceinline__ void split2(float x, float y, uint32_t* hi, uint32_t* lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  const float2 hf = __bfloat1622float2(h);
  *hi = *reinterpret_cast<const uint32_t*>(&h);
  *lo = pack2(x - hf.x, y - hf.y);
}
// fp32 global [rows][C] -> bf16 hi (/ lo when HL) smem tiles [rows][pitch].  (row, column) of a thread's elements advance
// incrementally: the i / c4n, i % c4n pair this loop used to evaluate per element was half of all instructions the token
// kernels executed (ncu source view, profiles/r1_ncu_brief_prof_tok_b4736.txt).
template <bool HL>
__device__ __forceinline__ void tile_to_bf16(bf16* dhi, bf16* dlo, int pitch, const float* __restrict__ src, int rows, int C) {
  const int c4n = C / 4, total = rows * c4n;
  const int step = blockDim.x, srow = step / c4n, scol = step - srow * c4n;
  int row = threadIdx.x / c4n, col = threadIdx.x - row * c4n;
  for (int base = 0; base < total; base += 8 * step) {     // 8 independent 16 B loads in flight per thread
    float4 v[8];
    int o[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + u * step + threadIdx.x;
      if (i < total) v[u] = *reinterpret_cast<const float4*>(src + (long)i * 4);   // full rows: the tile is contiguous
      o[u] = row * pitch + col * 4;
      row += srow; col += scol;
      if (col >= c4n) { col -= c4n; ++row; }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + u * step + threadIdx.x;
      if (i < total) {
        uint32_t h0, l0, h1, l1;
        split2(v[u].x, v[u].y, &h0, &l0);
        split2(v[u].z, v[u].w, &h1, &l1);
        *reinterpret_cast<uint2*>(dhi + o[u]) = make_uint2(h0, h1);
        if (HL) *reinterpret_cast<uint2*>(dlo + o[u]) = make_uint2(l0, l1);
      }
    }
  }
}
template <bool HL>
__device__ __forceinline__ void store_split(bf16* hi, bf16* lo, int idx, float v) {
  const bf16 h = __float2bfloat16_rn(v);
  hi[idx] = h;
  if (HL) lo[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
}
// acc += (ah + al) (bh + bl) without the lo * lo term; HL == false: plain bf16 product
template <bool HL>
__device__ __forceinline__ void mma3(float* acc, const uint32_t* ah, const uint32_t* al, uint32_t bh0, uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma16816(acc, ah, bh0, bh1);
  if (HL) {
    mma16816(acc, al, bh0, bh1);
    mma16816(acc, ah, bl0, bl1);
  }
}

struct Smem {
  bf16 *X, *XL;   // [N][XP]   stream tile (x or dup), hi / lo
  bf16 *D, *DL;   // [16][XP]  compressed tile (xc or dxc)
  bf16 *S, *SL;   // [N][SP]   S or W
  float* F;       // [N][16]   fp32 S / logits
  float* R;       // [8][16] + [16] reductions
};
__device__ __host__ __forceinline__ size_t x_elems(int N, int XP, bool needX) { return needX ? (size_t)N * XP : 0; }
__device__ __host__ __forceinline__ size_t d_elems(int XP, bool needD) { return needD ? (size_t)M16 * XP : 0; }
__device__ __forceinline__ Smem carve(uint8_t* base, int N, int XP, bool hl, bool needX = true, bool needD = true) {
  Smem s;
  const size_t k = hl ? 1 : 0;             // lo tiles exist only in the split-precision kernels
  s.X = reinterpret_cast<bf16*>(base);
  s.XL = s.X + k * x_elems(N, XP, needX);
  s.D = s.XL + x_elems(N, XP, needX);
  s.DL = s.D + k * d_elems(XP, needD);
  s.S = s.DL + d_elems(XP, needD);
  s.SL = s.S + k * (size_t)N * SP;
  s.F = reinterpret_cast<float*>(s.SL + (size_t)N * SP);
  s.R = s.F + (size_t)N * M16;
  return s;
}
size_t smem_bytes(int N, int C, bool hl, bool needX = true, bool needD = true) {
  const int XP = C + 8;
  return (hl ? 2 : 1) * (x_elems(N, XP, needX) + d_elems(XP, needD) + (size_t)N * SP) * 2 + ((size_t)N * M16 + 9 * M16) * 4 + 16;
}

// ---------------------------------------------------------------------------------------------- TokenLearner forward
__global__ void __launch_bounds__(NWARP * 32) tlm_fwd_kernel(const float* __restrict__ x, const bf16* __restrict__ logits, int B, int N,
                                                             int C, float* __restrict__ Sout, float* __restrict__ xc) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  const int XP = C + 8;
  Smem sm = carve(smraw, N, XP, true, true, false);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int col = tid % M16, part = tid / M16;          // softmax: 8 threads per slot
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    for (int i = tid; i < N * M16 / 8; i += blockDim.x) {      // 16 B loads (the scalar loop was 15 % of the stall samples)
      const uint4 q = *reinterpret_cast<const uint4*>(logits + (long)b * N * M16 + i * 8);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
        sm.F[i * 8 + 2 * j] = f.x; sm.F[i * 8 + 2 * j + 1] = f.y;
      }
    }
    tile_to_bf16<true>(sm.X, sm.XL, XP, x + (long)b * N * C, N, C);
    __syncthreads();
    float mx = -INFINITY;
    for (int n = part; n < N; n += 8) mx = fmaxf(mx, sm.F[n * M16 + col]);
    sm.R[part * M16 + col] = mx;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) mx = fmaxf(mx, sm.R[k * M16 + col]);
    float z = 0.f;
    for (int n = part; n < N; n += 8) { const float e = __expf(sm.F[n * M16 + col] - mx); sm.F[n * M16 + col] = e; z += e; }
    __syncthreads();
    sm.R[part * M16 + col] = z;
    __syncthreads();
    z = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) z += sm.R[k * M16 + col];
    z = 1.f / z;
    for (int n = part; n < N; n += 8) {
      const float s = sm.F[n * M16 + col] * z;
      Sout[(long)b * N * M16 + n * M16 + col] = s;
      store_split<true>(sm.S, sm.SL, n * SP + col, s);
    }
    __syncthreads();
    // xc[16, C] = S^T x : A = S^T (from [k = token][m = slot]), B = x (from [k = token][n = channel])
    for (int pair = warp; pair < C / 16; pair += NWARP) {
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      for (int ks = 0; ks < N / 16; ++ks) {
        uint32_t a[4], al[4], bb[4], bl[4];
        ldAt(a, sm.S, SP, 0, ks * 16, lane);
        ldAt(al, sm.SL, SP, 0, ks * 16, lane);
        ldBt(bb, sm.X, XP, pair * 16, ks * 16, lane);
        ldBt(bl, sm.XL, XP, pair * 16, ks * 16, lane);
        mma3<true>(acc[0], a, al, bb[0], bb[1], bl[0], bl[1]);
        mma3<true>(acc[1], a, al, bb[2], bb[3], bl[2], bl[3]);
      }
      float* o = xc + (long)b * M16 * C + pair * 16 + 2 * t;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        *reinterpret_cast<float2*>(o + (long)g * C + h * 8) = make_float2(acc[h][0], acc[h][1]);
        *reinterpret_cast<float2*>(o + (long)(g + 8) * C + h * 8) = make_float2(acc[h][2], acc[h][3]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- TokenLearner backward
__global__ void __launch_bounds__(NWARP * 32) tlm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ S,
                                                             const float* __restrict__ dxc, int B, int N, int C,
                                                             bf16* __restrict__ dlogits, float* __restrict__ dx) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  const int XP = C + 8;
  Smem sm = carve(smraw, N, XP, false);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    for (int i = tid; i < N * M16; i += blockDim.x) {
      const float s = S[(long)b * N * M16 + i];
      sm.F[i] = s;
      store_split<false>(sm.S, sm.SL, (i / M16) * SP + i % M16, s);
    }
    if (tid < M16) sm.R[tid] = 0.f;
    tile_to_bf16<false>(sm.X, sm.XL, XP, x + (long)b * N * C, N, C);
    tile_to_bf16<false>(sm.D, sm.DL, XP, dxc + (long)b * M16 * C, M16, C);
    __syncthreads();
    // dS[N, 16] = x dxc^T (this warp's token tiles), kept in registers until the column sums are known
    constexpr int MAXT = 2;                 // token tiles per warp: N <= 128
    float dS[MAXT][2][4];
#pragma unroll
    for (int nt = 0; nt < MAXT; ++nt) {
      const int mt = warp + nt * NWARP;
      if (mt >= N / 16) break;
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < 4; ++j) dS[nt][h][j] = 0.f;
      for (int ks = 0; ks < C / 16; ++ks) {
        uint32_t a[4], al[4] = {0u, 0u, 0u, 0u}, bb[4], bl[4] = {0u, 0u, 0u, 0u};
        ldA(a, sm.X, XP, mt * 16, ks * 16, lane);
        ldB(bb, sm.D, XP, 0, ks * 16, lane);
        mma3<false>(dS[nt][0], a, al, bb[0], bb[1], bl[0], bl[1]);
        mma3<false>(dS[nt][1], a, al, bb[2], bb[3], bl[2], bl[3]);
      }
      // partial column sums of S * dS over this tile's rows
      float cs[2][2];
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = h * 8 + 2 * t + j;
          cs[h][j] = sm.F[(mt * 16 + g) * M16 + c] * dS[nt][h][j] + sm.F[(mt * 16 + g + 8) * M16 + c] * dS[nt][h][2 + j];
        }
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float v = cs[h][j];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (g == 0) atomicAdd(sm.R + h * 8 + 2 * t + j, v);
        }
    }
    __syncthreads();
#pragma unroll
    for (int nt = 0; nt < MAXT; ++nt) {
      const int mt = warp + nt * NWARP;
      if (mt >= N / 16) break;
      // dlogits = S (dS - colsum)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = h * 8 + 2 * t;
        const float t0 = sm.R[c], t1 = sm.R[c + 1];
        const int r0 = mt * 16 + g, r1 = r0 + 8;
        const uint32_t lo = pack2(sm.F[r0 * M16 + c] * (dS[nt][h][0] - t0), sm.F[r0 * M16 + c + 1] * (dS[nt][h][1] - t1));
        const uint32_t hi = pack2(sm.F[r1 * M16 + c] * (dS[nt][h][2] - t0), sm.F[r1 * M16 + c + 1] * (dS[nt][h][3] - t1));
        *reinterpret_cast<uint32_t*>(dlogits + ((long)b * N + r0) * M16 + c) = lo;
        *reinterpret_cast<uint32_t*>(dlogits + ((long)b * N + r1) * M16 + c) = hi;
      }
      // dx[N, C] = S dxc : A = S (from [m = token][k = slot]), B = dxc (from [k = slot][n = channel])
      uint32_t a[4], al[4] = {0u, 0u, 0u, 0u};
      ldA(a, sm.S, SP, mt * 16, 0, lane);
      float* o = dx + ((long)b * N + mt * 16) * C + 2 * t;
      for (int pair = 0; pair < C / 16; ++pair) {
        uint32_t bb[4], bl[4] = {0u, 0u, 0u, 0u};
        ldBt(bb, sm.D, XP, pair * 16, 0, lane);
        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        mma3<false>(acc[0], a, al, bb[0], bb[1], bl[0], bl[1]);
        mma3<false>(acc[1], a, al, bb[2], bb[3], bl[2], bl[3]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          *reinterpret_cast<float2*>(o + (long)g * C + pair * 16 + h * 8) = make_float2(acc[h][0], acc[h][1]);
          *reinterpret_cast<float2*>(o + (long)(g + 8) * C + pair * 16 + h * 8) = make_float2(acc[h][2], acc[h][3]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- TokenUpMix forward
__global__ void __launch_bounds__(NWARP * 32) upm_fwd_kernel(const float* __restrict__ xc, int B, int N, int C,
                                                             const float* __restrict__ W, const float* __restrict__ bias,
                                                             float* __restrict__ up) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  const int XP = C + 8;
  Smem sm = carve(smraw, N, XP, true, false, true);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  for (int i = tid; i < N * M16; i += blockDim.x) store_split<true>(sm.S, sm.SL, (i / M16) * SP + i % M16, W[i]);
  for (int i = tid; i < N; i += blockDim.x) sm.F[i] = bias[i];
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    tile_to_bf16<true>(sm.D, sm.DL, XP, xc + (long)b * M16 * C, M16, C);
    __syncthreads();
    for (int mt = warp; mt < N / 16; mt += NWARP) {
      uint32_t a[4], al[4];
      ldA(a, sm.S, SP, mt * 16, 0, lane);
      ldA(al, sm.SL, SP, mt * 16, 0, lane);
      const float b0 = sm.F[mt * 16 + g], b1 = sm.F[mt * 16 + g + 8];
      float* o = up + ((long)b * N + mt * 16) * C + 2 * t;
      for (int pair = 0; pair < C / 16; ++pair) {
        uint32_t bb[4], bl[4];
        ldBt(bb, sm.D, XP, pair * 16, 0, lane);
        ldBt(bl, sm.DL, XP, pair * 16, 0, lane);
        float acc[2][4] = {{b0, b0, b1, b1}, {b0, b0, b1, b1}};
        mma3<true>(acc[0], a, al, bb[0], bb[1], bl[0], bl[1]);
        mma3<true>(acc[1], a, al, bb[2], bb[3], bl[2], bl[3]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          *reinterpret_cast<float2*>(o + (long)g * C + pair * 16 + h * 8) = make_float2(acc[h][0], acc[h][1]);
          *reinterpret_cast<float2*>(o + (long)(g + 8) * C + pair * 16 + h * 8) = make_float2(acc[h][2], acc[h][3]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- TokenUpMix backward
__global__ void __launch_bounds__(NWARP * 32) upm_bwd_kernel(const float* __restrict__ xc, const float* __restrict__ dup, int B,
                                                             int N, int C, const float* __restrict__ W, float* __restrict__ dxc,
                                                             float* __restrict__ dW, float* __restrict__ dbias) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  const int XP = C + 8;
  Smem sm = carve(smraw, N, XP, false);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  for (int i = tid; i < N * M16; i += blockDim.x) store_split<false>(sm.S, sm.SL, (i / M16) * SP + i % M16, W[i]);
  constexpr int MAXT = 2;
  float aW[MAXT][2][4], aB[MAXT][4];
#pragma unroll
  for (int q = 0; q < MAXT; ++q) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { aW[q][0][j] = aW[q][1][j] = 0.f; aB[q][j] = 0.f; }
  }
  const uint32_t ones = pack2(1.f, 1.f);
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    tile_to_bf16<false>(sm.X, sm.XL, XP, dup + (long)b * N * C, N, C);
    tile_to_bf16<false>(sm.D, sm.DL, XP, xc + (long)b * M16 * C, M16, C);
    __syncthreads();
    // dxc[16, C] = W^T dup : A = W^T (from [k = token][m = slot]), B = dup (from [k = token][n = channel])
    for (int pair = warp; pair < C / 16; pair += NWARP) {
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      for (int ks = 0; ks < N / 16; ++ks) {
        uint32_t a[4], al[4] = {0u, 0u, 0u, 0u}, bb[4], bl[4] = {0u, 0u, 0u, 0u};
        ldAt(a, sm.S, SP, 0, ks * 16, lane);
        ldBt(bb, sm.X, XP, pair * 16, ks * 16, lane);
        mma3<false>(acc[0], a, al, bb[0], bb[1], bl[0], bl[1]);
        mma3<false>(acc[1], a, al, bb[2], bb[3], bl[2], bl[3]);
      }
      float* o = dxc + (long)b * M16 * C + pair * 16 + 2 * t;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        *reinterpret_cast<float2*>(o + (long)g * C + h * 8) = make_float2(acc[h][0], acc[h][1]);
        *reinterpret_cast<float2*>(o + (long)(g + 8) * C + h * 8) = make_float2(acc[h][2], acc[h][3]);
      }
    }
    // dW[N, 16] += dup xc^T, dbias += rowsum(dup) (a B tile of ones): accumulated in registers over the CTA's images
#pragma unroll
    for (int nt = 0; nt < MAXT; ++nt) {
      const int mt = warp + nt * NWARP;
      if (mt >= N / 16) break;
      for (int ks = 0; ks < C / 16; ++ks) {
        uint32_t a[4], al[4] = {0u, 0u, 0u, 0u}, bb[4], bl[4] = {0u, 0u, 0u, 0u};
        ldA(a, sm.X, XP, mt * 16, ks * 16, lane);
        ldB(bb, sm.D, XP, 0, ks * 16, lane);
        mma3<false>(aW[nt][0], a, al, bb[0], bb[1], bl[0], bl[1]);
        mma3<false>(aW[nt][1], a, al, bb[2], bb[3], bl[2], bl[3]);
        mma16816(aB[nt], a, ones, ones);
      }
    }
  }
#pragma unroll
  for (int nt = 0; nt < MAXT; ++nt) {
    const int mt = warp + nt * NWARP;
    if (mt >= N / 16) break;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = h * 8 + 2 * t;
      atomicAdd(dW + (mt * 16 + g) * M16 + c, aW[nt][h][0]);
      atomicAdd(dW + (mt * 16 + g) * M16 + c + 1, aW[nt][h][1]);
      atomicAdd(dW + (mt * 16 + g + 8) * M16 + c, aW[nt][h][2]);
      atomicAdd(dW + (mt * 16 + g + 8) * M16 + c + 1, aW[nt][h][3]);
    }
    if (t == 0) {
      atomicAdd(dbias + mt * 16 + g, aB[nt][0]);
      atomicAdd(dbias + mt * 16 + g + 8, aB[nt][2]);
    }
  }
}

// ============================================================================================== 32 / 48 / 64 learned tokens
// The same four products for MS = 16 m learned tokens (m <= 4) and up to 256 stream tokens -- HQAViT-TinyImageNet has
// 64 / 256 -- where the one-tile kernels above do not fit (their SIMT stand-ins were 58 % of the TinyImageNet step).  The
// slot dimension becomes a loop (K loop in the dx / up products, M loop in the xc / dxc products); the stream tile is staged
// in two channel halves where the split-precision copy (or the fp32 dW accumulator) would not fit next to it.
constexpr int MAXS = 64;              // learned tokens
constexpr int MSP = MAXS + 8;         // pitch (bf16) of the [N][MS] matrices: 144 B rows
constexpr int MAXN = 256;             // stream tokens
constexpr int NW64 = 8;               // warps per CTA: the staged tiles allow one CTA per SM, so the CTA itself must hide latency

// fp32 global rows (row stride ld, columns [c0, c0 + cols)) -> bf16 hi (/ lo) tiles [rows][pitch]
template <bool HL>
__device__ __forceinline__ void tile_to_bf16_cols(bf16* dhi, bf16* dlo, int pitch, const float* __restrict__ src, int rows, int ld, int c0,
                                                  int cols) {
  const int c4n = cols / 4, total = rows * c4n;
  const int step = blockDim.x, srow = step / c4n, scol = step - srow * c4n;
  int row = threadIdx.x / c4n, col = threadIdx.x - row * c4n;
  for (int base = 0; base < total; base += 8 * step) {
    float4 v[8];
    int o[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + u * step + threadIdx.x;
      if (i < total) v[u] = *reinterpret_cast<const float4*>(src + (long)row * ld + c0 + col * 4);
      o[u] = row * pitch + col * 4;
      row += srow; col += scol;
      if (col >= c4n) { col -= c4n; ++row; }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + u * step + threadIdx.x;
      if (i < total) {
        uint32_t h0, l0, h1, l1;
        split2(v[u].x, v[u].y, &h0, &l0);
        split2(v[u].z, v[u].w, &h1, &l1);
        *reinterpret_cast<uint2*>(dhi + o[u]) = make_uint2(h0, h1);
        if (HL) *reinterpret_cast<uint2*>(dlo + o[u]) = make_uint2(l0, l1);
      }
    }
  }
}

// ---- TokenLearner forward: xc[MS, C] = softmax_tokens(logits)^T x          (split-precision operands)
__global__ void __launch_bounds__(NW64 * 32) tlm64_fwd_kernel(const float* __restrict__ x, const bf16* __restrict__ logits, int B, int N,
                                                               int MS, int C, float* __restrict__ Sout, float* __restrict__ xc) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  const int CH = C / 2, XPc = CH + 8;
  bf16* S = reinterpret_cast<bf16*>(smraw);                    // [N][MSP] hi, lo
  bf16* SL = S + (size_t)N * MSP;
  uint8_t* U = reinterpret_cast<uint8_t*>(SL + (size_t)N * MSP);
  float* F = reinterpret_cast<float*>(U);                      // [N][MS] fp32 logits / exp (dead before X is staged)
  bf16* X = reinterpret_cast<bf16*>(U);                        // [N][XPc] hi, lo: one channel half
  bf16* XL = X + (size_t)N * XPc;
  const size_t ubytes = max((size_t)N * MS * 4, (size_t)2 * N * XPc * 2);
  float* R = reinterpret_cast<float*>(U + ((ubytes + 15) & ~(size_t)15));   // [16][16]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int c16 = tid % M16, part = tid / M16;                 // softmax: 16 threads per slot, 16 slots per pass
  constexpr int NP = NW64 * 32 / M16;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    for (int i = tid; i < N * MS / 8; i += blockDim.x) {       // 16 B loads: a scalar loop here serialised on load latency
      const uint4 q = *reinterpret_cast<const uint4*>(logits + (long)b * N * MS + i * 8);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
        F[i * 8 + 2 * j] = f.x; F[i * 8 + 2 * j + 1] = f.y;
      }
    }
    __syncthreads();
    for (int sg = 0; sg < MS / 16; ++sg) {
      const int col = sg * 16 + c16;
      float mx = -INFINITY;
      for (int n = part; n < N; n += NP) mx = fmaxf(mx, F[n * MS + col]);
      R[part * M16 + c16] = mx;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < NP; ++k) mx = fmaxf(mx, R[k * M16 + c16]);
      float z = 0.f;
      for (int n = part; n < N; n += NP) { const float e = __expf(F[n * MS + col] - mx); F[n * MS + col] = e; z += e; }
      __syncthreads();
      R[part * M16 + c16] = z;
      __syncthreads();
      z = 0.f;
#pragma unroll
      for (int k = 0; k < NP; ++k) z += R[k * M16 + c16];
      z = 1.f / z;
      for (int n = part; n < N; n += NP) {
        const float sv = F[n * MS + col] * z;
        Sout[(long)b * N * MS + n * MS + col] = sv;
        store_split<true>(S, SL, n * MSP + col, sv);
      }
      __syncthreads();
    }
    for (int ch = 0; ch < 2; ++ch) {
      __syncthreads();                                         // previous half consumed (and F dead)
      tile_to_bf16_cols<true>(X, XL, XPc, x + (long)b * N * C, N, C, ch * CH, CH);
      __syncthreads();
      const int npair = CH / 16;
      for (int item = warp; item < (MS / 16) * npair; item += NW64) {
        const int ms = item / npair, pair = item % npair;
        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        for (int ks = 0; ks < N / 16; ++ks) {
          uint32_t a[4], al[4], bb[4], bl[4];
          ldAt(a, S, MSP, ms * 16, ks * 16, lane);             // A(m = slot, k = token) = S[token][slot]
          ldAt(al, SL, MSP, ms * 16, ks * 16, lane);
          ldBt(bb, X, XPc, pair * 16, ks * 16, lane);          // B(k = token, n = channel)
          ldBt(bl, XL, XPc, pair * 16, ks * 16, lane);
          mma3<true>(acc[0], a, al, bb[0], bb[1], bl[0], bl[1]);
          mma3<true>(acc[1], a, al, bb[2], bb[3], bl[2], bl[3]);
        }
        float* o = xc + ((long)b * MS + ms * 16) * C + ch * CH + pair * 16 + 2 * t;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          *reinterpret_cast<float2*>(o + (long)g * C + h * 8) = make_float2(acc[h][0], acc[h][1]);
          *reinterpret_cast<float2*>(o + (long)(g + 8) * C + h * 8) = make_float2(acc[h][2], acc[h][3]);
        }
      }
    }
  }
}

// ---- TokenUpMix forward: up[N, C] = W[N, MS] xc[MS, C] + bias                (split-precision operands)
__global__ void __launch_bounds__(NW64 * 32) upm64_fwd_kernel(const float* __restrict__ xc, int B, int N, int MS, int C,
                                                               const float* __restrict__ W, const float* __restrict__ bias,
                                                               float* __restrict__ up) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  const int XP = C + 8;
  bf16* S = reinterpret_cast<bf16*>(smraw);                    // W hi, lo [N][MSP]
  bf16* SL = S + (size_t)N * MSP;
  bf16* D = SL + (size_t)N * MSP;                              // xc hi, lo [MS][XP]
  bf16* DL = D + (size_t)MS * XP;
  float* F = reinterpret_cast<float*>(DL + (size_t)MS * XP);   // bias [N]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  for (int i = tid; i < N * MS; i += blockDim.x) store_split<true>(S, SL, (i / MS) * MSP + i % MS, W[i]);
  for (int i = tid; i < N; i += blockDim.x) F[i] = bias[i];
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    tile_to_bf16<true>(D, DL, XP, xc + (long)b * MS * C, MS, C);
    __syncthreads();
    for (int mt = warp; mt < N / 16; mt += NW64) {
      const float b0 = F[mt * 16 + g], b1 = F[mt * 16 + g + 8];
      float* o = up + ((long)b * N + mt * 16) * C + 2 * t;
      for (int pair = 0; pair < C / 16; ++pair) {
        float acc[2][4] = {{b0, b0, b1, b1}, {b0, b0, b1, b1}};
        for (int ks = 0; ks < MS / 16; ++ks) {
          uint32_t a[4], al[4], bb[4], bl[4];
          ldA(a, S, MSP, mt * 16, ks * 16, lane);              // A(m = token, k = slot) = W
          ldA(al, SL, MSP, mt * 16, ks * 16, lane);
          ldBt(bb, D, XP, pair * 16, ks * 16, lane);           // B(k = slot, n = channel) = xc
          ldBt(bl, DL, XP, pair * 16, ks * 16, lane);
          mma3<true>(acc[0], a, al, bb[0], bb[1], bl[0], bl[1]);
          mma3<true>(acc[1], a, al, bb[2], bb[3], bl[2], bl[3]);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          *reinterpret_cast<float2*>(o + (long)g * C + pair * 16 + h * 8) = make_float2(acc[h][0], acc[h][1]);
          *reinterpret_cast<float2*>(o + (long)(g + 8) * C + pair * 16 + h * 8) = make_float2(acc[h][2], acc[h][3]);
        }
      }
    }
  }
}

// ---- TokenLearner backward: dS = x dxc^T; dlogits = S (dS - colsum(S dS)); dx = S dxc
__global__ void __launch_bounds__(NW64 * 32) tlm64_bwd_kernel(const float* __restrict__ x, const float* __restrict__ Sg,
                                                               const float* __restrict__ dxc, int B, int N, int MS, int C,
                                                               bf16* __restrict__ dlogits, float* __restrict__ dx) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  const int XP = C + 8;
  bf16* X = reinterpret_cast<bf16*>(smraw);                    // x [N][XP]
  bf16* D = X + (size_t)N * XP;                                // dxc [MS][XP]
  bf16* S = D + (size_t)MS * XP;                               // S [N][MSP]
  float* R = reinterpret_cast<float*>(S + (size_t)N * MSP);    // column sums [MS]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  constexpr int MAXT = MAXN / 16 / NW64;                       // token tiles per warp
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    for (int i = tid; i < N * MS / 4; i += blockDim.x) {       // 16 B loads
      const float4 v = *reinterpret_cast<const float4*>(Sg + (long)b * N * MS + i * 4);
      const int e = i * 4;
      *reinterpret_cast<uint2*>(S + (e / MS) * MSP + e % MS) = make_uint2(pack2(v.x, v.y), pack2(v.z, v.w));
    }
    if (tid < MS) R[tid] = 0.f;
    tile_to_bf16<false>(X, X, XP, x + (long)b * N * C, N, C);
    tile_to_bf16<false>(D, D, XP, dxc + (long)b * MS * C, MS, C);
    __syncthreads();
    for (int sg = 0; sg < MS / 16; ++sg) {
      float dS[MAXT][2][4];
#pragma unroll
      for (int nt = 0; nt < MAXT; ++nt) {
        const int mt = warp + nt * NW64;
        if (mt >= N / 16) break;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int j = 0; j < 4; ++j) dS[nt][h][j] = 0.f;
        for (int ks = 0; ks < C / 16; ++ks) {
          uint32_t a[4], bb[4];
          ldA(a, X, XP, mt * 16, ks * 16, lane);               // A(m = token, k = channel)
          ldB(bb, D, XP, sg * 16, ks * 16, lane);              // B(n = slot, k = channel)
          mma16816(dS[nt][0], a, bb[0], bb[1]);
          mma16816(dS[nt][1], a, bb[2], bb[3]);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int c = sg * 16 + h * 8 + 2 * t + j;
            float v = __bfloat162float(S[(mt * 16 + g) * MSP + c]) * dS[nt][h][j] +
                      __bfloat162float(S[(mt * 16 + g + 8) * MSP + c]) * dS[nt][h][2 + j];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (g == 0) atomicAdd(R + c, v);
          }
      }
      __syncthreads();                                         // column sums of this slot group complete
#pragma unroll
      for (int nt = 0; nt < MAXT; ++nt) {
        const int mt = warp + nt * NW64;
        if (mt >= N / 16) break;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = sg * 16 + h * 8 + 2 * t;
          const float t0 = R[c], t1 = R[c + 1];
          const int r0 = mt * 16 + g, r1 = r0 + 8;
          const uint32_t lo = pack2(__bfloat162float(S[r0 * MSP + c]) * (dS[nt][h][0] - t0),
                                    __bfloat162float(S[r0 * MSP + c + 1]) * (dS[nt][h][1] - t1));
          const uint32_t hi = pack2(__bfloat162float(S[r1 * MSP + c]) * (dS[nt][h][2] - t0),
                                    __bfloat162float(S[r1 * MSP + c + 1]) * (dS[nt][h][3] - t1));
          *reinterpret_cast<uint32_t*>(dlogits + ((long)b * N + r0) * MS + c) = lo;
          *reinterpret_cast<uint32_t*>(dlogits + ((long)b * N + r1) * MS + c) = hi;
        }
      }
    }
    // dx[N, C] = S dxc: K loop over the slots
    for (int mt = warp; mt < N / 16; mt += NW64) {
      float* o = dx + ((long)b * N + mt * 16) * C + 2 * t;
      for (int pair = 0; pair < C / 16; ++pair) {
        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        for (int ks = 0; ks < MS / 16; ++ks) {
          uint32_t a[4], bb[4];
          ldA(a, S, MSP, mt * 16, ks * 16, lane);              // A(m = token, k = slot)
          ldBt(bb, D, XP, pair * 16, ks * 16, lane);           // B(k = slot, n = channel)
          mma16816(acc[0], a, bb[0], bb[1]);
          mma16816(acc[1], a, bb[2], bb[3]);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          *reinterpret_cast<float2*>(o + (long)g * C + pair * 16 + h * 8) = make_float2(acc[h][0], acc[h][1]);
          *reinterpret_cast<float2*>(o + (long)(g + 8) * C + pair * 16 + h * 8) = make_float2(acc[h][2], acc[h][3]);
        }
      }
    }
  }
}

// ---- TokenUpMix backward: dxc = W^T dup; dW += dup xc^T; dbias += rowsum(dup)
__global__ void __launch_bounds__(NW64 * 32) upm64_bwd_kernel(const float* __restrict__ xc, const float* __restrict__ dup, int B, int N,
                                                               int MS, int C, const float* __restrict__ W, float* __restrict__ dxc,
                                                               float* __restrict__ dW, float* __restrict__ dbias) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  const int CH = C / 2, XPc = CH + 8;
  bf16* S = reinterpret_cast<bf16*>(smraw);                    // W [N][MSP]
  bf16* X = S + (size_t)N * MSP;                               // dup, one channel half [N][XPc]
  bf16* D = X + (size_t)N * XPc;                               // xc, one channel half [MS][XPc]
  float* dWs = reinterpret_cast<float*>(D + (size_t)MS * XPc); // [N][MS] fp32, accumulated over this CTA's images
  float* dbs = dWs + (size_t)N * MS;                           // [N]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  for (int i = tid; i < N * MS; i += blockDim.x) { S[(i / MS) * MSP + i % MS] = __float2bfloat16_rn(W[i]); dWs[i] = 0.f; }
  for (int i = tid; i < N; i += blockDim.x) dbs[i] = 0.f;
  const uint32_t ones = pack2(1.f, 1.f);
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    for (int ch = 0; ch < 2; ++ch) {
      __syncthreads();
      tile_to_bf16_cols<false>(X, X, XPc, dup + (long)b * N * C, N, C, ch * CH, CH);
      tile_to_bf16_cols<false>(D, D, XPc, xc + (long)b * MS * C, MS, C, ch * CH, CH);
      __syncthreads();
      // dxc[MS, half] = W^T dup: A(m = slot, k = token) = W[token][slot], B(k = token, n = channel) = dup
      const int npair = CH / 16;
      for (int item = warp; item < (MS / 16) * npair; item += NW64) {
        const int ms = item / npair, pair = item % npair;
        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        for (int ks = 0; ks < N / 16; ++ks) {
          uint32_t a[4], bb[4];
          ldAt(a, S, MSP, ms * 16, ks * 16, lane);
          ldBt(bb, X, XPc, pair * 16, ks * 16, lane);
          mma16816(acc[0], a, bb[0], bb[1]);
          mma16816(acc[1], a, bb[2], bb[3]);
        }
        float* o = dxc + ((long)b * MS + ms * 16) * C + ch * CH + pair * 16 + 2 * t;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          *reinterpret_cast<float2*>(o + (long)g * C + h * 8) = make_float2(acc[h][0], acc[h][1]);
          *reinterpret_cast<float2*>(o + (long)(g + 8) * C + h * 8) = make_float2(acc[h][2], acc[h][3]);
        }
      }
      // dW[N, MS] += dup xc^T over this half's channels (each warp owns its token tiles: plain shared-memory adds)
      for (int mt = warp; mt < N / 16; mt += NW64) {
        float aB[4] = {0.f, 0.f, 0.f, 0.f};
        for (int sg = 0; sg < MS / 16; ++sg) {
          float c[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
          for (int ks = 0; ks < CH / 16; ++ks) {
            uint32_t a[4], bb[4];
            ldA(a, X, XPc, mt * 16, ks * 16, lane);            // A(m = token, k = channel) = dup
            ldB(bb, D, XPc, sg * 16, ks * 16, lane);           // B(n = slot, k = channel) = xc
            mma16816(c[0], a, bb[0], bb[1]);
            mma16816(c[1], a, bb[2], bb[3]);
            if (sg == 0) mma16816(aB, a, ones, ones);          // row sums of dup (a B tile of ones)
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float* w0 = dWs + (mt * 16 + g) * MS + sg * 16 + h * 8 + 2 * t;
            float* w1 = dWs + (mt * 16 + g + 8) * MS + sg * 16 + h * 8 + 2 * t;
            w0[0] += c[h][0]; w0[1] += c[h][1];
            w1[0] += c[h][2]; w1[1] += c[h][3];
          }
        }
        if (t == 0) { dbs[mt * 16 + g] += aB[0]; dbs[mt * 16 + g + 8] += aB[2]; }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < N * MS; i += blockDim.x) atomicAdd(dW + i, dWs[i]);
  for (int i = tid; i < N; i += blockDim.x) atomicAdd(dbias + i, dbs[i]);
}

// ---------------------------------------------------------------------------------------------- bank write reduction
// GlobalTokenBank.write (H:303-311): per image  K_upd[16, d] = S^T c,  V_upd[16, d] = S^T tn  with S = softmax over the
// tokens of the gate logits; summed over the batch.  One product per image: [16 slots x Nt] x [Nt x 2d] on mma.sync;
// the accumulator fragments stay in registers across all images of the CTA and leave once as the CTA's partial sum.
__global__ void __launch_bounds__(NWARP * 32) bwr_mma_kernel(const bf16* __restrict__ tn, const bf16* __restrict__ cg, int ldcg, int B,
                                                             int Nt, int d, float* __restrict__ partial) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  const int XP = 2 * d + 8;
  bf16* sX = reinterpret_cast<bf16*>(smraw);                  // [Nt][XP]  [c | tn]
  bf16* sS = sX + (size_t)Nt * XP;                             // [Nt][SP]  gate, high bf16 part
  bf16* sL = sS + (size_t)Nt * SP;                             // [Nt][SP]  gate, low part: S = hi + lo to ~16 mantissa bits
  float* sF = reinterpret_cast<float*>(sL + (size_t)Nt * SP);  // [Nt][16]
  float* sR = sF + (size_t)Nt * M16;                           // [8][16]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int col = tid % M16, part = tid / M16;
  constexpr int MAXP = 8;                                      // n16 pairs per warp: 2d / 16 / 4 <= 8  (d <= 256)
  const int npairs = 2 * d / 16;
  float acc[MAXP][2][4];
#pragma unroll
  for (int q = 0; q < MAXP; ++q)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[q][0][j] = acc[q][1][j] = 0.f;
  const int v8 = d / 8;                                        // 16 B vectors per d-wide row
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const long row0 = (long)b * Nt;
    __syncthreads();
    for (int i = tid; i < Nt * M16; i += blockDim.x) sF[i] = __bfloat162float(cg[(row0 + i / M16) * ldcg + d + i % M16]);
    {   // (row, vector) advance incrementally: no per-element division by the runtime row length
      const int vpr = 2 * v8, sr = blockDim.x / vpr, sc = blockDim.x - sr * vpr;
      int r = tid / vpr, c = tid - r * vpr;
#pragma unroll 3
      for (int i = tid; i < Nt * vpr; i += blockDim.x) {
        const uint4 v = c < v8 ? *reinterpret_cast<const uint4*>(cg + (row0 + r) * ldcg + c * 8)
                               : *reinterpret_cast<const uint4*>(tn + (row0 + r) * d + (c - v8) * 8);
        *reinterpret_cast<uint4*>(sX + r * XP + c * 8) = v;
        r += sr; c += sc;
        if (c >= vpr) { c -= vpr; ++r; }
      }
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int n = part; n < Nt; n += 8) mx = fmaxf(mx, sF[n * M16 + col]);
    sR[part * M16 + col] = mx;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) mx = fmaxf(mx, sR[k * M16 + col]);
    float z = 0.f;
    for (int n = part; n < Nt; n += 8) { const float e = __expf(sF[n * M16 + col] - mx); sF[n * M16 + col] = e; z += e; }
    __syncthreads();
    sR[part * M16 + col] = z;
    __syncthreads();
    z = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) z += sR[k * M16 + col];
    z = 1.f / z;
    for (int n = part; n < Nt; n += 8) {
      const float sv = sF[n * M16 + col] * z;
      const bf16 hi = __float2bfloat16_rn(sv);
      sS[n * SP + col] = hi;
      sL[n * SP + col] = __float2bfloat16_rn(sv - __bfloat162float(hi));
    }
    __syncthreads();
    for (int ks = 0; ks < Nt / 16; ++ks) {
      uint32_t a[4], al[4];
      ldAt(a, sS, SP, 0, ks * 16, lane);
      ldAt(al, sL, SP, 0, ks * 16, lane);
#pragma unroll
      for (int q = 0; q < MAXP; ++q) {
        const int pair = warp + q * NWARP;
        if (pair >= npairs) break;
        uint32_t bb[4];
        ldBt(bb, sX, XP, pair * 16, ks * 16, lane);
        mma16816(acc[q][0], a, bb[0], bb[1]);
        mma16816(acc[q][1], a, bb[2], bb[3]);
        mma16816(acc[q][0], al, bb[0], bb[1]);
        mma16816(acc[q][1], al, bb[2], bb[3]);
      }
    }
  }
  float* pk = partial + (long)blockIdx.x * 2 * M16 * d;
#pragma unroll
  for (int q = 0; q < MAXP; ++q) {
    const int pair = warp + q * NWARP;
    if (pair >= npairs) break;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int cc = pair * 16 + h * 8 + 2 * t;              // column in [c | tn]
      const int role = cc >= d ? 1 : 0, ch = cc - role * d;
      *reinterpret_cast<float2*>(pk + ((long)role * M16 + g) * d + ch) = make_float2(acc[q][h][0], acc[q][h][1]);
      *reinterpret_cast<float2*>(pk + ((long)role * M16 + g + 8) * d + ch) = make_float2(acc[q][h][2], acc[q][h][3]);
    }
  }
}

// Same reduction for short images (Nt <= 64; HQAViT's 16 learned tokens, QAViTv2's 64): the sum over the batch lets G images be STACKED along the
// contraction (token) axis -- [16 slots x G Nt] x [G Nt x 2d] -- so one CTA iteration (one load phase, one softmax phase, three
// barriers) covers G images instead of one; the per-(image, slot) softmax over <= 32 tokens is a serial loop of one thread.
// (The one-image-per-iteration kernel above spent its time in 7 barriers per 16-row image: 55 us for 60 MB, ncu.)
constexpr int GWARP = 8;
__global__ void __launch_bounds__(GWARP * 32) bwr_mma_grp_kernel(const bf16* __restrict__ tn, const bf16* __restrict__ cg, int ldcg, int B,
                                                                 int Nt, int G, int d, float* __restrict__ partial) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  const int XP = 2 * d + 8, R = G * Nt;
  bf16* sX = reinterpret_cast<bf16*>(smraw);                  // [R][XP]  [c | tn]
  bf16* sS = sX + (size_t)R * XP;                              // [R][SP]  gate, high bf16 part
  bf16* sL = sS + (size_t)R * SP;                              // [R][SP]  gate, low part
  float* sF = reinterpret_cast<float*>(sL + (size_t)R * SP);   // [R][16]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  constexpr int MAXP = 4;                                      // n16 pairs per warp: 2d / 16 / 8 <= 4  (d <= 256)
  const int npairs = 2 * d / 16;
  float acc[MAXP][2][4];
#pragma unroll
  for (int q = 0; q < MAXP; ++q)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[q][0][j] = acc[q][1][j] = 0.f;
  const int v8 = d / 8, vpr = 2 * v8;
  const int ngroups = (B + G - 1) / G;
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int b0 = grp * G, nimg = min(G, B - b0), rows = nimg * Nt;
    const long row0 = (long)b0 * Nt;
    __syncthreads();
    for (int i = tid; i < rows * M16; i += blockDim.x) sF[i] = __bfloat162float(cg[(row0 + i / M16) * ldcg + d + i % M16]);
    {
      const int sr = blockDim.x / vpr, sc = blockDim.x - sr * vpr;
      int r = tid / vpr, c = tid - r * vpr;
#pragma unroll 4
      for (int i = tid; i < rows * vpr; i += blockDim.x) {
        const uint4 v = c < v8 ? *reinterpret_cast<const uint4*>(cg + (row0 + r) * ldcg + c * 8)
                               : *reinterpret_cast<const uint4*>(tn + (row0 + r) * d + (c - v8) * 8);
        *reinterpret_cast<uint4*>(sX + r * XP + c * 8) = v;
        r += sr; c += sc;
        if (c >= vpr) { c -= vpr; ++r; }
      }
    }
    __syncthreads();
    // softmax over the Nt tokens of (image, slot)
    if (Nt <= 32) {                                              // short images: one thread per pair (64 pairs per pass)
      for (int pq = tid; pq < nimg * M16; pq += blockDim.x) {
        const int col = pq % M16;
        float* f = sF + (pq / M16) * Nt * M16 + col;
        float mx = -INFINITY;
        for (int n = 0; n < Nt; ++n) mx = fmaxf(mx, f[n * M16]);
        float z = 0.f;
        for (int n = 0; n < Nt; ++n) { const float e = __expf(f[n * M16] - mx); f[n * M16] = e; z += e; }
        z = 1.f / z;
        const int r0 = (pq / M16) * Nt;
        for (int n = 0; n < Nt; ++n) {
          const float sv = f[n * M16] * z;
          const bf16 hi = __float2bfloat16_rn(sv);
          sS[(r0 + n) * SP + col] = hi;
          sL[(r0 + n) * SP + col] = __float2bfloat16_rn(sv - __bfloat162float(hi));
        }
      }
    } else {
      // 16 lanes per pair, tokens strided over them, shuffle reductions inside the half-warp (one thread per pair walking 64 tokens
      // three times was the longest phase of a pass; at 16 tokens the serial loop above is the faster one: 36 vs 47 us per launch)
      for (int pq0 = 0; pq0 < nimg * M16; pq0 += blockDim.x / 16) {
        const int pq = pq0 + tid / 16, part = tid & 15;
        const bool on = pq < nimg * M16;
        const int col = on ? pq % M16 : 0, r0 = on ? (pq / M16) * Nt : 0;
        float* f = sF + r0 * M16 + col;
        float mx = -INFINITY;
        if (on) for (int n = part; n < Nt; n += 16) mx = fmaxf(mx, f[n * M16]);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float z = 0.f;
        if (on) for (int n = part; n < Nt; n += 16) { const float e = __expf(f[n * M16] - mx); f[n * M16] = e; z += e; }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
        z = 1.f / z;
        if (on) for (int n = part; n < Nt; n += 16) {
          const float sv = f[n * M16] * z;
          const bf16 hi = __float2bfloat16_rn(sv);
          sS[(r0 + n) * SP + col] = hi;
          sL[(r0 + n) * SP + col] = __float2bfloat16_rn(sv - __bfloat162float(hi));
        }
      }
    }
    __syncthreads();
    for (int ks = 0; ks < rows / 16; ++ks) {
      uint32_t a[4], al[4];
      ldAt(a, sS, SP, 0, ks * 16, lane);
      ldAt(al, sL, SP, 0, ks * 16, lane);
#pragma unroll
      for (int q = 0; q < MAXP; ++q) {
        const int pair = warp + q * GWARP;
        if (pair >= npairs) break;
        uint32_t bb[4];
        ldBt(bb, sX, XP, pair * 16, ks * 16, lane);
        mma16816(acc[q][0], a, bb[0], bb[1]);
        mma16816(acc[q][1], a, bb[2], bb[3]);
        mma16816(acc[q][0], al, bb[0], bb[1]);
        mma16816(acc[q][1], al, bb[2], bb[3]);
      }
    }
  }
  float* pk = partial + (long)blockIdx.x * 2 * M16 * d;
#pragma unroll
  for (int q = 0; q < MAXP; ++q) {
    const int pair = warp + q * GWARP;
    if (pair >= npairs) break;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int cc = pair * 16 + h * 8 + 2 * t;              // column in [c | tn]
      const int role = cc >= d ? 1 : 0, ch = cc - role * d;
      *reinterpret_cast<float2*>(pk + ((long)role * M16 + g) * d + ch) = make_float2(acc[q][h][0], acc[q][h][1]);
      *reinterpret_cast<float2*>(pk + ((long)role * M16 + g + 8) * d + ch) = make_float2(acc[q][h][2], acc[q][h][3]);
    }
  }
}

template <typename K>
int opt_in(K kernel, size_t bytes) {
  QV_CHECK(bytes <= 200 * 1024, "token kernels need %zu B of shared memory", bytes);
  if (bytes > 48 * 1024) QV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}
int tok_grid(int B, size_t smem) {
  const int occ = max(1, min(8, (int)(200 * 1024 / (smem + 1024))));
  return max(1, min(B, qv_num_sms() * occ));
}

}  // namespace

bool tokens_mma_ok(int M, int N, int C) { return M == 16 && N % 16 == 0 && N >= 16 && N <= 128 && C % 16 == 0 && C <= 512; }
// 32 / 48 / 64 learned tokens, up to 256 stream tokens (tlm64_* / upm64_* kernels)
bool tokens_mma64_ok(int M, int N, int C) {
  return M % 16 == 0 && M >= 16 && M <= MAXS && N % 16 == 0 && N >= 16 && N <= MAXN && C % 32 == 0 && C <= 256 && !tokens_mma_ok(M, N, C);
}
int tlm64_fwd(cudaStream_t s, const float* x, const void* logits, int B, int N, int M, int C, float* S, float* xc) {
  const int XPc = C / 2 + 8;
  const size_t ub = max((size_t)N * M * 4, (size_t)2 * N * XPc * 2);
  const size_t smem = (size_t)2 * N * MSP * 2 + ((ub + 15) & ~(size_t)15) + 16 * M16 * 4 + 16;
  QV_TRY(opt_in(tlm64_fwd_kernel, smem));
  qv_launch(tlm64_fwd_kernel, tok_grid(B, smem), NW64 * 32, smem, s, x, (const bf16*)logits, B, N, M, C, S, xc);
  QV_LAUNCH_CHECK();
  return 0;
}
int tlm64_bwd(cudaStream_t s, const float* x, const float* S, const float* dxc, int B, int N, int M, int C, void* dlogits, float* dx) {
  const size_t smem = ((size_t)N * (C + 8) + (size_t)M * (C + 8) + (size_t)N * MSP) * 2 + (size_t)M * 4 + 16;
  QV_TRY(opt_in(tlm64_bwd_kernel, smem));
  qv_launch(tlm64_bwd_kernel, tok_grid(B, smem), NW64 * 32, smem, s, x, S, dxc, B, N, M, C, (bf16*)dlogits, dx);
  QV_LAUNCH_CHECK();
  return 0;
}
int upm64_fwd(cudaStream_t s, const float* xc, int B, int N, int M, int C, const float* W, const float* bias, float* up) {
  const size_t smem = ((size_t)2 * N * MSP + (size_t)2 * M * (C + 8)) * 2 + (size_t)N * 4 + 16;
  QV_TRY(opt_in(upm64_fwd_kernel, smem));
  qv_launch(upm64_fwd_kernel, tok_grid(B, smem), NW64 * 32, smem, s, xc, B, N, M, C, W, bias, up);
  QV_LAUNCH_CHECK();
  return 0;
}
int upm64_bwd(cudaStream_t s, const float* xc, const float* dup, int B, int N, int M, int C, const float* W, float* dxc, float* dW,
              float* dbias) {
  const int XPc = C / 2 + 8;
  const size_t smem = ((size_t)N * MSP + (size_t)N * XPc + (size_t)M * XPc) * 2 + ((size_t)N * M + N) * 4 + 16;
  QV_TRY(opt_in(upm64_bwd_kernel, smem));
  qv_launch(upm64_bwd_kernel, tok_grid(B, smem), NW64 * 32, smem, s, xc, dup, B, N, M, C, W, dxc, dW, dbias);
  QV_LAUNCH_CHECK();
  return 0;
}

int tlm_fwd(cudaStream_t s, const float* x, const void* logits, int B, int N, int C, float* S, float* xc) {
  const size_t smem = smem_bytes(N, C, true, true, false);
  QV_TRY(opt_in(tlm_fwd_kernel, smem));
  qv_launch(tlm_fwd_kernel, tok_grid(B, smem), NWARP * 32, smem, s, x, (const bf16*)logits, B, N, C, S, xc);
  QV_LAUNCH_CHECK();
  return 0;
}
int tlm_bwd(cudaStream_t s, const float* x, const float* S, const float* dxc, int B, int N, int C, void* dlogits, float* dx) {
  const size_t smem = smem_bytes(N, C, false);
  QV_TRY(opt_in(tlm_bwd_kernel, smem));
  qv_launch(tlm_bwd_kernel, tok_grid(B, smem), NWARP * 32, smem, s, x, S, dxc, B, N, C, (bf16*)dlogits, dx);
  QV_LAUNCH_CHECK();
  return 0;
}
int upm_fwd(cudaStream_t s, const float* xc, int B, int N, int C, const float* W, const float* bias, float* up) {
  const size_t smem = smem_bytes(N, C, true, false, true);
  QV_TRY(opt_in(upm_fwd_kernel, smem));
  qv_launch(upm_fwd_kernel, tok_grid(B, smem), NWARP * 32, smem, s, xc, B, N, C, W, bias, up);
  QV_LAUNCH_CHECK();
  return 0;
}
int upm_bwd(cudaStream_t s, const float* xc, const float* dup, int B, int N, int C, const float* W, float* dxc, float* dW,
            float* dbias) {
  const size_t smem = smem_bytes(N, C, false);
  QV_TRY(opt_in(upm_bwd_kernel, smem));
  // the dW / dbias accumulators are flushed with atomics once per CTA: keep the CTA count moderate
  const int grid = max(1, min(B, qv_num_sms() * 3));
  qv_launch(upm_bwd_kernel, grid, NWARP * 32, smem, s, xc, dup, B, N, C, W, dxc, dW, dbias);
  QV_LAUNCH_CHECK();
  return 0;
}

bool bank_write_mma_ok(int Nt, int d, int kb, int ldcg) {
  return kb == 16 && Nt % 16 == 0 && Nt >= 16 && Nt <= 256 && d % 16 == 0 && d <= 256 && ldcg % 8 == 0;
}
int bank_write_reduce_mma(cudaStream_t s, const void* tn, const void* cg, int ldcg, int B, int Nt, int d, float* partial, int* n_partial) {
  if (Nt <= 64) {   // stacked-image flavour (one 64-token image per pass at Nt = 64)
    const int G = 64 / Nt;
    const int R = G * Nt;
    const size_t smem = ((size_t)R * (2 * d + 8) + (size_t)2 * R * SP) * 2 + (size_t)R * M16 * 4 + 16;
    QV_TRY(opt_in(bwr_mma_grp_kernel, smem));
    const int occ = max(1, min(4, (int)(200 * 1024 / (smem + 1024))));
    const int grid = max(1, min(cdiv(B, G), min(592, qv_num_sms() * occ)));
    *n_partial = grid;
    qv_launch(bwr_mma_grp_kernel, grid, GWARP * 32, smem, s, (const bf16*)tn, (const bf16*)cg, ldcg, B, Nt, G, d, partial);
    QV_LAUNCH_CHECK();
    return 0;
  }
  const size_t smem = ((size_t)Nt * (2 * d + 8) + (size_t)2 * Nt * SP) * 2 + ((size_t)Nt * M16 + 8 * M16) * 4 + 16;
  QV_TRY(opt_in(bwr_mma_kernel, smem));
  const int occ = max(1, min(4, (int)(200 * 1024 / (smem + 1024))));
  const int grid = max(1, min(cdiv(B, 2), min(592, qv_num_sms() * occ)));
  *n_partial = grid;
  qv_launch(bwr_mma_kernel, grid, NWARP * 32, smem, s, (const bf16*)tn, (const bf16*)cg, ldcg, B, Nt, d, partial);
  QV_LAUNCH_CHECK();
  return 0;
}
