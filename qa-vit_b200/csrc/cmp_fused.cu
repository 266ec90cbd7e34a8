// Per-branch LayerNorm -> compress Linear(192 -> 48) -> fusion scale -> concat (H:1074-1079) and its backward as ONE launch each
// for all four branches of a block (bf16 runs, d = 192, compress_dim = 48):
//
//   cmpf_fwd : fused[:, 48 i : 48 i + 48] = alpha_i (LN_i(branch_i) W_i^T + b_i),  i = blockIdx.y          (4 x (ln_fwd + GEMM) before)
//   cmpf_bwd : d_branch_i = dropout-mask . LN_i-backward(alpha_i d_fused_i W_i), dW_i, db_i, dgamma_i, dbeta_i   (4 x (2 GEMMs + ln_bwd))
//
// The LayerNorm is folded into the GEMM algebraically (as in tokens_fused.cu): with Wg = gamma (.) W, c1 = rowsum(Wg),
// c0 = W beta + b:  LN(x) W^T + b = rstd (x Wg^T - mean c1) + c0, so the bf16 branch tile is the MMA operand as it is (exact) and the
// normalised tensor is never materialised -- forward no longer writes the four [R, 192] LayerNorm outputs, backward no longer
// reads them.  Backward uses the same identity for the weight gradient: with dz = alpha d_fused_i, dl' = dz rstd,
// P = dl'^T x, U = dl'^T mean, T = colsum(dz):  dW = gamma (P - U) + beta T, dgamma = colsum_j W (P - U), dbeta = colsum_j W T,
// db = T; U and T come out of the same MMA as P through two extra columns (mean, 1 / rstd) appended to the x tile.
// mma.sync.m16n8k16 bf16 with fp32 accumulation; LayerNorm statistics and all elementwise math in fp32.  HBM-bound: forward reads
// the branch once and writes a 48-column slice, backward reads the branch and the gradient slice and writes d_branch.
#include "kernels.h"

namespace {

constexpr int KC = 192;          // channels
constexpr int KD = 48;           // compress dim (one branch's slice of the fused tensor)
constexpr int KXP = 216;         // pitch (bf16) of the [rows][C + 8] tile: 432 B = 27 x 16 B (odd: conflict-free ldmatrix)
constexpr int KWP = 200;         // pitch (bf16) of [48][C] weight tiles
constexpr int KHP = 56;          // pitch (bf16) of [rows][48] tiles: 112 B = 7 x 16 B
constexpr int KTR = 64;          // rows per tile
constexpr int KNT = 256, KNW = 8;

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
__device__ __forceinline__ uint32_t pack2(float x, float y) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void store_split(bf16* hi, bf16* lo, int idx, float v) {
  const bf16 h = __float2bfloat16_rn(v);
  hi[idx] = h;
  lo[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
}
__device__ __forceinline__ void ldA(uint32_t* a, const bf16* base, int pitch, int m0, int k0, int lane) {        // [m][k]
  const int mat = lane >> 3, r = lane & 7;
  ldsm4(a, sa(base + (m0 + r + (mat & 1) * 8) * pitch + k0 + (mat >> 1) * 8));
}
__device__ __forceinline__ void ldAt(uint32_t* a, const bf16* base, int pitch, int m0, int k0, int lane) {       // [k][m]
  const int mat = lane >> 3, r = lane & 7;
  ldsm4t(a, sa(base + (k0 + r + (mat >> 1) * 8) * pitch + m0 + (mat & 1) * 8));
}
__device__ __forceinline__ void ldB2(uint32_t* b, const bf16* base, int pitch, int n0, int k0, int lane) {       // [n][k], one n8
  const int mat = (lane >> 3) & 1, r = lane & 7;
  ldsm2(b, sa(base + (n0 + r) * pitch + k0 + mat * 8));
}
__device__ __forceinline__ void ldBt(uint32_t* b, const bf16* base, int pitch, int n0, int k0, int lane) {       // [k][n], two n8
  const int mat = lane >> 3, r = lane & 7;
  ldsm4t(b, sa(base + (k0 + r + (mat & 1) * 8) * pitch + n0 + (mat >> 1) * 8));
}
__device__ __forceinline__ void ldBt2(uint32_t* b, const bf16* base, int pitch, int n0, int k0, int lane) {      // [k][n], one n8
  const int mat = (lane >> 3) & 1, r = lane & 7;
  ldsm2t(b, sa(base + (k0 + r + mat * 8) * pitch + n0));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
constexpr size_t al16(size_t b) { return (b + 15) & ~(size_t)15; }
struct Carve {
  uint8_t* p;
  __device__ explicit Carve(uint8_t* base) : p(base) {}
  template <typename T> __device__ T* take(size_t n) {
    T* r = reinterpret_cast<T*>(p);
    p += (n * sizeof(T) + 15) & ~(size_t)15;
    return r;
  }
};

// a [64, 192] bf16 tile in registers: 1536 16 B chunks over 256 threads
struct TRegs { uint4 v[6]; };
__device__ __forceinline__ void tload(TRegs& r, const bf16* __restrict__ x, long row0, long R, int tid) {
#pragma unroll
  for (int u = 0; u < 6; ++u) {
    const int i = tid + KNT * u, row = i / 24, ch = i - row * 24;
    r.v[u] = (row0 + row < R) ? __ldg(reinterpret_cast<const uint4*>(x + (row0 + row) * KC) + ch) : make_uint4(0u, 0u, 0u, 0u);
  }
}
__device__ __forceinline__ void tstore(const TRegs& r, bf16* X, int tid) {
#pragma unroll
  for (int u = 0; u < 6; ++u) {
    const int i = tid + KNT * u, row = i / 24, ch = i - row * 24;
    *reinterpret_cast<uint4*>(X + row * KXP + ch * 8) = r.v[u];
  }
}
// mean / rstd of the tile's rows from shared memory: warp w owns rows w, w + 8, ...; a lane reads 6 contiguous channels
__device__ __forceinline__ void tile_stats(const bf16* X, float* mean_s, float* rstd_s, int warp, int lane, float eps) {
  for (int r = warp; r < KTR; r += KNW) {
    const bf16* p = X + r * KXP + lane * 6;
    float v[6];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p + 2 * i));
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) s += v[i];
    const float m = warp_sum(s) * (1.f / KC);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) { const float d = v[i] - m; q += d * d; }
    const float rs = rsqrtf(warp_sum(q) * (1.f / KC) + eps);
    if (lane == 0) { mean_s[r] = m; rstd_s[r] = rs; }
  }
}

struct CmpFwdArgs {
  const bf16* x[4];
  const float* gamma[4]; const float* beta[4]; const float* W[4]; const float* bias[4];
  float* stats[4];
  const float* alpha;
  bf16* out;
  long R;
  float eps;
};
constexpr size_t CMPF_FWD_SMEM = al16(KTR * KXP * 2) + 2 * al16(KD * KWP * 2) + 2 * al16(KTR * 4) + 2 * al16(KD * 4);

__global__ void __launch_bounds__(KNT, 2) cmpf_fwd_kernel(CmpFwdArgs a) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  Carve cv(smraw);
  bf16* X = cv.take<bf16>(KTR * KXP);
  bf16* Wg = cv.take<bf16>(KD * KWP); bf16* WgL = cv.take<bf16>(KD * KWP);
  float* mean_s = cv.take<float>(KTR); float* rstd_s = cv.take<float>(KTR);
  float* c0 = cv.take<float>(KD); float* c1 = cv.take<float>(KD);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int br = blockIdx.y;
  const bf16* __restrict__ x = a.x[br];
  const float* __restrict__ W = a.W[br];
  const float* __restrict__ gamma = a.gamma[br];
  const float* __restrict__ beta = a.beta[br];
  const float al = a.alpha[br];
  for (int i = tid; i < KD * KC; i += KNT) {
    const int j = i / KC, k = i - j * KC;
    store_split(Wg, WgL, j * KWP + k, W[i] * gamma[k]);
  }
  for (int j = warp; j < KD; j += KNW) {
    float a1 = 0.f, a0 = 0.f;
    for (int k = lane; k < KC; k += 32) { const float w = W[j * KC + k]; a1 = fmaf(w, gamma[k], a1); a0 = fmaf(w, beta[k], a0); }
    a1 = warp_sum(a1); a0 = warp_sum(a0);
    if (lane == 0) { c1[j] = a1; c0[j] = a0 + a.bias[br][j]; }
  }
  const long ntiles = (a.R + KTR - 1) / KTR;
  TRegs xr;
  if ((long)blockIdx.x < ntiles) tload(xr, x, (long)blockIdx.x * KTR, a.R, tid);
  const int mt = warp & 3, ng = warp >> 2;
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long row0 = tile * KTR;
    tstore(xr, X, tid);
    if (tile + gridDim.x < ntiles) tload(xr, x, (tile + gridDim.x) * KTR, a.R, tid);
    __syncthreads();
    tile_stats(X, mean_s, rstd_s, warp, lane, a.eps);
    __syncthreads();
    if (tid < KTR && row0 + tid < a.R) *reinterpret_cast<float2*>(a.stats[br] + (row0 + tid) * 2) = make_float2(mean_s[tid], rstd_s[tid]);
    float acc[3][4];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int ks = 0; ks < KC / 16; ++ks) {
      uint32_t af[4];
      ldA(af, X, KXP, mt * 16, ks * 16, lane);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        uint32_t bh[2], bl[2];
        ldB2(bh, Wg, KWP, (ng * 3 + i) * 8, ks * 16, lane);
        ldB2(bl, WgL, KWP, (ng * 3 + i) * 8, ks * 16, lane);
        mma16816(acc[i], af, bh[0], bh[1]);
        mma16816(acc[i], af, bl[0], bl[1]);
      }
    }
    const int n0 = mt * 16 + g, n1 = n0 + 8;
    const float r0 = rstd_s[n0], u0 = mean_s[n0], r1 = rstd_s[n1], u1 = mean_s[n1];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int j = (ng * 3 + i) * 8 + 2 * t;
      const float v00 = al * fmaf(r0, acc[i][0] - u0 * c1[j], c0[j]), v01 = al * fmaf(r0, acc[i][1] - u0 * c1[j + 1], c0[j + 1]);
      const float v10 = al * fmaf(r1, acc[i][2] - u1 * c1[j], c0[j]), v11 = al * fmaf(r1, acc[i][3] - u1 * c1[j + 1], c0[j + 1]);
      if (row0 + n0 < a.R) *reinterpret_cast<uint32_t*>(a.out + (row0 + n0) * KC + br * KD + j) = pack2(v00, v01);
      if (row0 + n1 < a.R) *reinterpret_cast<uint32_t*>(a.out + (row0 + n1) * KC + br * KD + j) = pack2(v10, v11);
    }
    __syncthreads();
  }
}

struct CmpBwdArgs {
  const bf16* x[4];            // branch outputs [R, 192]
  const float* stats[4];       // (mean, rstd) per row
  const float* gamma[4]; const float* beta[4]; const float* W[4];
  float* dW[4]; float* db[4]; float* dgamma[4]; float* dbeta[4];
  bf16* dx[4];                 // d_branch [R, 192]
  const bf16* dfused;          // [R, 192]: branch i owns columns [48 i, 48 i + 48)
  const float* alpha;
  long R;
  DropP drop[4];               // branch-output dropout site (mask applied to d_branch; ids of drop_rows on [R, 192])
};
constexpr size_t CMPF_BWD_SMEM = 2 * al16(KTR * KXP * 2) + al16(KD * KWP * 2) + 2 * al16(KTR * KHP * 2) + 2 * al16(KTR * 4) +
                                 al16(2 * KTR * 2 * 4) + 2 * al16(KC * 4);

__global__ void __launch_bounds__(KNT, 1) cmpf_bwd_kernel(CmpBwdArgs a) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  Carve cv(smraw);
  bf16* X = cv.take<bf16>(KTR * KXP);          // x tile + 8 extra columns: (mean, 1/rstd) as hi | lo pairs, zeros
  bf16* O = cv.take<bf16>(KTR * KXP);          // d_branch staging (pitch shared with X)
  bf16* Ws = cv.take<bf16>(KD * KWP);          // alpha W   [j][k]
  bf16* Gd = cv.take<bf16>(KTR * KHP);         // d_fused slice  [row][j]
  bf16* Hd = cv.take<bf16>(KTR * KHP);         // d_fused slice * rstd
  float* mean_s = cv.take<float>(KTR); float* rstd_s = cv.take<float>(KTR);
  float* RS = cv.take<float>(2 * KTR * 2);
  float* gam = cv.take<float>(KC); float* bet = cv.take<float>(KC);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int br = blockIdx.y;
  const bf16* __restrict__ x = a.x[br];
  const float* __restrict__ stats = a.stats[br];
  const float* __restrict__ W = a.W[br];
  const float al = a.alpha[br];
  const bool masked = a.drop[br].p > 0.f;
  DropState dst{};
  if (masked) dst = drop_state(a.drop[br]);
  for (int i = tid; i < KD * KC; i += KNT) {
    const int j = i / KC, k = i - j * KC;
    Ws[j * KWP + k] = __float2bfloat16_rn(W[i] * al);
  }
  for (int k = tid; k < KC; k += KNT) { gam[k] = a.gamma[br][k]; bet[k] = a.beta[br][k]; }
  // persistent accumulators of P^T-ish products: warp owns (3 m-tiles of j) x (n8 tiles nt = warp * 3 + i, plus the stats tile for warp 0)
  float Pacc[3][3][4];
  float Sacc[3][4];            // extra tile (columns 192..199) -- only warp 0 accumulates it
#pragma unroll
  for (int m = 0; m < 3; ++m) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) Pacc[m][i][j] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) Sacc[m][j] = 0.f;
  }
  const long ntiles = (a.R + KTR - 1) / KTR;
  TRegs xr;
  uint4 gr[2];                 // d_fused slice: 64 rows x 6 chunks of 16 B = 384 chunks over 256 threads
  float2 sr = make_float2(0.f, 1.f);
  auto prefetch = [&](long tile) {
    const long row0 = tile * KTR;
    tload(xr, x, row0, a.R, tid);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = tid + KNT * u, row = i / 6, ch = i - row * 6;
      gr[u] = (i < KTR * 6 && row0 + row < a.R) ? __ldg(reinterpret_cast<const uint4*>(a.dfused + (row0 + row) * KC + br * KD) + ch)
                                               : make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid < KTR) sr = (row0 + tid < a.R) ? __ldg(reinterpret_cast<const float2*>(stats + (row0 + tid) * 2)) : make_float2(0.f, 0.f);
  };
  if ((long)blockIdx.x < ntiles) prefetch(blockIdx.x);
  const int mt = warp & 3, hf = warp >> 2;
  const int n0 = mt * 16 + g, n1 = n0 + 8;
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long row0 = tile * KTR;
    tstore(xr, X, tid);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = tid + KNT * u, row = i / 6, ch = i - row * 6;
      if (i < KTR * 6) *reinterpret_cast<uint4*>(Gd + row * KHP + ch * 8) = gr[u];
    }
    if (tid < KTR) {
      // rows past R carry rstd = 0: their (zero) gradient rows then contribute nothing anywhere
      mean_s[tid] = sr.x; rstd_s[tid] = sr.y;
      const float sd = sr.y > 0.f ? 1.f / sr.y : 0.f;
      const bf16 mh = __float2bfloat16_rn(sr.x), sh = __float2bfloat16_rn(sd);
      bf16* e = X + tid * KXP + KC;
      e[0] = mh; e[1] = sh;
      e[2] = __float2bfloat16_rn(sr.x - __bfloat162float(mh)); e[3] = __float2bfloat16_rn(sd - __bfloat162float(sh));
      e[4] = e[5] = e[6] = e[7] = __float2bfloat16_rn(0.f);
    }
    if (tile + gridDim.x < ntiles) prefetch(tile + gridDim.x);
    __syncthreads();
    // ---- dl' = d_fused slice * rstd (bf16): thread = (row, 12 contiguous columns)
    {
      const int row = tid >> 2, c = (tid & 3) * 12;
      const float rs = rstd_s[row];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(Gd + row * KHP + c + 2 * i));
        *reinterpret_cast<uint32_t*>(Hd + row * KHP + c + 2 * i) = pack2(f.x * rs, f.y * rs);
      }
    }
    // ---- d_ln = d_fused (alpha W) (K = 48), row sums for the LayerNorm backward over this warp's 96 channels
    uint32_t ag[3][4];
#pragma unroll
    for (int ks = 0; ks < 3; ++ks) ldA(ag[ks], Gd, KHP, mt * 16, ks * 16, lane);
    const float r0 = rstd_s[n0], r1 = rstd_s[n1], u0 = mean_s[n0], u1 = mean_s[n1];
    float s1a = 0.f, s2a = 0.f, s1b = 0.f, s2b = 0.f;
#pragma unroll
    for (int pr = 0; pr < 6; ++pr) {
      const int c0 = hf * 96 + pr * 16;
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int ks = 0; ks < 3; ++ks) {
        uint32_t bw[4];
        ldBt(bw, Ws, KWP, c0, ks * 16, lane);
        mma16816(acc[0], ag[ks], bw[0], bw[1]);
        mma16816(acc[1], ag[ks], bw[2], bw[3]);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = c0 + h * 8 + 2 * t;
        const float2 xa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(X + n0 * KXP + c));
        const float2 xb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(X + n1 * KXP + c));
        const float g0 = gam[c], g1 = gam[c + 1];
        const float va0 = acc[h][0] * g0, va1 = acc[h][1] * g1, vb0 = acc[h][2] * g0, vb1 = acc[h][3] * g1;
        s1a += va0 + va1; s2a += va0 * ((xa.x - u0) * r0) + va1 * ((xa.y - u0) * r0);
        s1b += vb0 + vb1; s2b += vb0 * ((xb.x - u1) * r1) + vb1 * ((xb.y - u1) * r1);
      }
    }
    s1a = quad_sum(s1a); s2a = quad_sum(s2a); s1b = quad_sum(s1b); s2b = quad_sum(s2b);
    if (t == 0) {
      *reinterpret_cast<float2*>(RS + (hf * KTR + n0) * 2) = make_float2(s1a, s2a);
      *reinterpret_cast<float2*>(RS + (hf * KTR + n1) * 2) = make_float2(s1b, s2b);
    }
    __syncthreads();
    // ---- dx = rstd (g - mean(g) - xhat mean(g xhat)) -> staging tile (bf16)
    {
      const float2 pa0 = *reinterpret_cast<const float2*>(RS + n0 * 2), pa1 = *reinterpret_cast<const float2*>(RS + (KTR + n0) * 2);
      const float2 pb0 = *reinterpret_cast<const float2*>(RS + n1 * 2), pb1 = *reinterpret_cast<const float2*>(RS + (KTR + n1) * 2);
      const float m1a = (pa0.x + pa1.x) * (1.f / KC), m2a = (pa0.y + pa1.y) * (1.f / KC);
      const float m1b = (pb0.x + pb1.x) * (1.f / KC), m2b = (pb0.y + pb1.y) * (1.f / KC);
#pragma unroll
      for (int pr = 0; pr < 6; ++pr) {
        const int c0 = hf * 96 + pr * 16;
        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) {
          uint32_t bw[4];
          ldBt(bw, Ws, KWP, c0, ks * 16, lane);
          mma16816(acc[0], ag[ks], bw[0], bw[1]);
          mma16816(acc[1], ag[ks], bw[2], bw[3]);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = c0 + h * 8 + 2 * t;
          const float2 xa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(X + n0 * KXP + c));
          const float2 xb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(X + n1 * KXP + c));
          const float g0 = gam[c], g1 = gam[c + 1];
          const float ya0 = r0 * (acc[h][0] * g0 - m1a - (xa.x - u0) * r0 * m2a), ya1 = r0 * (acc[h][1] * g1 - m1a - (xa.y - u0) * r0 * m2a);
          const float yb0 = r1 * (acc[h][2] * g0 - m1b - (xb.x - u1) * r1 * m2b), yb1 = r1 * (acc[h][3] * g1 - m1b - (xb.y - u1) * r1 * m2b);
          *reinterpret_cast<uint32_t*>(O + n0 * KXP + c) = pack2(ya0, ya1);
          *reinterpret_cast<uint32_t*>(O + n1 * KXP + c) = pack2(yb0, yb1);
        }
      }
    }
    // ---- P[48, C (+ 8)] += dl'^T [x | mean, 1/rstd]: warp = n8 tiles warp * 3 + i (and tile 24 for warp 0), all 3 m-tiles of j
    for (int ks = 0; ks < KTR / 16; ++ks) {
      uint32_t ah[3][4];
#pragma unroll
      for (int m = 0; m < 3; ++m) ldAt(ah[m], Hd, KHP, m * 16, ks * 16, lane);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        uint32_t bx[2];
        ldBt2(bx, X, KXP, (warp * 3 + i) * 8, ks * 16, lane);
#pragma unroll
        for (int m = 0; m < 3; ++m) mma16816(Pacc[m][i], ah[m], bx[0], bx[1]);
      }
      if (warp == 0) {
        uint32_t bx[2];
        ldBt2(bx, X, KXP, KC, ks * 16, lane);
#pragma unroll
        for (int m = 0; m < 3; ++m) mma16816(Sacc[m], ah[m], bx[0], bx[1]);
      }
    }
    __syncthreads();
    // ---- staging tile -> d_branch: a thread moves 16 B chunks (8 elements = one dropout id), rows coalesced
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      const int i = tid + KNT * u, row = i / 24, ch = i - row * 24;
      if (row0 + row < a.R) {
        uint4 v = *reinterpret_cast<const uint4*>(O + row * KXP + ch * 8);
        if (masked) {
          float k[8];
          drop_keep8(dst, (unsigned long long)((row0 + row) * KC + ch * 8) >> 3, k);
          uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[q]));
            w[q] = pack2(f.x * k[2 * q], f.y * k[2 * q + 1]);
          }
          v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        *(reinterpret_cast<uint4*>(a.dx[br] + (row0 + row) * KC) + ch) = v;
      }
    }
    __syncthreads();
  }
  // ---- flush.  Pb[j][k] (k < 192) and the stats columns: U[j] = P[j][192] + P[j][194], Td[j] = P[j][193] + P[j][195] (= colsum of the
  // raw d_fused slice; T = alpha Td).  dl' carried no alpha: dW = alpha (gamma (P - U) + beta Td), dgamma = colsum_j (alpha W)(P - U), ...
  float* Pb = reinterpret_cast<float*>(X);     // [48][200] fp32 = 38.4 KB over the X | O tiles (55 KB)
  constexpr int PBP = 200;
#pragma unroll
  for (int m = 0; m < 3; ++m) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = (warp * 3 + i) * 8 + 2 * t;
      *reinterpret_cast<float2*>(Pb + (m * 16 + g) * PBP + c) = make_float2(Pacc[m][i][0], Pacc[m][i][1]);
      *reinterpret_cast<float2*>(Pb + (m * 16 + g + 8) * PBP + c) = make_float2(Pacc[m][i][2], Pacc[m][i][3]);
    }
    if (warp == 0) {
      *reinterpret_cast<float2*>(Pb + (m * 16 + g) * PBP + KC + 2 * t) = make_float2(Sacc[m][0], Sacc[m][1]);
      *reinterpret_cast<float2*>(Pb + (m * 16 + g + 8) * PBP + KC + 2 * t) = make_float2(Sacc[m][2], Sacc[m][3]);
    }
  }
  __syncthreads();
  for (int i = tid * 4; i < KD * KC; i += KNT * 4) {          // 16 B vector reductions: a quarter of the atomic operations
    const int j = i / KC, k = i - j * KC;
    const float U = Pb[j * PBP + KC] + Pb[j * PBP + KC + 2], Td = Pb[j * PBP + KC + 1] + Pb[j * PBP + KC + 3];
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = al * (gam[k + e] * (Pb[j * PBP + k + e] - U) + bet[k + e] * Td);
    red_add_v4(a.dW[br] + i, v[0], v[1], v[2], v[3]);
  }
  for (int k = tid; k < KC; k += KNT) {
    float dg = 0.f, db = 0.f;
    for (int j = 0; j < KD; ++j) {
      const float U = Pb[j * PBP + KC] + Pb[j * PBP + KC + 2], Td = Pb[j * PBP + KC + 1] + Pb[j * PBP + KC + 3];
      const float w = W[j * KC + k] * al;
      dg = fmaf(w, Pb[j * PBP + k] - U, dg);
      db = fmaf(w, Td, db);
    }
    atomicAdd(a.dgamma[br] + k, dg);
    atomicAdd(a.dbeta[br] + k, db);
  }
  if (tid < KD) atomicAdd(a.db[br] + tid, al * (Pb[tid * PBP + KC + 1] + Pb[tid * PBP + KC + 3]));
}

template <typename K>
int opt_in(K kernel, size_t bytes) {
  QV_CHECK(bytes <= 227 * 1024, "kernel needs %zu B of shared memory (> 227 KB)", bytes);
  if (bytes > 48 * 1024) QV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

}  // namespace

bool cmp_fused_ok(int d, int cd) { return d == KC && cd == KD; }

int cmpf_fwd(cudaStream_t s, long R, const void* const* x, const float* const* gamma, const float* const* beta, const float* const* W,
             const float* const* bias, const float* alpha, float eps, void* fused, float* const* stats) {
  if (R <= 0) return 0;
  CmpFwdArgs a{};
  for (int i = 0; i < 4; ++i) {
    a.x[i] = static_cast<const bf16*>(x[i]); a.gamma[i] = gamma[i]; a.beta[i] = beta[i]; a.W[i] = W[i]; a.bias[i] = bias[i];
    a.stats[i] = stats[i];
  }
  a.alpha = alpha; a.out = static_cast<bf16*>(fused); a.R = R; a.eps = eps;
  QV_TRY(opt_in(cmpf_fwd_kernel, CMPF_FWD_SMEM));
  const long ntiles = (R + KTR - 1) / KTR;
  const int gx = (int)max(1L, min(ntiles, (long)(qv_num_sms() * 2 + 3) / 4));
  qv_launch(cmpf_fwd_kernel, dim3(gx, 4), KNT, CMPF_FWD_SMEM, s, a);
  QV_LAUNCH_CHECK();
  return 0;
}

int cmpf_bwd(cudaStream_t s, long R, const void* const* x, const float* const* stats, const float* const* gamma, const float* const* beta,
             const float* const* W, const float* alpha, const void* dfused, void* const* dx, float* const* dW, float* const* db,
             float* const* dgamma, float* const* dbeta, const DropP* drop) {
  if (R <= 0) return 0;
  CmpBwdArgs a{};
  for (int i = 0; i < 4; ++i) {
    a.x[i] = static_cast<const bf16*>(x[i]); a.stats[i] = stats[i]; a.gamma[i] = gamma[i]; a.beta[i] = beta[i]; a.W[i] = W[i];
    a.dW[i] = dW[i]; a.db[i] = db[i]; a.dgamma[i] = dgamma[i]; a.dbeta[i] = dbeta[i]; a.dx[i] = static_cast<bf16*>(dx[i]);
    if (drop) a.drop[i] = drop[i];
  }
  a.dfused = static_cast<const bf16*>(dfused); a.alpha = alpha; a.R = R;
  for (int i = 0; i < 4; ++i) QV_CHECK(((uintptr_t)dW[i] & 15) == 0, "cmpf_bwd: dW[%d] must be 16 B aligned (vector reductions)", i);
  QV_TRY(opt_in(cmpf_bwd_kernel, CMPF_BWD_SMEM));
  const long ntiles = (R + KTR - 1) / KTR;
  const int gx = (int)max(1L, min(ntiles, (long)(qv_num_sms() + 3) / 4));
  qv_launch(cmpf_bwd_kernel, dim3(gx, 4), KNT, CMPF_BWD_SMEM, s, a);
  QV_LAUNCH_CHECK();
  return 0;
}
