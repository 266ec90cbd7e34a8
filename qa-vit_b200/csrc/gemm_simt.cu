// fp32 SIMT GEMM family: the GEMM of fp32 parity runs (1e-4 gate; tensor cores round operands to
// bf16/tf32 and cannot meet it) and of shapes too small for a 128-row tcgen05 tile in bf16 runs.
// One strided kernel:  C[i, j] = sum_k A(i, k) * B(k, j),  64x64x16 tiles, 256 threads, 4x4 per thread,
// optional split along k (grid.z) with atomic accumulation.
#include "kernels.h"

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

struct SimtArgs {
  const void* A;
  long as0, as1;   // A(i, k) = A[i * as0 + k * as1]
  const void* B;
  long bs0, bs1;   // B(k, j) = B[k * bs0 + j * bs1]
  int M, N, K;     // C is [M, N], contraction K
  int kchunk;      // contraction elements per grid.z slice
  GemmEpi e;
  float* db;       // optional: db[i] += scale * sum_k A(i, k)  (TN flavour: column sums of dY)
};

template <typename TA, typename TB, bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256) simt_gemm_kernel(SimtArgs p) {
  QV_PDL_ENTRY();
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const TA* A = static_cast<const TA*>(p.A);
  const TB* B = static_cast<const TB*>(p.B);
  const int tid = threadIdx.x;
  const int i0 = blockIdx.y * TM, j0 = blockIdx.x * TN;
  const int kbeg = blockIdx.z * p.kchunk;
  const int kend = min(p.K, kbeg + p.kchunk);
  const int ty = tid / 16, tx = tid % 16;
  float acc[4][4] = {};
  float rowsum = 0.f;   // for db (thread tid < 64 owns row i0 + tid)

  for (int k0 = kbeg; k0 < kend; k0 += TK) {
    // ---- stage A tile [TM x TK] and B tile [TK x TN]; 1024 elements each, 4 per thread
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = tid + r * 256;
      int ii, kk;
      if (A_KC) { ii = idx / TK; kk = idx % TK; } else { kk = idx / TM; ii = idx % TM; }
      const int gi = i0 + ii, gk = k0 + kk;
      float v = 0.f;
      if (gi < p.M && gk < kend) v = ldf(A + (long)gi * p.as0 + (long)gk * p.as1);
      As[kk][ii] = v;
      int jj, kb;
      if (B_KC) { jj = idx / TK; kb = idx % TK; } else { kb = idx / TN; jj = idx % TN; }
      const int gj = j0 + jj, gk2 = k0 + kb;
      float w = 0.f;
      if (gj < p.N && gk2 < kend) w = ldf(B + (long)gk2 * p.bs0 + (long)gj * p.bs1);
      Bs[kb][jj] = w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = As[kk][ty * 4 + r];
#pragma unroll
      for (int c = 0; c < 4; ++c) b[c] = Bs[kk][tx * 4 + c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
    }
    if (p.db != nullptr && blockIdx.x == 0 && tid < TM) {
#pragma unroll
      for (int kk = 0; kk < TK; ++kk) rowsum += As[kk][tid];
    }
    __syncthreads();
  }

  const GemmEpi& e = p.e;
  const float s_pre = e.scale_pre ? *e.scale_pre : 1.f;
  const float s_res = e.scale_res ? *e.scale_res : 1.f;
  const bool split = gridDim.z > 1;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int gi = i0 + ty * 4 + r;
    if (gi >= p.M) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int gj = j0 + tx * 4 + c;
      if (gj >= p.N) continue;
      float val = acc[r][c];
      if (e.bias && blockIdx.z == 0) val += e.bias[gj];
      val *= s_pre;
      if (e.gmul) {
        const long gi_ = (long)gi * e.ldg + gj;
        const float gv = e.g_bf16 ? ldf(static_cast<const bf16*>(e.gmul) + gi_) : static_cast<const float*>(e.gmul)[gi_];
        val *= e.gmul_raw ? gv : gelu_grad_f(gv);
      }
      const float pre = val;
      if (e.gelu && e.gelu_dgrad) val = gelu_grad_f(pre);
      if (e.C) {
        if (e.c_f32) {
          float* c_ = static_cast<float*>(e.C) + (long)gi * e.ldc + gj;
          if (split) atomicAdd(c_, val);
          else if (e.c_accum) *c_ += val;
          else *c_ = val;
        } else {
          stf(static_cast<bf16*>(e.C) + (long)gi * e.ldc + gj, val);
        }
      }
      if (e.C2) {
        float rv = 0.f;
        if (e.resid) {
          const long ri = (long)gi * e.ldr + gj;
          rv = e.r_bf16 ? ldf(static_cast<const bf16*>(e.resid) + ri) : static_cast<const float*>(e.resid)[ri];
        }
        float v2 = e.gelu ? gelu_f(pre) : rv + s_res * val;
        if (e.c2_f32) static_cast<float*>(e.C2)[(long)gi * e.ldc2 + gj] = v2;
        else stf(static_cast<bf16*>(e.C2) + (long)gi * e.ldc2 + gj, v2);
      }
    }
  }
  if (p.db != nullptr && blockIdx.x == 0 && tid < TM && i0 + tid < p.M) atomicAdd(p.db + i0 + tid, rowsum * s_pre);
}

template <bool A_KC, bool B_KC>
int launch(cudaStream_t s, int dtA, int dtB, SimtArgs& a, int splits) {
  dim3 grid(cdiv(a.N, TN), cdiv(a.M, TM), splits);
  a.kchunk = cdiv(cdiv(a.K, splits), TK) * TK;
  grid.z = cdiv(a.K, a.kchunk);
  if (dtA == QV_F32 && dtB == QV_F32) qv_launch(simt_gemm_kernel<float, float, A_KC, B_KC>, grid, 256, 0, s, a);
  else if (dtA == QV_BF16 && dtB == QV_F32) qv_launch(simt_gemm_kernel<bf16, float, A_KC, B_KC>, grid, 256, 0, s, a);
  else if (dtA == QV_BF16 && dtB == QV_BF16) qv_launch(simt_gemm_kernel<bf16, bf16, A_KC, B_KC>, grid, 256, 0, s, a);
  else { qv_set_error("simt_gemm: unsupported dtype pair %d/%d", dtA, dtB); return 1; }
  QV_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int simt_gemm_nt(cudaStream_t s, int dtA, const void* A, int lda, int M, int N, int K, const float* W, const GemmEpi& e) {
  if (M <= 0) return 0;
  QV_CHECK(!(e.c_accum && !e.c_f32), "gemm: accumulate needs an fp32 C");
  SimtArgs a{A, lda, 1, W, 1, K, M, N, K, 0, e, nullptr};
  return launch<true, true>(s, dtA, QV_F32, a, 1);
}

int simt_gemm_nn(cudaStream_t s, int dtA, const void* dY, int ldy, int M, int N, int K, const float* W, const GemmEpi& e) {
  if (M <= 0) return 0;
  QV_CHECK(!(e.c_accum && !e.c_f32), "gemm: accumulate needs an fp32 C");
  // dX[M, K] = dY[M, N] W[N, K]: contraction over N
  SimtArgs a{dY, ldy, 1, W, K, 1, M, K, N, 0, e, nullptr};
  return launch<true, false>(s, dtA, QV_F32, a, 1);
}

int simt_gemm_tn(cudaStream_t s, int dt, const void* dY, int ldy, const void* X, int ldx, int M, int N, int K, float* dW,
                 float* db, const float* scale) {
  if (M <= 0) return 0;
  // dW[N, K] += dY^T X: A(i = n, k = m) = dY[m * ldy + n]; B(k = m, j) = X[m * ldx + j]
  GemmEpi e;
  e.C = dW; e.ldc = K; e.c_f32 = 1; e.c_accum = 1; e.scale_pre = scale;
  SimtArgs a{dY, 1, ldy, X, ldx, 1, N, K, M, 0, e, db};
  const int tiles = cdiv(N, TM) * cdiv(K, TN);
  int splits = (2 * qv_num_sms() + tiles - 1) / tiles;
  splits = max(1, min(splits, cdiv(M, 4 * TK)));
  if (splits == 1) splits = 2;   // keep the atomic (+=) path: dW always accumulates
  if (dt == QV_F32) return launch<false, false>(s, QV_F32, QV_F32, a, splits);
  return launch<false, false>(s, QV_BF16, QV_BF16, a, splits);
}
