// Channel-group attention (H:559-595) and the batch-invariant bank projections.  See attn.cu for the other branches.
#include "kernels.h"

namespace {
constexpr int NT = 128;  // threads per CTA
// compile-time shapes of every shipped config (d = 192, 6 groups, 4 heads): 32 channels / group -> 16 -> 4 heads of 4
constexpr int CG = 32, CPG = 16, NH = 4, HDC = 4;
template <typename K>
int set_smem(K kernel, size_t bytes) {
  QV_CHECK(bytes <= 227 * 1024, "attention kernel needs %zu B of shared memory (> 227 KB): config not supported", bytes);
  if (bytes > 48 * 1024) QV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}
}  // namespace

// =====================================================================================================
// Channel-group attention
// =====================================================================================================
namespace {

struct CgaLay { int NKV, SP, QC, oX, oQ, oK, oV, oP, odS, odO, odQ, odK, odV, oW, odW, odkb, total; };

// per-group projections for image b: q/k/v[Nt, cpg] from xg[Nt, cg]; bank rows appended to k, v.
template <typename T>
__device__ void cga_project(const CgaP& p, const CgaLay& ly, float* sm, int b, int g) {
  const int tid = threadIdx.x, Nt = p.Nt;
  constexpr int cg = CG, cpg = CPG;
  const T* xn = static_cast<const T*>(p.xn);
  float *X = sm + ly.oX, *Q = sm + ly.oQ, *K = sm + ly.oK, *V = sm + ly.oV;
  const float* W = sm + ly.oW;  // [3][cpg][cg] then biases [3][cpg]
  const float* Bv = W + 3 * cpg * cg;
  for (int idx = tid; idx < Nt * cg; idx += NT) {
    const int n = idx / cg, c = idx % cg;
    X[n * (cg + 1) + c] = ldf(xn + ((long)b * Nt + n) * p.ldx + g * cg + c);
  }
  __syncthreads();
  for (int idx = tid; idx < Nt * cpg; idx += NT) {
    const int n = idx / cpg, o = idx % cpg;
    float aq = Bv[o], ak = Bv[cpg + o], av = Bv[2 * cpg + o];
    for (int c = 0; c < cg; ++c) {
      const float x = X[n * (cg + 1) + c];
      aq = fmaf(x, W[o * cg + c], aq);
      ak = fmaf(x, W[(cpg + o) * cg + c], ak);
      av = fmaf(x, W[(2 * cpg + o) * cg + c], av);
    }
    Q[n * cpg + o] = aq; K[n * cpg + o] = ak; V[n * cpg + o] = av;
  }
  for (int idx = tid; idx < p.kb * cpg; idx += NT) {
    K[Nt * cpg + idx] = p.kbp[idx];
    V[Nt * cpg + idx] = p.vbp[idx];
  }
  __syncthreads();
}

// P[h][i][j] for query rows q0..q0+n of the current group.  Ends with __syncthreads().
__device__ void cga_scores(const CgaP& p, const CgaLay& ly, float* sm, int q0, int n) {
  const int tid = threadIdx.x, NKV = ly.NKV, SP = ly.SP;
  constexpr int cpg = CPG, hdc = HDC;
  const float *Q = sm + ly.oQ, *K = sm + ly.oK;
  float* P = sm + ly.oP;
  const float scale = rsqrtf((float)hdc);
  for (int idx = tid; idx < NH * n * NKV; idx += NT) {
    const int h = idx / (n * NKV), r = idx % (n * NKV), i = r / NKV, j = r % NKV;
    float a = 0.f;
    _Pragma("unroll") for (int d = 0; d < hdc; ++d) a = fmaf(Q[(q0 + i) * cpg + h * hdc + d], K[j * cpg + h * hdc + d], a);
    P[(h * ly.QC + i) * SP + j] = a * scale;
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  for (int r = warp; r < NH * n; r += NT / 32) {
    const int h = r / n, i = r % n;
    float* row = P + (h * ly.QC + i) * SP;
    float m = -INFINITY;
    for (int j = lane; j < NKV; j += 32) m = fmaxf(m, row[j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < NKV; j += 32) { const float e = __expf(row[j] - m); row[j] = e; s += e; }
    s = 1.f / warp_sum(s);
    for (int j = lane; j < NKV; j += 32) row[j] *= s;
  }
  __syncthreads();
}

__device__ void cga_load_weights(const CgaP& p, const CgaLay& ly, float* sm) {
  float* W = sm + ly.oW;
  constexpr int n = CPG * CG;
  for (int idx = threadIdx.x; idx < n; idx += NT) { W[idx] = p.Wq[idx]; W[n + idx] = p.Wk[idx]; W[2 * n + idx] = p.Wv[idx]; }
  for (int idx = threadIdx.x; idx < CPG; idx += NT) {
    W[3 * n + idx] = p.bq[idx]; W[3 * n + CPG + idx] = p.bk[idx]; W[3 * n + 2 * CPG + idx] = p.bv[idx];
  }
}

// dropout id of probability (token row r, group g, head h, key j): unique within the site (NKV <= 128, G * NH <= 32)
__device__ __forceinline__ unsigned long long cga_drop_id(const CgaP& p, long r, int g, int h, int j) {
  return ((unsigned long long)((r * p.G + g) * NH + h) << 7) | (unsigned)j;
}

template <typename T>
__global__ void __launch_bounds__(NT) cga_fwd_kernel(CgaP p, CgaLay ly) {
  QV_PDL_ENTRY();
  extern __shared__ float sm[];
  const int tid = threadIdx.x;
  constexpr int cpg = CPG, hdc = HDC;
  cga_load_weights(p, ly, sm);
  __syncthreads();
  T* out = static_cast<T*>(p.out);
  const float *V = sm + ly.oV, *P = sm + ly.oP;
  for (int task = blockIdx.x; task < p.B * p.G; task += gridDim.x) {
    const int b = task / p.G, g = task % p.G;
    cga_project<T>(p, ly, sm, b, g);
    for (int q0 = 0; q0 < p.Nt; q0 += ly.QC) {
      const int n = min(ly.QC, p.Nt - q0);
      cga_scores(p, ly, sm, q0, n);
      if (p.drop.p > 0.f) {   // SDPA dropout_p on the probabilities
        const DropState ds = drop_state(p.drop);
        float* Pw = sm + ly.oP;
        for (int idx = tid; idx < NH * n * ly.NKV; idx += NT) {
          const int h = idx / (n * ly.NKV), r = idx % (n * ly.NKV), i = r / ly.NKV, j = r % ly.NKV;
          Pw[(h * ly.QC + i) * ly.SP + j] *= drop_keep1(ds, cga_drop_id(p, (long)b * p.Nt + q0 + i, g, h, j));
        }
        __syncthreads();
      }
      for (int idx = tid; idx < n * cpg; idx += NT) {
        const int i = idx / cpg, o = idx % cpg, h = o / hdc;
        float a = 0.f;
        for (int j = 0; j < ly.NKV; ++j) a = fmaf(P[(h * ly.QC + i) * ly.SP + j], V[j * cpg + o], a);
        stf(out + ((long)b * p.Nt + q0 + i) * p.ldo + g * cpg + o, a);
      }
      __syncthreads();
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(NT) cga_bwd_kernel(CgaP p, CgaLay ly) {
  QV_PDL_ENTRY();
  extern __shared__ float sm[];
  const int tid = threadIdx.x, Nt = p.Nt, NKV = ly.NKV, SP = ly.SP;
  constexpr int cpg = CPG, cg = CG, hdc = HDC;
  cga_load_weights(p, ly, sm);
  const int nW = 3 * cpg * cg + 3 * cpg;
  for (int idx = tid; idx < nW; idx += NT) sm[ly.odW + idx] = 0.f;
  for (int idx = tid; idx < 2 * p.kb * cpg; idx += NT) sm[ly.odkb + idx] = 0.f;
  __syncthreads();
  const T* dout = static_cast<const T*>(p.dout);
  float *X = sm + ly.oX, *Q = sm + ly.oQ, *K = sm + ly.oK, *V = sm + ly.oV, *P = sm + ly.oP, *dS = sm + ly.odS;
  float *dO = sm + ly.odO, *dQ = sm + ly.odQ, *dK = sm + ly.odK, *dV = sm + ly.odV;
  const float* W = sm + ly.oW;
  float* dW = sm + ly.odW;
  const float scale = rsqrtf((float)hdc);
  const bool drop = p.drop.p > 0.f;
  DropState dst{};
  if (drop) dst = drop_state(p.drop);
  for (int task = blockIdx.x; task < p.B * p.G; task += gridDim.x) {
    const int b = task / p.G, g = task % p.G;
    cga_project<T>(p, ly, sm, b, g);
    for (int idx = tid; idx < NKV * cpg; idx += NT) { dK[idx] = 0.f; dV[idx] = 0.f; }
    for (int q0 = 0; q0 < Nt; q0 += ly.QC) {
      const int n = min(ly.QC, Nt - q0);
      for (int idx = tid; idx < n * cpg; idx += NT) {
        const int i = idx / cpg, o = idx % cpg;
        dO[idx] = ldf(dout + ((long)b * Nt + q0 + i) * p.lddo + g * cpg + o);
      }
      cga_scores(p, ly, sm, q0, n);   // also orders the dO stores
      for (int idx = tid; idx < NH * n * NKV; idx += NT) {
        const int h = idx / (n * NKV), r = idx % (n * NKV), i = r / NKV, j = r % NKV;
        float a = 0.f;
        _Pragma("unroll") for (int d = 0; d < hdc; ++d) a = fmaf(dO[i * cpg + h * hdc + d], V[j * cpg + h * hdc + d], a);
        if (drop) a *= drop_keep1(dst, cga_drop_id(p, (long)b * Nt + q0 + i, g, h, j));   // d(P_dropped) -> dP
        dS[(h * ly.QC + i) * SP + j] = a;
      }
      __syncthreads();
      {
        const int lane = tid & 31, warp = tid >> 5;
        for (int r = warp; r < NH * n; r += NT / 32) {
          const int h = r / n, i = r % n;
          float* ds = dS + (h * ly.QC + i) * SP;
          float* pr = P + (h * ly.QC + i) * SP;
          float s = 0.f;
          for (int j = lane; j < NKV; j += 32) s += ds[j] * pr[j];
          s = warp_sum(s);
          for (int j = lane; j < NKV; j += 32) {
            const float pv = pr[j];
            ds[j] = pv * (ds[j] - s) * scale;
            if (drop) pr[j] = pv * drop_keep1(dst, cga_drop_id(p, (long)b * Nt + q0 + i, g, h, j));   // dV below needs the dropped P
          }
        }
      }
      __syncthreads();
      for (int idx = tid; idx < n * cpg; idx += NT) {
        const int i = idx / cpg, o = idx % cpg, h = o / hdc;
        float a = 0.f;
        for (int j = 0; j < NKV; ++j) a = fmaf(dS[(h * ly.QC + i) * SP + j], K[j * cpg + o], a);
        dQ[(q0 + i) * cpg + o] = a;
      }
      for (int idx = tid; idx < NKV * cpg; idx += NT) {
        const int j = idx / cpg, o = idx % cpg, h = o / hdc;
        float av = 0.f, ak = 0.f;
        for (int i = 0; i < n; ++i) {
          av = fmaf(P[(h * ly.QC + i) * SP + j], dO[i * cpg + o], av);
          ak = fmaf(dS[(h * ly.QC + i) * SP + j], Q[(q0 + i) * cpg + o], ak);
        }
        dV[idx] += av;
        dK[idx] += ak;
      }
      __syncthreads();
    }
    // bank rows -> d(projected bank); own rows -> projections backward
    for (int idx = tid; idx < p.kb * cpg; idx += NT) {
      sm[ly.odkb + idx] += dK[Nt * cpg + idx];
      sm[ly.odkb + p.kb * cpg + idx] += dV[Nt * cpg + idx];
    }
    // dx[n, c] += sum_o dq[n,o] Wq[o,c] + dk[n,o] Wk[o,c] + dv[n,o] Wv[o,c]
    for (int idx = tid; idx < Nt * cg; idx += NT) {
      const int n = idx / cg, c = idx % cg;
      float a = 0.f;
      for (int o = 0; o < cpg; ++o) {
        a = fmaf(dQ[n * cpg + o], W[o * cg + c], a);
        a = fmaf(dK[n * cpg + o], W[(cpg + o) * cg + c], a);
        a = fmaf(dV[n * cpg + o], W[(2 * cpg + o) * cg + c], a);
      }
      p.dxn[((long)b * Nt + n) * p.lddx + g * cg + c] += a;
    }
    // dW[o, c] += sum_n d{q,k,v}[n, o] x[n, c] ; db[o] += sum_n d{q,k,v}[n, o]
    for (int idx = tid; idx < 3 * cpg * cg; idx += NT) {
      const int which = idx / (cpg * cg), r = idx % (cpg * cg), o = r / cg, c = r % cg;
      const float* dsrc = which == 0 ? dQ : (which == 1 ? dK : dV);
      float a = 0.f;
      for (int n = 0; n < Nt; ++n) a = fmaf(dsrc[n * cpg + o], X[n * (cg + 1) + c], a);
      dW[idx] += a;
    }
    for (int idx = tid; idx < 3 * cpg; idx += NT) {
      const int which = idx / cpg, o = idx % cpg;
      const float* dsrc = which == 0 ? dQ : (which == 1 ? dK : dV);
      float a = 0.f;
      for (int n = 0; n < Nt; ++n) a += dsrc[n * cpg + o];
      dW[3 * cpg * cg + idx] += a;
    }
    __syncthreads();
  }
  const int n1 = cpg * cg;
  for (int idx = tid; idx < n1; idx += NT) {
    atomicAdd(p.dWq + idx, dW[idx]); atomicAdd(p.dWk + idx, dW[n1 + idx]); atomicAdd(p.dWv + idx, dW[2 * n1 + idx]);
  }
  for (int idx = tid; idx < cpg; idx += NT) {
    atomicAdd(p.dbq + idx, dW[3 * n1 + idx]); atomicAdd(p.dbk + idx, dW[3 * n1 + cpg + idx]);
    atomicAdd(p.dbv + idx, dW[3 * n1 + 2 * cpg + idx]);
  }
  for (int idx = tid; idx < p.kb * cpg; idx += NT) {
    atomicAdd(p.dkbp + idx, sm[ly.odkb + idx]);
    atomicAdd(p.dvbp + idx, sm[ly.odkb + p.kb * cpg + idx]);
  }
}

CgaLay cga_layout_qc(const CgaP& p, bool bwd, int qc) {
  CgaLay ly{};
  ly.NKV = p.Nt + p.kb;
  ly.SP = ly.NKV + 1;
  ly.QC = p.Nt < qc ? p.Nt : qc;
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
  ly.oX = take(p.Nt * (p.cg + 1));
  ly.oQ = take(p.Nt * p.cpg);
  ly.oK = take(ly.NKV * p.cpg);
  ly.oV = take(ly.NKV * p.cpg);
  ly.oP = take(p.H * ly.QC * ly.SP);
  ly.odS = take(bwd ? p.H * ly.QC * ly.SP : 0);
  ly.odO = take(bwd ? ly.QC * p.cpg : 0);
  ly.odQ = take(bwd ? p.Nt * p.cpg : 0);
  ly.odK = take(bwd ? ly.NKV * p.cpg : 0);
  ly.odV = take(bwd ? ly.NKV * p.cpg : 0);
  ly.oW = take(3 * p.cpg * p.cg + 3 * p.cpg);
  ly.odW = take(bwd ? 3 * p.cpg * p.cg + 3 * p.cpg : 0);
  ly.odkb = take(bwd ? 2 * p.kb * p.cpg : 0);
  ly.total = o;
  return ly;
}
// query-chunk height: 32 rows, halved while the probability tiles [heads][QC][Nkv] do not fit (196 + 16 keys at 224 / 16)
CgaLay cga_layout(const CgaP& p, bool bwd) {
  int qc = 32;
  CgaLay ly = cga_layout_qc(p, bwd, qc);
  while (qc > 4 && (size_t)ly.total * sizeof(float) > 200 * 1024) {
    qc >>= 1;
    ly = cga_layout_qc(p, bwd, qc);
  }
  return ly;
}

}  // namespace

int cga_fwd(cudaStream_t s, int dt, const CgaP& p) {
  if (p.B <= 0) return 0;
  if (dt == QV_BF16 && cga_mma_ok(p)) return cga_mma_fwd(s, p);   // tensor-core path (cga_mma.cu)
  if (dt == QV_BF16 && cga_mma64_ok(p)) return cga_mma64_fwd(s, p);   // 64-token blocks (cga_mma64.cu)
  QV_CHECK(p.cg == 32 && p.cpg == 16 && p.H == 4, "CGA kernels are instantiated for 32 -> 16 channels per group and 4 heads (got %d -> %d, %d heads)", p.cg, p.cpg, p.H);
  const CgaLay ly = cga_layout(p, false);
  const size_t smem = (size_t)ly.total * sizeof(float);
  const int occ = max(1, (int)(200 * 1024 / (smem + 1024)));
  const int grid = min(p.B * p.G, qv_num_sms() * min(occ, 8));
  if (dt == QV_F32) { QV_TRY(set_smem(cga_fwd_kernel<float>, smem)); qv_launch(cga_fwd_kernel<float>, grid, NT, smem, s, p, ly); }
  else { QV_TRY(set_smem(cga_fwd_kernel<bf16>, smem)); qv_launch(cga_fwd_kernel<bf16>, grid, NT, smem, s, p, ly); }
  QV_LAUNCH_CHECK();
  return 0;
}

int cga_bwd(cudaStream_t s, int dt, const CgaP& p) {
  if (p.B <= 0) return 0;
  if (dt == QV_BF16 && cga_mma_ok(p)) return cga_mma_bwd(s, p);
  if (dt == QV_BF16 && cga_mma64_ok(p)) return cga_mma64_bwd(s, p);
  QV_CHECK(p.cg == 32 && p.cpg == 16 && p.H == 4, "CGA kernels are instantiated for 32 -> 16 channels per group and 4 heads");
  const CgaLay ly = cga_layout(p, true);
  const size_t smem = (size_t)ly.total * sizeof(float);
  const int occ = max(1, (int)(200 * 1024 / (smem + 1024)));
  const int grid = min(p.B * p.G, qv_num_sms() * min(occ, 8));
  if (dt == QV_F32) { QV_TRY(set_smem(cga_bwd_kernel<float>, smem)); qv_launch(cga_bwd_kernel<float>, grid, NT, smem, s, p, ly); }
  else { QV_TRY(set_smem(cga_bwd_kernel<bf16>, smem)); qv_launch(cga_bwd_kernel<bf16>, grid, NT, smem, s, p, ly); }
  QV_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================
// Bank projections: Y[rows, N] = X[rows, K] W^T + b with rows = 16 (batch invariant; the reference recomputes
// them per image, H:617-619).  Single small grid, fp32.
// =====================================================================================================
namespace {
struct SmallLin {
  const float* X; const float* W; const float* b; float* Y;      // forward
  const float* dY; float* dW; float* db; float* dX;              // backward (accumulated)
};
struct SmallLin2 { SmallLin p[2]; };

// One warp per output column n for every row (W[n, :] is read once): lanes stride over K (coalesced), shuffle reduction.
// blockIdx.y selects the problem: the K and V projections of a bank snapshot go out as ONE launch.
// These kernels move a few hundred KB: what they cost is dependent memory latencies, so every loop is shaped to have all the
// loads of a pass in flight at once (the rolled loops spent one L2 round trip per iteration: 16 / 41 us per launch).
__global__ void __launch_bounds__(128) small_linear_fwd_kernel(SmallLin2 pp, int rows, int K, int N) {
  QV_PDL_ENTRY();
  const SmallLin& q = pp.p[blockIdx.y];
  const float* __restrict__ X = q.X;
  const float* __restrict__ W = q.W;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= N) return;
  const int n = warp;
  for (int r0 = 0; r0 < rows; r0 += 16) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 0.f;
    for (int k0 = 0; k0 < K; k0 += 128) {       // 4 k per lane and pass: 4 + 64 independent loads
      float w[4], x[4][16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + j * 32 + lane;
        w[j] = k < K ? W[n * K + k] : 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) x[j][i] = (k < K && r0 + i < rows) ? X[(r0 + i) * K + k] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(x[j][i], w[j], a[i]);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float v = warp_sum(a[i]);
      if (lane == 0 && r0 + i < rows) q.Y[(r0 + i) * N + n] = v + (q.b ? q.b[n] : 0.f);
    }
  }
}
// dW[n,k] += sum_r dY[r,n] X[r,k]; db[n] += sum_r dY[r,n]; dX[r,k] += sum_n dY[r,n] W[n,k].
// blockIdx.z = a slice of 32 output columns n for the dX sum (atomic accumulation); slice 0 also does dW and db.
__global__ void __launch_bounds__(128) small_linear_bwd_kernel(SmallLin2 pp, int rows, int K, int N) {
  QV_PDL_ENTRY();
  const SmallLin& q = pp.p[blockIdx.y];
  const float* __restrict__ X = q.X;
  const float* __restrict__ W = q.W;
  const float* __restrict__ dY = q.dY;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.z == 0) {
    if (idx < N * K) {
      const int n = idx / K, k = idx % K;
      float a = 0.f;
      for (int r0 = 0; r0 < rows; r0 += 16) {
        float dy[16], x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const bool ok = r0 + i < rows;
          dy[i] = ok ? dY[(r0 + i) * N + n] : 0.f;
          x[i] = ok ? X[(r0 + i) * K + k] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) a = fmaf(dy[i], x[i], a);
      }
      q.dW[idx] += a;
    }
    if (idx < N) {
      float a = 0.f;
      for (int r = 0; r < rows; ++r) a += dY[r * N + idx];
      q.db[idx] += a;
    }
  }
  if (idx < rows * K) {
    const int r = idx / K, k = idx % K, n0 = blockIdx.z * 32;
    float dy[32], w[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const bool ok = n0 + i < N;
      dy[i] = ok ? dY[r * N + n0 + i] : 0.f;
      w[i] = ok ? W[(n0 + i) * K + k] : 0.f;                          // lanes = consecutive k: coalesced
    }
    float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 32; ++i) a[i & 3] = fmaf(dy[i], w[i], a[i & 3]);
    atomicAdd(q.dX + idx, (a[0] + a[1]) + (a[2] + a[3]));
  }
}
}  // namespace

int small_linear_fwd2(cudaStream_t s, int rows, int K, int N, const float* X0, const float* W0, const float* b0, float* Y0,
                      const float* X1, const float* W1, const float* b1, float* Y1) {
  SmallLin2 pp{};
  pp.p[0].X = X0; pp.p[0].W = W0; pp.p[0].b = b0; pp.p[0].Y = Y0;
  pp.p[1].X = X1; pp.p[1].W = W1; pp.p[1].b = b1; pp.p[1].Y = Y1;
  qv_launch(small_linear_fwd_kernel, dim3(cdiv(N * 32, 128), X1 ? 2 : 1), 128, 0, s, pp, rows, K, N);
  QV_LAUNCH_CHECK();
  return 0;
}
int small_linear_fwd(cudaStream_t s, const float* X, int rows, int K, const float* W, const float* b, int N, float* Y) {
  return small_linear_fwd2(s, rows, K, N, X, W, b, Y, nullptr, nullptr, nullptr, nullptr);
}

// NOTE: the two problems of a pair must not share an accumulated output (both bank projections add into different dX here)
int small_linear_bwd2(cudaStream_t s, int rows, int K, int N, const float* X0, const float* W0, const float* dY0, float* dW0, float* db0,
                      float* dX0, const float* X1, const float* W1, const float* dY1, float* dW1, float* db1, float* dX1) {
  SmallLin2 pp{};
  pp.p[0].X = X0; pp.p[0].W = W0; pp.p[0].dY = dY0; pp.p[0].dW = dW0; pp.p[0].db = db0; pp.p[0].dX = dX0;
  pp.p[1].X = X1; pp.p[1].W = W1; pp.p[1].dY = dY1; pp.p[1].dW = dW1; pp.p[1].db = db1; pp.p[1].dX = dX1;
  const int n = max(N * K, rows * K);
  qv_launch(small_linear_bwd_kernel, dim3(cdiv(n, 128), X1 ? 2 : 1, cdiv(N, 32)), 128, 0, s, pp, rows, K, N);
  QV_LAUNCH_CHECK();
  return 0;
}
int small_linear_bwd(cudaStream_t s, const float* X, int rows, int K, const float* W, int N, const float* dY, float* dW,
                     float* db, float* dX_accum) {
  return small_linear_bwd2(s, rows, K, N, X, W, dY, dW, db, dX_accum, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}
