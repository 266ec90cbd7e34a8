// The repo's attention variants as single fused kernels (forward and backward):
//   mode 0  SWA   (H:441-469): window partition -> Linformer K' = E_k^T K, V' = E_v^T V over the window's
//                 tokens -> concat 16 bank keys -> softmax(QK^T/sqrt(hd)) V -> window reverse
//   mode 1  MSDA  (H:496-532): K,V from the pooled dilated tokens (rows >= NM are the zero padding of
//                 H:507-509 and are skipped), Linformer, bank concat, full-resolution Q
//   mode 2  Cross (H:613-626): K,V = projected bank (batch invariant), no Linformer
// (CGA lives in cga.cu.)  Every attention problem here fits one KV tile (Nkv <= 80), so softmax is a single pass in
// shared memory.  One CTA walks (window, head) tasks grid-stride; the batch reductions (dE_k, dE_v, d bank)
// accumulate in shared memory across a CTA's tasks and are flushed once with atomics.
// Thread mapping: warp = one query / key row, lanes = head-dim channels or keys; head_dim is a template constant
// (48 in every shipped config) so all index arithmetic is multiply-shift and the inner products unroll.
#include "kernels.h"

namespace {

constexpr int NT = 128;  // threads per CTA
constexpr int NW = NT / 32;

struct Lay {  // shared-memory carve-up (float offsets), computed on host and passed by value
  int HDP, NKV, SP, QC, nq, L;
  int e_smem;   // E_k / E_v (and their gradient accumulators) staged in shared memory; 0: read from / atomically added to global
  int oEk, oEv, odEk, odEv, oKs, oVs, oKf, oVf, odKf, odVf, oQ, odO, oP, odS, odbk, odbv, oRow, total;
};

__device__ __forceinline__ int q_row(const AttnP& p, int w, int i) {
  if (p.mode == 0) {
    const int nws = p.side / p.ws, nW = nws * nws;
    const int b = w / nW, wi = w % nW;
    const int r = (wi / nws) * p.ws + i / p.ws, c = (wi % nws) * p.ws + i % p.ws;
    return b * p.Nt + r * p.side + c;
  }
  return w * p.Nt + i;
}

// Row indices of the task (queries, then kv sources) -> smem, once per task instead of once per element.
__device__ __forceinline__ void task_rows(const AttnP& p, const Lay& ly, int* rows, int w) {
  for (int i = threadIdx.x; i < ly.nq; i += NT) rows[i] = q_row(p, w, i);
  if (p.mode != 2) {
    const int nl = (p.mode == 1) ? p.NM : ly.L;
    for (int l = threadIdx.x; l < nl; l += NT) rows[ly.nq + l] = (p.mode == 0) ? q_row(p, w, l) : w * p.NM + l;
  }
}

// Loads K/V sources, builds Kf/Vf = [E^T Ksrc ; bank] for task (w, h).  Ends with __syncthreads().
template <typename T, int HD>
__device__ void build_kv(const AttnP& p, const Lay& ly, float* sm, const int* rows, int h) {
  const int tid = threadIdx.x, HDP = ly.HDP, D = p.H * HD;
  float *Ks = sm + ly.oKs, *Vs = sm + ly.oVs, *Kf = sm + ly.oKf, *Vf = sm + ly.oVf;
  if (p.mode == 2) {
    for (int idx = tid; idx < p.kb * HD; idx += NT) {
      const int j = idx / HD, d = idx % HD;
      Kf[j * HDP + d] = p.kc[j * D + h * HD + d];
      Vf[j * HDP + d] = p.vc[j * D + h * HD + d];
    }
    __syncthreads();
    return;
  }
  const T* kv = static_cast<const T*>(p.kv);
  for (int idx = tid; idx < ly.L * HD; idx += NT) {
    const int l = idx / HD, d = idx % HD;
    const long r = rows[ly.nq + l];
    Ks[l * HDP + d] = ldf(kv + r * p.ldkv + p.kcol + h * HD + d);
    Vs[l * HDP + d] = ldf(kv + r * p.ldkv + p.vcol + h * HD + d);
  }
  for (int idx = tid; idx < p.kb * HD; idx += NT) {
    const int j = idx / HD, d = idx % HD;
    Kf[(p.klin + j) * HDP + d] = p.bank_k[j * D + h * HD + d];
    Vf[(p.klin + j) * HDP + d] = p.bank_v[j * D + h * HD + d];
  }
  __syncthreads();
  const float *Ek = ly.e_smem ? sm + ly.oEk : p.Ek, *Ev = ly.e_smem ? sm + ly.oEv : p.Ev;
  for (int idx = tid; idx < p.klin * HD; idx += NT) {
    const int j = idx / HD, d = idx % HD;
    float ak = 0.f, av = 0.f;
    for (int l = 0; l < ly.L; ++l) {
      ak = fmaf(Ek[l * p.klin + j], Ks[l * HDP + d], ak);
      av = fmaf(Ev[l * p.klin + j], Vs[l * HDP + d], av);
    }
    Kf[j * HDP + d] = ak;
    Vf[j * HDP + d] = av;
  }
  __syncthreads();
}

// P = softmax(scale * Q Kf^T) for the loaded Q chunk of n rows: one warp per query row, lanes over keys.
template <int HD>
__device__ void scores_softmax(const Lay& ly, float* sm, int n, float scale) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, HDP = ly.HDP, NKV = ly.NKV, SP = ly.SP;
  const float *Q = sm + ly.oQ, *Kf = sm + ly.oKf;
  float* P = sm + ly.oP;
  for (int i = warp; i < n; i += NW) {
    float m = -INFINITY;
    for (int j = lane; j < NKV; j += 32) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) a = fmaf(Q[i * HDP + d], Kf[j * HDP + d], a);
      a *= scale;
      P[i * SP + j] = a;
      m = fmaxf(m, a);
    }
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < NKV; j += 32) { const float e = __expf(P[i * SP + j] - m); P[i * SP + j] = e; s += e; }
    s = 1.f / warp_sum(s);
    for (int j = lane; j < NKV; j += 32) P[i * SP + j] *= s;
  }
  __syncthreads();
}

// dropout id of probability (query row r, head h, key j): unique within the site (NKV <= 128)
__device__ __forceinline__ unsigned long long drop_id(const AttnP& p, long r, int h, int j) {
  return ((unsigned long long)(r * p.H + h) << 7) | (unsigned)j;
}

template <typename T, int HD>
__global__ void __launch_bounds__(NT) attn_fwd_kernel(AttnP p, Lay ly, int ntask) {
  QV_PDL_ENTRY();
  extern __shared__ float sm[];
  const int tid = threadIdx.x, HDP = ly.HDP;
  int* rows = reinterpret_cast<int*>(sm + ly.oRow);
  if (p.mode != 2 && ly.e_smem) {
    for (int idx = tid; idx < ly.L * p.klin; idx += NT) { sm[ly.oEk + idx] = p.Ek[idx]; sm[ly.oEv + idx] = p.Ev[idx]; }
  }
  const float scale = rsqrtf((float)HD);
  const T* q = static_cast<const T*>(p.q);
  T* out = static_cast<T*>(p.out);
  float *Q = sm + ly.oQ, *P = sm + ly.oP, *Vf = sm + ly.oVf;
  for (int task = blockIdx.x; task < ntask; task += gridDim.x) {
    const int w = task / p.H, h = task % p.H;
    __syncthreads();
    task_rows(p, ly, rows, w);
    __syncthreads();
    build_kv<T, HD>(p, ly, sm, rows, h);
    for (int q0 = 0; q0 < ly.nq; q0 += ly.QC) {
      const int n = min(ly.QC, ly.nq - q0);
      for (int idx = tid; idx < n * HD; idx += NT) {
        const int i = idx / HD, d = idx % HD;
        Q[i * HDP + d] = ldf(q + (long)rows[q0 + i] * p.ldq + p.qcol + h * HD + d);
      }
      __syncthreads();
      scores_softmax<HD>(ly, sm, n, scale);
      if (p.drop.p > 0.f) {   // SDPA dropout_p: P <- P * keep / (1 - p), no renormalisation
        const DropState ds = drop_state(p.drop);
        for (int idx = tid; idx < n * ly.NKV; idx += NT) {
          const int i = idx / ly.NKV, j = idx % ly.NKV;
          P[i * ly.SP + j] *= drop_keep1(ds, drop_id(p, rows[q0 + i], h, j));
        }
        __syncthreads();
      }
      for (int idx = tid; idx < n * HD; idx += NT) {
        const int i = idx / HD, d = idx % HD;
        float a = 0.f;
        for (int j = 0; j < ly.NKV; ++j) a = fmaf(P[i * ly.SP + j], Vf[j * HDP + d], a);
        stf(out + (long)rows[q0 + i] * p.ldo + h * HD + d, a);
      }
      __syncthreads();
    }
  }
}

template <typename T, int HD>
__global__ void __launch_bounds__(NT) attn_bwd_kernel(AttnP p, Lay ly, int ntask) {
  QV_PDL_ENTRY();
  extern __shared__ float sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int HDP = ly.HDP, NKV = ly.NKV, SP = ly.SP, D = p.H * HD;
  int* rows = reinterpret_cast<int*>(sm + ly.oRow);
  if (p.mode != 2 && ly.e_smem) {
    for (int idx = tid; idx < ly.L * p.klin; idx += NT) {
      sm[ly.oEk + idx] = p.Ek[idx]; sm[ly.oEv + idx] = p.Ev[idx];
      sm[ly.odEk + idx] = 0.f; sm[ly.odEv + idx] = 0.f;
    }
  }
  for (int idx = tid; idx < p.kb * D; idx += NT) { sm[ly.odbk + idx] = 0.f; sm[ly.odbv + idx] = 0.f; }
  const float scale = rsqrtf((float)HD);
  const T* q = static_cast<const T*>(p.q);
  const T* dout = static_cast<const T*>(p.dout);
  T* dq = static_cast<T*>(p.dq);
  T* dkv = static_cast<T*>(p.dkv);
  float *Q = sm + ly.oQ, *dO = sm + ly.odO, *P = sm + ly.oP, *dS = sm + ly.odS;
  float *Kf = sm + ly.oKf, *Vf = sm + ly.oVf, *dKf = sm + ly.odKf, *dVf = sm + ly.odVf;
  for (int task = blockIdx.x; task < ntask; task += gridDim.x) {
    const int w = task / p.H, h = task % p.H;
    __syncthreads();
    task_rows(p, ly, rows, w);
    __syncthreads();
    build_kv<T, HD>(p, ly, sm, rows, h);
    for (int idx = tid; idx < NKV * HDP; idx += NT) { dKf[idx] = 0.f; dVf[idx] = 0.f; }
    for (int q0 = 0; q0 < ly.nq; q0 += ly.QC) {
      const int n = min(ly.QC, ly.nq - q0);
      for (int idx = tid; idx < n * HD; idx += NT) {
        const int i = idx / HD, d = idx % HD;
        const long r = rows[q0 + i];
        Q[i * HDP + d] = ldf(q + r * p.ldq + p.qcol + h * HD + d);
        dO[i * HDP + d] = ldf(dout + r * p.lddo + h * HD + d);
      }
      __syncthreads();
      scores_softmax<HD>(ly, sm, n, scale);
      // dS = P * (dO Vf^T - rowsum(P * dO Vf^T)) * scale : one warp per query row
      const bool drop = p.drop.p > 0.f;
      DropState dst{};
      if (drop) dst = drop_state(p.drop);
      for (int i = warp; i < n; i += NW) {
        float s = 0.f;
        for (int j = lane; j < NKV; j += 32) {
          float a = 0.f;
#pragma unroll
          for (int d = 0; d < HD; ++d) a = fmaf(dO[i * HDP + d], Vf[j * HDP + d], a);
          if (drop) a *= drop_keep1(dst, drop_id(p, rows[q0 + i], h, j));   // d(P_dropped) -> dP
          dS[i * SP + j] = a;
          s = fmaf(a, P[i * SP + j], s);
        }
        s = warp_sum(s);
        for (int j = lane; j < NKV; j += 32) {
          const float pv = P[i * SP + j];
          dS[i * SP + j] = pv * (dS[i * SP + j] - s) * scale;
          if (drop) P[i * SP + j] = pv * drop_keep1(dst, drop_id(p, rows[q0 + i], h, j));   // dVf below needs the dropped P
        }
      }
      __syncthreads();
      // dQ = dS Kf ; dVf += P^T dO ; dKf += dS^T Q
      for (int idx = tid; idx < n * HD; idx += NT) {
        const int i = idx / HD, d = idx % HD;
        float a = 0.f;
        for (int j = 0; j < NKV; ++j) a = fmaf(dS[i * SP + j], Kf[j * HDP + d], a);
        stf(dq + (long)rows[q0 + i] * p.lddq + p.dqcol + h * HD + d, a);
      }
      for (int idx = tid; idx < NKV * HD; idx += NT) {
        const int j = idx / HD, d = idx % HD;
        float av = 0.f, ak = 0.f;
        for (int i = 0; i < n; ++i) {
          av = fmaf(P[i * SP + j], dO[i * HDP + d], av);
          ak = fmaf(dS[i * SP + j], Q[i * HDP + d], ak);
        }
        dVf[j * HDP + d] += av;
        dKf[j * HDP + d] += ak;
      }
      __syncthreads();
    }
    // ---- distribute dKf / dVf
    const int boff = (p.mode == 2) ? 0 : p.klin;
    for (int idx = tid; idx < p.kb * HD; idx += NT) {
      const int j = idx / HD, d = idx % HD;
      sm[ly.odbk + j * D + h * HD + d] += dKf[(boff + j) * HDP + d];
      sm[ly.odbv + j * D + h * HD + d] += dVf[(boff + j) * HDP + d];
    }
    if (p.mode != 2) {
      const float *Ek = ly.e_smem ? sm + ly.oEk : p.Ek, *Ev = ly.e_smem ? sm + ly.oEv : p.Ev, *Ks = sm + ly.oKs, *Vs = sm + ly.oVs;
      float *dEk = sm + ly.odEk, *dEv = sm + ly.odEv;
      // dKsrc[l, d] = sum_j E[l, j] dK'[j, d]
      for (int idx = tid; idx < ly.L * HD; idx += NT) {
        const int l = idx / HD, d = idx % HD;
        float ak = 0.f, av = 0.f;
        for (int j = 0; j < p.klin; ++j) {
          ak = fmaf(Ek[l * p.klin + j], dKf[j * HDP + d], ak);
          av = fmaf(Ev[l * p.klin + j], dVf[j * HDP + d], av);
        }
        const long r = rows[ly.nq + l];
        stf(dkv + r * p.lddkv + p.dkcol + h * HD + d, ak);
        stf(dkv + r * p.lddkv + p.dvcol + h * HD + d, av);
      }
      if (p.mode == 1) {  // pooled rows beyond the Linformer length were truncated away (H:510-512): zero grad
        for (int idx = tid; idx < (p.NM - ly.L) * HD; idx += NT) {
          const int l = ly.L + idx / HD, d = idx % HD;
          const long r = rows[ly.nq + l];
          stf(dkv + r * p.lddkv + p.dkcol + h * HD + d, 0.f);
          stf(dkv + r * p.lddkv + p.dvcol + h * HD + d, 0.f);
        }
      }
      // dE[l, j] += sum_d Ksrc[l, d] dK'[j, d] : warp per l, lanes over j
      for (int l = warp; l < ly.L; l += NW) {
        for (int j = lane; j < p.klin; j += 32) {
          float ak = 0.f, av = 0.f;
#pragma unroll
          for (int d = 0; d < HD; ++d) {
            ak = fmaf(Ks[l * HDP + d], dKf[j * HDP + d], ak);
            av = fmaf(Vs[l * HDP + d], dVf[j * HDP + d], av);
          }
          if (ly.e_smem) {
            dEk[l * p.klin + j] += ak;
            dEv[l * p.klin + j] += av;
          } else {   // large Linformer matrices (128 x 64 at 224 / 16): accumulate straight into the parameter gradient
            atomicAdd(p.dEk + l * p.klin + j, ak);
            atomicAdd(p.dEv + l * p.klin + j, av);
          }
        }
      }
    }
  }
  __syncthreads();
  // ---- flush the batch-reduced accumulators
  if (p.mode != 2 && ly.e_smem) {
    for (int idx = tid; idx < ly.L * p.klin; idx += NT) {
      atomicAdd(p.dEk + idx, sm[ly.odEk + idx]);
      atomicAdd(p.dEv + idx, sm[ly.odEv + idx]);
    }
  }
  for (int idx = tid; idx < p.kb * D; idx += NT) {
    atomicAdd(p.dbank_k + idx, sm[ly.odbk + idx]);
    atomicAdd(p.dbank_v + idx, sm[ly.odbv + idx]);
  }
}

Lay make_layout(const AttnP& p, bool bwd) {
  Lay ly{};
  ly.HDP = p.hd + 1;
  ly.nq = (p.mode == 0) ? p.ws * p.ws : p.Nt;
  ly.L = (p.mode == 2) ? 0 : p.L;
  ly.NKV = (p.mode == 2) ? p.kb : p.klin + p.kb;
  ly.SP = ly.NKV + 1;
  ly.QC = ly.nq < 32 ? ly.nq : 32;
  ly.e_smem = (ly.L * p.klin * (bwd ? 4 : 2)) <= 16384 ? 1 : 0;     // <= 64 KB of E / dE in shared memory
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
  ly.oEk = take(ly.e_smem ? ly.L * p.klin : 0); ly.oEv = take(ly.e_smem ? ly.L * p.klin : 0);
  ly.odEk = take(bwd && ly.e_smem ? ly.L * p.klin : 0); ly.odEv = take(bwd && ly.e_smem ? ly.L * p.klin : 0);
  ly.oKs = take(ly.L * ly.HDP); ly.oVs = take(ly.L * ly.HDP);
  ly.oKf = take(ly.NKV * ly.HDP); ly.oVf = take(ly.NKV * ly.HDP);
  ly.odKf = take(bwd ? ly.NKV * ly.HDP : 0); ly.odVf = take(bwd ? ly.NKV * ly.HDP : 0);
  ly.oQ = take(ly.QC * ly.HDP); ly.odO = take(bwd ? ly.QC * ly.HDP : 0);
  ly.oP = take(ly.QC * ly.SP); ly.odS = take(bwd ? ly.QC * ly.SP : 0);
  ly.odbk = take(bwd ? p.kb * p.H * p.hd : 0); ly.odbv = take(bwd ? p.kb * p.H * p.hd : 0);
  ly.oRow = take(ly.nq + (p.mode == 1 ? p.NM : ly.L));
  ly.total = o;
  return ly;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  QV_CHECK(bytes <= 227 * 1024, "attention kernel needs %zu B of shared memory (> 227 KB): config not supported", bytes);
  if (bytes > 48 * 1024) QV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

template <typename T, bool BWD>
int launch_attn(cudaStream_t s, const AttnP& p) {
  QV_CHECK(p.hd == 48, "attention kernels are instantiated for head_dim 48 (got %d)", p.hd);
  const Lay ly = make_layout(p, BWD);
  const int nwin = (p.mode == 0) ? p.B * (p.side / p.ws) * (p.side / p.ws) : p.B;
  const int ntask = nwin * p.H;
  if (ntask <= 0) return 0;
  const size_t smem = (size_t)ly.total * sizeof(float);
  const int occ = max(1, min(8, (int)(220 * 1024 / (smem + 1024))));
  const int grid = min(ntask, qv_num_sms() * occ);
  if (BWD) {
    QV_TRY(set_smem(attn_bwd_kernel<T, 48>, smem));
    qv_launch(attn_bwd_kernel<T, 48>, grid, NT, smem, s, p, ly, ntask);
  } else {
    QV_TRY(set_smem(attn_fwd_kernel<T, 48>, smem));
    qv_launch(attn_fwd_kernel<T, 48>, grid, NT, smem, s, p, ly, ntask);
  }
  QV_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// Cross attention (mode 2) has no per-image keys (K / V are the batch-invariant bank projections), so an image with
// Nt = 16 m query tokens is the same problem as m images of 16 tokens: lets the 64-token blocks (QAViTv2, TinyImageNet)
// use the 16-query tensor-core kernel.
static AttnP as_tiles16(const AttnP& p) {
  AttnP q = p;
  if (p.mode == 2 && p.Nt > 16 && p.Nt % 16 == 0) {
    q.B = p.B * (p.Nt / 16);
    q.Nt = 16;
  }
  return q;
}

int attn_fwd(cudaStream_t s, int dt, const AttnP& p0) {
  const AttnP p = (dt == QV_BF16) ? as_tiles16(p0) : p0;
  if (dt == QV_BF16 && attn_mma_ok(p)) return attn_mma_fwd(s, p);   // tensor-core path (attn_mma.cu)
  if (dt == QV_BF16 && p.wsp && attn_msda64_ok(p)) return attn_msda64_fwd(s, p, p.wsp);
  return dt == QV_F32 ? launch_attn<float, false>(s, p) : launch_attn<bf16, false>(s, p);
}
int attn_bwd(cudaStream_t s, int dt, const AttnP& p0) {
  const AttnP p = (dt == QV_BF16) ? as_tiles16(p0) : p0;
  if (dt == QV_BF16 && attn_mma_ok(p)) return attn_mma_bwd(s, p);
  if (dt == QV_BF16 && p.wsp && attn_msda64_ok(p)) return attn_msda64_bwd(s, p, p.wsp);
  return dt == QV_F32 ? launch_attn<float, true>(s, p) : launch_attn<bf16, true>(s, p);
}
