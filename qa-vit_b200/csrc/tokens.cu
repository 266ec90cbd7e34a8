// TokenLearner / TokenUpMix (H:971-1031) for 16 learned tokens: register-blocked, one thread per (image, channel)
// for the token-mixing products (each thread streams its own channel column straight from global memory, the tiny
// [N x 16] mixing matrix is broadcast from shared memory as float4), and a (token, 4-slot) thread tiling for the
// products that reduce over channels.  The generic kernels in misc.cu remain for other M.
#include "kernels.h"

namespace {

constexpr int M16 = 16;

template <typename K>
int opt_in(K kernel, size_t bytes) {
  QV_CHECK(bytes <= 227 * 1024, "kernel needs %zu B of shared memory (> 227 KB): config not supported", bytes);
  if (bytes > 48 * 1024) QV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

// ---------------------------------------------------------------------------------------------- TokenLearner forward
// S = softmax over tokens of logits[b, n, m];  xc[b, m, c] = sum_n S[n, m] x[b, n, c].   One CTA per image.
template <typename T, int CT>
__global__ void __launch_bounds__(192) tl16_fwd_kernel(const float* __restrict__ x, const T* __restrict__ logits, int B, int N,
                                                       int Crt, float* __restrict__ S, float* __restrict__ xc) {
  QV_PDL_ENTRY();
  const int C = CT ? CT : Crt;
  extern __shared__ __align__(16) float sS[];   // [N][16]
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int idx = tid; idx < N * M16; idx += blockDim.x) sS[idx] = ldf(logits + (long)b * N * M16 + idx);
  __syncthreads();
  {  // column softmax: 12 threads per slot share the N tokens (192 threads = 16 slots x 12)
    const int m = tid / 12, part = tid % 12;
    float mx = -INFINITY;
    for (int n = part; n < N; n += 12) mx = fmaxf(mx, sS[n * M16 + m]);
    // reduce over the 12 threads of the slot through shared memory
    __shared__ float red[16][12];
    red[m][part] = mx;
    __syncthreads();
    mx = red[m][0];
#pragma unroll
    for (int k = 1; k < 12; ++k) mx = fmaxf(mx, red[m][k]);
    __syncthreads();
    float z = 0.f;
    for (int n = part; n < N; n += 12) { const float e = __expf(sS[n * M16 + m] - mx); sS[n * M16 + m] = e; z += e; }
    red[m][part] = z;
    __syncthreads();
    z = 0.f;
#pragma unroll
    for (int k = 0; k < 12; ++k) z += red[m][k];
    z = 1.f / z;
    for (int n = part; n < N; n += 12) sS[n * M16 + m] *= z;
  }
  __syncthreads();
  for (int idx = tid; idx < N * M16; idx += blockDim.x) S[(long)b * N * M16 + idx] = sS[idx];
  for (int c = tid; c < C; c += blockDim.x) {
    float acc[M16];
#pragma unroll
    for (int m = 0; m < M16; ++m) acc[m] = 0.f;
    const float* xp = x + (long)b * N * C + c;
    for (int n0 = 0; n0 < N; n0 += 8) {
      float xv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) xv[k] = (n0 + k < N) ? xp[(long)(n0 + k) * C] : 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (n0 + k < N) {
          const float4* sr = reinterpret_cast<const float4*>(sS + (n0 + k) * M16);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 w = sr[q];
            acc[4 * q] = fmaf(w.x, xv[k], acc[4 * q]); acc[4 * q + 1] = fmaf(w.y, xv[k], acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(w.z, xv[k], acc[4 * q + 2]); acc[4 * q + 3] = fmaf(w.w, xv[k], acc[4 * q + 3]);
          }
        }
      }
    }
#pragma unroll
    for (int m = 0; m < M16; ++m) xc[((long)b * M16 + m) * C + c] = acc[m];
  }
}

// ---------------------------------------------------------------------------------------------- TokenLearner backward
// dS[n, m] = x[n, :] . dxc[m, :];  dlogits = S (dS - colsum_n(S dS));  dx[n, :] = sum_m S[n, m] dxc[m, :]
template <typename T, int CT>
__global__ void __launch_bounds__(256) tl16_bwd_kernel(const float* __restrict__ x, const float* __restrict__ S,
                                                       const float* __restrict__ dxc, int B, int N, int Crt,
                                                       T* __restrict__ dlogits, float* __restrict__ dx) {
  QV_PDL_ENTRY();
  const int C = CT ? CT : Crt;
  extern __shared__ __align__(16) float sm[];
  constexpr int RC = 64;                          // tokens staged per pass
  const int CP = C + 1;
  float* sS = sm;                                 // [N][16]
  float* sdS = sS + N * M16;                      // [N][16]
  float* sDT = sdS + N * M16;                     // [C][16]   dxc transposed
  float* st = sDT + C * M16;                      // [16]
  float* sX = st + M16;                           // [RC][C + 1]
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int idx = tid; idx < N * M16; idx += blockDim.x) sS[idx] = S[(long)b * N * M16 + idx];
  for (int idx = tid; idx < M16 * C; idx += blockDim.x) sDT[(idx % C) * M16 + idx / C] = dxc[(long)b * M16 * C + idx];
  const int nl = tid >> 2, mq = tid & 3;          // thread tile: token nl (0..63), slots 4 mq .. 4 mq + 3
  for (int n0 = 0; n0 < N; n0 += RC) {
    const int nr = min(RC, N - n0);
    __syncthreads();
    for (int idx = tid; idx < nr * (C / 4); idx += blockDim.x) {
      const int r = idx / (C / 4), c4 = idx % (C / 4);
      const float4 v = *reinterpret_cast<const float4*>(x + ((long)b * N + n0 + r) * C + 4 * c4);
      float* d = sX + r * CP + 4 * c4;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    __syncthreads();
    if (nl < nr) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      const float* xr = sX + nl * CP;
#pragma unroll 8
      for (int c = 0; c < C; ++c) {
        const float xv = xr[c];
        const float4 dv = *reinterpret_cast<const float4*>(sDT + c * M16 + 4 * mq);
        a0 = fmaf(xv, dv.x, a0); a1 = fmaf(xv, dv.y, a1); a2 = fmaf(xv, dv.z, a2); a3 = fmaf(xv, dv.w, a3);
      }
      *reinterpret_cast<float4*>(sdS + (n0 + nl) * M16 + 4 * mq) = make_float4(a0, a1, a2, a3);
    }
  }
  __syncthreads();
  if (tid < M16) {
    float t = 0.f;
    for (int n = 0; n < N; ++n) t = fmaf(sS[n * M16 + tid], sdS[n * M16 + tid], t);
    st[tid] = t;
  }
  __syncthreads();
  for (int idx = tid; idx < N * M16; idx += blockDim.x) stf(dlogits + (long)b * N * M16 + idx, sS[idx] * (sdS[idx] - st[idx % M16]));
  for (int c = tid; c < C; c += blockDim.x) {
    float dv[M16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 w = *reinterpret_cast<const float4*>(sDT + c * M16 + 4 * q);
      dv[4 * q] = w.x; dv[4 * q + 1] = w.y; dv[4 * q + 2] = w.z; dv[4 * q + 3] = w.w;
    }
    float* op = dx + (long)b * N * C + c;
    for (int n = 0; n < N; ++n) {
      const float4* sr = reinterpret_cast<const float4*>(sS + n * M16);
      float a = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 w = sr[q];
        a = fmaf(w.x, dv[4 * q], a); a = fmaf(w.y, dv[4 * q + 1], a); a = fmaf(w.z, dv[4 * q + 2], a); a = fmaf(w.w, dv[4 * q + 3], a);
      }
      op[(long)n * C] = a;
    }
  }
}

// ---------------------------------------------------------------------------------------------- TokenUpMix forward
// up[b, n, c] = sum_m W[n, m] xc[b, m, c] + bias[n]
__global__ void __launch_bounds__(192) up16_fwd_kernel(const float* __restrict__ xc, int B, int N, int C,
                                                       const float* __restrict__ W, const float* __restrict__ bias,
                                                       float* __restrict__ up) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) float sW[];    // [N][16] then bias [N]
  float* sb = sW + N * M16;
  for (int idx = threadIdx.x; idx < N * M16; idx += blockDim.x) sW[idx] = W[idx];
  for (int idx = threadIdx.x; idx < N; idx += blockDim.x) sb[idx] = bias[idx];
  __syncthreads();
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float xr[M16];
#pragma unroll
      for (int m = 0; m < M16; ++m) xr[m] = xc[((long)b * M16 + m) * C + c];
      float* op = up + (long)b * N * C + c;
#pragma unroll 4
      for (int n = 0; n < N; ++n) {
        const float4* wr = reinterpret_cast<const float4*>(sW + n * M16);
        float a = sb[n];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 w = wr[q];
          a = fmaf(w.x, xr[4 * q], a); a = fmaf(w.y, xr[4 * q + 1], a); a = fmaf(w.z, xr[4 * q + 2], a); a = fmaf(w.w, xr[4 * q + 3], a);
        }
        op[(long)n * C] = a;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- TokenUpMix backward
// dxc[b, m, c] = sum_n W[n, m] dup[b, n, c]
__global__ void __launch_bounds__(192) up16_dx_kernel(const float* __restrict__ dup, int B, int N, int C,
                                                      const float* __restrict__ W, float* __restrict__ dxc) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) float sW[];    // [N][16]
  for (int idx = threadIdx.x; idx < N * M16; idx += blockDim.x) sW[idx] = W[idx];
  __syncthreads();
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float acc[M16];
#pragma unroll
      for (int m = 0; m < M16; ++m) acc[m] = 0.f;
      const float* gp = dup + (long)b * N * C + c;
      for (int n0 = 0; n0 < N; n0 += 8) {
        float gv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) gv[k] = (n0 + k < N) ? gp[(long)(n0 + k) * C] : 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (n0 + k < N) {
            const float4* wr = reinterpret_cast<const float4*>(sW + (n0 + k) * M16);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 w = wr[q];
              acc[4 * q] = fmaf(w.x, gv[k], acc[4 * q]); acc[4 * q + 1] = fmaf(w.y, gv[k], acc[4 * q + 1]);
              acc[4 * q + 2] = fmaf(w.z, gv[k], acc[4 * q + 2]); acc[4 * q + 3] = fmaf(w.w, gv[k], acc[4 * q + 3]);
            }
          }
        }
      }
#pragma unroll
      for (int m = 0; m < M16; ++m) dxc[((long)b * M16 + m) * C + c] = acc[m];
    }
  }
}

// dW[n, m] += sum_{b, c} dup[b, n, c] xc[b, m, c];  dbias[n] += sum_{b, c} dup[b, n, c].
// Thread tile (token, 4 slots); 64 tokens staged per pass; accumulators in shared memory across the CTA's images.
template <int CT>
__global__ void __launch_bounds__(256) up16_dw_kernel(const float* __restrict__ xc, const float* __restrict__ dup, int B, int N,
                                                      int Crt, float* __restrict__ dW, float* __restrict__ dbias) {
  QV_PDL_ENTRY();
  const int C = CT ? CT : Crt;
  extern __shared__ __align__(16) float sm[];
  constexpr int RC = 64;
  const int CP = C + 1;
  float* sdW = sm;                    // [N][16]
  float* sdb = sdW + N * M16;         // [N]
  float* sXT = sdb + ((N + 3) & ~3);  // [C][16]   xc transposed
  float* sG = sXT + C * M16;          // [RC][C + 1]
  const int tid = threadIdx.x, nl = tid >> 2, mq = tid & 3;
  for (int idx = tid; idx < N * M16; idx += blockDim.x) sdW[idx] = 0.f;
  for (int idx = tid; idx < N; idx += blockDim.x) sdb[idx] = 0.f;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    for (int idx = tid; idx < M16 * C; idx += blockDim.x) sXT[(idx % C) * M16 + idx / C] = xc[(long)b * M16 * C + idx];
    for (int n0 = 0; n0 < N; n0 += RC) {
      const int nr = min(RC, N - n0);
      __syncthreads();
      for (int idx = tid; idx < nr * (C / 4); idx += blockDim.x) {
        const int r = idx / (C / 4), c4 = idx % (C / 4);
        const float4 v = *reinterpret_cast<const float4*>(dup + ((long)b * N + n0 + r) * C + 4 * c4);
        float* d = sG + r * CP + 4 * c4;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      }
      __syncthreads();
      if (nl < nr) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, sb = 0.f;
        const float* gr = sG + nl * CP;
#pragma unroll 8
        for (int c = 0; c < C; ++c) {
          const float gv = gr[c];
          const float4 xv = *reinterpret_cast<const float4*>(sXT + c * M16 + 4 * mq);
          a0 = fmaf(gv, xv.x, a0); a1 = fmaf(gv, xv.y, a1); a2 = fmaf(gv, xv.z, a2); a3 = fmaf(gv, xv.w, a3);
          sb += gv;
        }
        float4* d = reinterpret_cast<float4*>(sdW + (n0 + nl) * M16 + 4 * mq);
        float4 o = *d;
        o.x += a0; o.y += a1; o.z += a2; o.w += a3;
        *d = o;
        if (mq == 0) sdb[n0 + nl] += sb;
      }
    }
  }
  __syncthreads();
  for (int idx = tid; idx < N * M16; idx += blockDim.x) atomicAdd(dW + idx, sdW[idx]);
  for (int idx = tid; idx < N; idx += blockDim.x) atomicAdd(dbias + idx, sdb[idx]);
}

}  // namespace

bool tokens16_ok(int M, int C) { return M == 16 && C % 4 == 0 && C <= 1024; }

int tl16_fwd(cudaStream_t s, int dt, const float* x, const void* logits, int B, int N, int C, float* S, float* xc) {
  const size_t smem = (size_t)N * M16 * sizeof(float);
#define TLF(T, CT) do { QV_TRY(opt_in(tl16_fwd_kernel<T, CT>, smem)); qv_launch(tl16_fwd_kernel<T, CT>, B, 192, smem, s, x, (const T*)logits, B, N, C, S, xc); } while (0)
  if (dt == QV_F32) { if (C == 192) TLF(float, 192); else TLF(float, 0); }
  else { if (C == 192) TLF(bf16, 192); else TLF(bf16, 0); }
#undef TLF
  QV_LAUNCH_CHECK();
  return 0;
}
int tl16_bwd(cudaStream_t s, int dt, const float* x, const float* S, const float* dxc, int B, int N, int C, void* dlogits, float* dx) {
  const size_t smem = (size_t)(2 * N * M16 + C * M16 + M16 + 64 * (C + 1)) * sizeof(float);
#define TLB(T, CT) do { QV_TRY(opt_in(tl16_bwd_kernel<T, CT>, smem)); qv_launch(tl16_bwd_kernel<T, CT>, B, 256, smem, s, x, S, dxc, B, N, C, (T*)dlogits, dx); } while (0)
  if (dt == QV_F32) { if (C == 192) TLB(float, 192); else TLB(float, 0); }
  else { if (C == 192) TLB(bf16, 192); else TLB(bf16, 0); }
#undef TLB
  QV_LAUNCH_CHECK();
  return 0;
}
int up16_fwd(cudaStream_t s, const float* xc, int B, int N, int C, const float* W, const float* bias, float* up) {
  const size_t smem = (size_t)(N * M16 + N) * sizeof(float);
  QV_TRY(opt_in(up16_fwd_kernel, smem));
  qv_launch(up16_fwd_kernel, min(B, qv_num_sms() * 16), 192, smem, s, xc, B, N, C, W, bias, up);
  QV_LAUNCH_CHECK();
  return 0;
}
int up16_bwd(cudaStream_t s, const float* xc, const float* dup, int B, int N, int C, const float* W, float* dxc, float* dW,
             float* dbias) {
  size_t smem = (size_t)N * M16 * sizeof(float);
  QV_TRY(opt_in(up16_dx_kernel, smem));
  qv_launch(up16_dx_kernel, min(B, qv_num_sms() * 16), 192, smem, s, dup, B, N, C, W, dxc);
  QV_LAUNCH_CHECK();
  smem = (size_t)(N * M16 + ((N + 3) & ~3) + C * M16 + 64 * (C + 1)) * sizeof(float);
  const int occ = max(1, (int)(220 * 1024 / (smem + 1024)));
  const int grid = min(B, qv_num_sms() * min(occ, 4));
  if (C == 192) { QV_TRY(opt_in(up16_dw_kernel<192>, smem)); qv_launch(up16_dw_kernel<192>, grid, 256, smem, s, xc, dup, B, N, C, dW, dbias); }
  else { QV_TRY(opt_in(up16_dw_kernel<0>, smem)); qv_launch(up16_dw_kernel<0>, grid, 256, smem, s, xc, dup, B, N, C, dW, dbias); }
  QV_LAUNCH_CHECK();
  return 0;
}
