// Patch embedding (im2col-free), classifier head, label-smoothed cross entropy, gradient clipping and the
// fused AdamW step.  Warp-shuffle reductions, coalesced/vectorised streams; all fp32.
#include "kernels.h"

// =============================================================================== patch embed (H:1129-1138, :1250)
namespace {
constexpr int PE_KC = 48;  // contraction chunk staged in shared memory
// One CTA per (image, chunk of <= 64 patches).  stride == kernel, so patch row (py, px) is the [Cin, p, p] box at
// (py*p, px*p): read straight from the image, no im2col buffer.  pre = conv (saved for backward), out = LN(pre) + pos.
constexpr int PE_NC = 64;
__global__ void __launch_bounds__(192) patch_embed_fwd_kernel(const float* __restrict__ img, int B, int Cin, int S, int p,
                                                              int d, const float* __restrict__ W,
                                                              const float* __restrict__ bias,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta,
                                                              const float* __restrict__ pos, float* __restrict__ pre,
                                                              float* __restrict__ stats, float* __restrict__ out) {
  QV_PDL_ENTRY();
  extern __shared__ float sm[];
  const int n_side = S / p, N = n_side * n_side, K = Cin * p * p;
  const int chunks = (N + PE_NC - 1) / PE_NC;
  float* sP = sm;                      // [PE_NC][PE_KC]  patch chunk
  float* sW = sP + PE_NC * PE_KC;      // [d][PE_KC + 1]  weight chunk
  float* sO = sW + d * (PE_KC + 1);    // [PE_NC][d]      conv output
  const int tid = threadIdx.x;
  for (int task = blockIdx.x; task < B * chunks; task += gridDim.x) {
    const int b = task / chunks, n0 = (task % chunks) * PE_NC, nn = min(PE_NC, N - n0);
    for (int idx = tid; idx < nn * d; idx += blockDim.x) sO[idx] = bias[idx % d];
    for (int k0 = 0; k0 < K; k0 += PE_KC) {
      const int kc = min(PE_KC, K - k0);
      __syncthreads();
      for (int idx = tid; idx < nn * kc; idx += blockDim.x) {
        const int n = n0 + idx / kc, k = k0 + idx % kc;
        const int c = k / (p * p), r = (k / p) % p, q = k % p;
        sP[(idx / kc) * PE_KC + idx % kc] = img[(((long)b * Cin + c) * S + (n / n_side) * p + r) * S + (n % n_side) * p + q];
      }
      for (int idx = tid; idx < d * kc; idx += blockDim.x) sW[(idx / kc) * (PE_KC + 1) + idx % kc] = W[(long)(idx / kc) * K + k0 + idx % kc];
      __syncthreads();
      for (int o = tid; o < d; o += blockDim.x) {
        for (int n = 0; n < nn; ++n) {
          float a = sO[n * d + o];
          for (int k = 0; k < kc; ++k) a = fmaf(sP[n * PE_KC + k], sW[o * (PE_KC + 1) + k], a);
          sO[n * d + o] = a;
        }
      }
    }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    for (int n = warp; n < nn; n += nwarp) {
      float s = 0.f;
      for (int c = lane; c < d; c += 32) s += sO[n * d + c];
      const float mean = warp_sum(s) / d;
      float q = 0.f;
      for (int c = lane; c < d; c += 32) { const float t = sO[n * d + c] - mean; q += t * t; }
      const float rstd = rsqrtf(warp_sum(q) / d + 1e-5f);
      const long row = (long)b * N + n0 + n;
      if (lane == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
      for (int c = lane; c < d; c += 32) {
        const float v = sO[n * d + c];
        pre[row * d + c] = v;
        out[row * d + c] = (v - mean) * rstd * gamma[c] + beta[c] + (pos ? pos[(n0 + n) * d + c] : 0.f);
      }
    }
    __syncthreads();
  }
}
// dW[o, k] += sum_{b, n} dpre[b, n, o] * patch[b, n, k].  thread = output channel, 48 accumulators in registers.
__global__ void __launch_bounds__(192) patch_embed_dw_kernel(const float* __restrict__ img, const float* __restrict__ dpre,
                                                             int B, int Cin, int S, int p, int d,
                                                             float* __restrict__ dW) {
  QV_PDL_ENTRY();
  extern __shared__ float sP[];  // [PE_NC][PE_KC]
  const int n_side = S / p, N = n_side * n_side, K = Cin * p * p;
  const int chunks = (N + PE_NC - 1) / PE_NC;
  const int tid = threadIdx.x;
  for (int k0 = 0; k0 < K; k0 += PE_KC) {
    const int kc = min(PE_KC, K - k0);
    float acc[PE_KC];
#pragma unroll
    for (int k = 0; k < PE_KC; ++k) acc[k] = 0.f;
    for (int task = blockIdx.x; task < B * chunks; task += gridDim.x) {
      const int b = task / chunks, n0 = (task % chunks) * PE_NC, nn = min(PE_NC, N - n0);
      __syncthreads();
      for (int idx = tid; idx < nn * kc; idx += blockDim.x) {
        const int n = n0 + idx / kc, k = k0 + idx % kc;
        const int c = k / (p * p), r = (k / p) % p, q = k % p;
        sP[(idx / kc) * PE_KC + idx % kc] = img[(((long)b * Cin + c) * S + (n / n_side) * p + r) * S + (n % n_side) * p + q];
      }
      __syncthreads();
      if (tid < d) {
        for (int n = 0; n < nn; ++n) {
          const float g = dpre[((long)b * N + n0 + n) * d + tid];
#pragma unroll
          for (int k = 0; k < PE_KC; ++k) if (k < kc) acc[k] = fmaf(g, sP[n * PE_KC + k], acc[k]);
        }
      }
    }
    if (tid < d) {
#pragma unroll
      for (int k = 0; k < PE_KC; ++k) if (k < kc) atomicAdd(dW + (long)tid * K + k0 + k, acc[k]);
    }
  }
}
// ---- bf16 runs: the patch rows are gathered once into a bf16 [B*N, K] matrix (stride == kernel: every image element
// lands in exactly one row) and the convolution is the tcgen05 GEMM; LayerNorm + pos is a warp-per-row kernel.
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ img, int B, int Cin, int S, int p,
                                                       bf16* __restrict__ col) {
  QV_PDL_ENTRY();
  const int n_side = S / p, N = n_side * n_side, K = Cin * p * p;
  const long total = (long)B * N * K / 2;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const int k = (int)((2 * i) % K);
    const long row = (2 * i) / K;
    const int n = (int)(row % N);
    const long b = row / N;
    const int c = k / (p * p), r = (k / p) % p, q = k % p;      // p is even: (k, k + 1) share the image row
    const float2 v = *reinterpret_cast<const float2*>(img + ((b * Cin + c) * S + (n / n_side) * p + r) * S + (n % n_side) * p + q);
    *reinterpret_cast<__nv_bfloat162*>(col + 2 * i) = __floats2bfloat162_rn(v.x, v.y);
  }
}
template <int EPL>
__global__ void __launch_bounds__(256) ln_pos_fwd_kernel(const float* __restrict__ pre, long rows, int N, int C,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         const float* __restrict__ pos, float* __restrict__ stats,
                                                         float* __restrict__ out) {
  QV_PDL_ENTRY();
  const int lane = threadIdx.x & 31, c0 = lane * EPL;
  const bool act = c0 < C;
  const float invC = 1.f / (float)C;
  float gm[EPL], bt[EPL];
#pragma unroll
  for (int i = 0; i < EPL; ++i) gm[i] = bt[i] = 0.f;
  if (act) { load_vec<EPL>(gamma + c0, gm); load_vec<EPL>(beta + c0, bt); }
  for (long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += (long)gridDim.x * 8) {
    float v[EPL], ps[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) v[i] = ps[i] = 0.f;
    if (act) {
      load_vec<EPL>(pre + row * C + c0, v);
      if (pos) load_vec<EPL>(pos + (row % N) * C + c0, ps);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) s += v[i];
    const float mean = warp_sum(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) { const float t = act ? v[i] - mean : 0.f; q += t * t; }
    const float rstd = rsqrtf(warp_sum(q) * invC + 1e-5f);
    if (lane == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
#pragma unroll
    for (int i = 0; i < EPL; ++i) v[i] = (v[i] - mean) * rstd * gm[i] + bt[i] + ps[i];
    if (act) store_vec<EPL>(out + row * C + c0, v);
  }
}
int patchify(cudaStream_t s, const float* img, int B, int Cin, int S, int p, bf16* col) {
  const long total = (long)B * (S / p) * (S / p) * Cin * p * p / 2;
  qv_launch(patchify_kernel, (int)max(1L, min((total + 255) / 256, (long)qv_num_sms() * 8)), 256, 0, s, img, B, Cin, S, p, col);
  QV_LAUNCH_CHECK();
  return 0;
}
bool pe_tc_ok(int Cin, int S, int p, int d) {
  const int K = Cin * p * p;
  return p % 2 == 0 && S % p == 0 && K % 8 == 0 && d % 16 == 0 && d <= 256 && d % 4 == 0;
}
}  // namespace

size_t patch_embed_scratch_bytes(int B, int Cin, int S, int p, int d) {
  const size_t K = (size_t)Cin * p * p, rows = (size_t)B * (S / p) * (S / p);
  return ((rows * K * 2 + 255) & ~(size_t)255) + ((rows * d * 2 + 255) & ~(size_t)255) + 2 * (((size_t)d * K * 2 + 255) & ~(size_t)255);
}

int patch_embed_fwd(cudaStream_t s, int dt, const float* img, int B, int Cin, int S, int p, int d, const float* W,
                    const float* bias, const float* gamma, const float* beta, const float* pos, float* pre, float* stats,
                    float* out, void* scratch) {
  if (B <= 0) return 0;
  const int N = (S / p) * (S / p);
  if (dt == QV_BF16 && scratch && pe_tc_ok(Cin, S, p, d)) {
    const int K = Cin * p * p;
    const long rows = (long)B * N;
    uint8_t* sc = static_cast<uint8_t*>(scratch);
    bf16* col = reinterpret_cast<bf16*>(sc);
    bf16* wb = reinterpret_cast<bf16*>(sc + (((size_t)rows * K * 2 + 255) & ~(size_t)255) + (((size_t)rows * d * 2 + 255) & ~(size_t)255));
    QV_TRY(patchify(s, img, B, Cin, S, p, col));
    QV_TRY(convert_weight(s, W, d, K, wb, nullptr));
    GemmEpi e;
    e.bias = bias; e.C = pre; e.ldc = d; e.c_f32 = 1;
    QV_TRY(tc_gemm_nt(s, col, K, (int)rows, d, K, wb, e));
    const int grid = (int)max(1L, min((rows + 7) / 8, (long)qv_num_sms() * 8));
    if (d > 128) qv_launch(ln_pos_fwd_kernel<8>, grid, 256, 0, s, pre, rows, N, d, gamma, beta, pos, stats, out);
    else qv_launch(ln_pos_fwd_kernel<4>, grid, 256, 0, s, pre, rows, N, d, gamma, beta, pos, stats, out);
    QV_LAUNCH_CHECK();
    return 0;
  }
  const size_t smem = (size_t)(PE_NC * PE_KC + d * (PE_KC + 1) + PE_NC * d) * sizeof(float);
  QV_CHECK(smem <= 227 * 1024, "patch_embed: d=%d needs %zu B smem: not supported", d, smem);
  QV_CUDA(cudaFuncSetAttribute(patch_embed_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  qv_launch(patch_embed_fwd_kernel, min(B * cdiv(N, PE_NC), qv_num_sms() * 2), 192, smem, s, img, B, Cin, S, p, d, W, bias, gamma, beta, pos, pre, stats, out);
  QV_LAUNCH_CHECK();
  return 0;
}

int patch_embed_bwd(cudaStream_t s, int dt, const float* img, const float* dout, int B, int Cin, int S, int p, int d,
                    const float* pre, const float* stats, const float* gamma, float* dpre, float* dW, float* dbias,
                    float* dgamma, float* dbeta, float* dpos, void* scratch) {
  if (B <= 0) return 0;
  const int N = (S / p) * (S / p);
  // dpos[n, c] += sum_b dout[b, n, c]
  if (dpos) QV_TRY(colsum_accum(s, QV_F32, dout, N * d, B, N * d, dpos, nullptr));
  if (dt == QV_BF16 && scratch && pe_tc_ok(Cin, S, p, d)) {
    const int K = Cin * p * p;
    const long rows = (long)B * N;
    uint8_t* sc = static_cast<uint8_t*>(scratch);
    bf16* col = reinterpret_cast<bf16*>(sc);
    bf16* dpre_b = reinterpret_cast<bf16*>(sc + (((size_t)rows * K * 2 + 255) & ~(size_t)255));
    QV_TRY(ln_bwd(s, QV_F32, pre, d, QV_F32, dout, d, (int)rows, d, gamma, stats, 0, QV_BF16, dpre_b, nullptr, nullptr, dgamma, dbeta));
    QV_TRY(patchify(s, img, B, Cin, S, p, col));
    return gemm_tn(s, QV_BF16, dpre_b, d, col, K, (int)rows, d, K, dW, dbias, nullptr);
  }
  QV_CHECK(d <= 192, "patch_embed_bwd: d=%d > 192", d);
  QV_TRY(ln_bwd(s, QV_F32, pre, d, QV_F32, dout, d, B * N, d, gamma, stats, 0, QV_F32, nullptr, dpre, nullptr, dgamma, dbeta));
  QV_TRY(colsum_accum(s, QV_F32, dpre, d, B * N, d, dbias, nullptr));
  const size_t smem = (size_t)PE_NC * PE_KC * sizeof(float);
  qv_launch(patch_embed_dw_kernel, min(B * cdiv(N, PE_NC), qv_num_sms()), 192, smem, s, img, dpre, B, Cin, S, p, d, dW);
  QV_LAUNCH_CHECK();
  return 0;
}

// =============================================================================== head (H:1273-1275)
namespace {
// pooled[b, :] = mean_n LN(x[b, n, :]).  One CTA (8 warps) per image.
__global__ void __launch_bounds__(256) ln_mean_kernel(const float* __restrict__ x, int B, int N, int d,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      float* __restrict__ stats, float* __restrict__ pooled) {
  QV_PDL_ENTRY();
  __shared__ float acc[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 0.f;
    for (int n = warp; n < N; n += 8) {
      const long row = (long)b * N + n;
      float v[8], s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) { const int c = lane + 32 * i; v[i] = c < d ? x[row * d + c] : 0.f; s += v[i]; }
      const float mean = warp_sum(s) / d;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) { const int c = lane + 32 * i; if (c < d) { const float t = v[i] - mean; q += t * t; } }
      const float rstd = rsqrtf(warp_sum(q) / d + 1e-5f);
      if (lane == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
#pragma unroll
      for (int i = 0; i < 8; ++i) { const int c = lane + 32 * i; if (c < d) a[i] += (v[i] - mean) * rstd * gamma[c] + beta[c]; }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[warp][lane + 32 * i] = a[i];
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += acc[w][c];
      pooled[(long)b * d + c] = t / N;
    }
  }
}
// LN backward where every token row of image b receives dy = dpooled[b, :] / N.
__global__ void __launch_bounds__(256) ln_mean_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dpooled,
                                                          int B, int N, int d, const float* __restrict__ gamma,
                                                          const float* __restrict__ stats, float* __restrict__ dx,
                                                          float* __restrict__ dgamma, float* __restrict__ dbeta) {
  QV_PDL_ENTRY();
  __shared__ float red[2][8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float ag[8], ab[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) ag[i] = ab[i] = 0.f;
  const float invN = 1.f / N, invd = 1.f / d;
  for (long row = (long)blockIdx.x * 8 + warp; row < (long)B * N; row += (long)gridDim.x * 8) {
    const long b = row / N;
    const float mean = stats[2 * row], rstd = stats[2 * row + 1];
    float xh[8], g[8], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      xh[i] = g[i] = 0.f;
      if (c < d) {
        const float dyv = dpooled[b * d + c] * invN;
        xh[i] = (x[row * d + c] - mean) * rstd;
        g[i] = dyv * gamma[c];
        ag[i] += dyv * xh[i];
        ab[i] += dyv;
        s1 += g[i];
        s2 += g[i] * xh[i];
      }
    }
    const float c1 = warp_sum(s1) * invd, c2 = warp_sum(s2) * invd;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const int c = lane + 32 * i; if (c < d) dx[row * d + c] = rstd * (g[i] - c1 - xh[i] * c2); }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { red[0][warp][lane + 32 * i] = ag[i]; red[1][warp][lane + 32 * i] = ab[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float a = 0.f, bsum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { a += red[0][w][c]; bsum += red[1][w][c]; }
    atomicAdd(dgamma + c, a);
    atomicAdd(dbeta + c, bsum);
  }
}
}  // namespace

int head_fwd(cudaStream_t s, const float* x, int B, int N, int d, const float* gamma, const float* beta, const float* W,
             const float* bias, int ncls, float* stats, float* pooled, float* logits) {
  if (B <= 0) return 0;
  QV_CHECK(d <= 256, "head: d=%d > 256", d);
  qv_launch(ln_mean_kernel, min(B, qv_num_sms() * 4), 256, 0, s, x, B, N, d, gamma, beta, stats, pooled);
  QV_LAUNCH_CHECK();
  GemmEpi e;
  e.bias = bias; e.C = logits; e.ldc = ncls; e.c_f32 = 1;
  return simt_gemm_nt(s, QV_F32, pooled, d, B, ncls, d, W, e);
}

int head_bwd(cudaStream_t s, const float* x, const float* dlogits, int B, int N, int d, const float* gamma,
             const float* stats, const float* pooled, const float* W, int ncls, float* dpooled, float* dx,
             float* dgamma, float* dbeta, float* dW, float* dbias) {
  if (B <= 0) return 0;
  QV_TRY(simt_gemm_tn(s, QV_F32, dlogits, ncls, pooled, d, B, ncls, d, dW, dbias, nullptr));
  GemmEpi e;
  e.C = dpooled; e.ldc = d; e.c_f32 = 1;
  QV_TRY(simt_gemm_nn(s, QV_F32, dlogits, ncls, B, ncls, d, W, e));
  qv_launch(ln_mean_bwd_kernel, min(cdiv((long)B * N, 8), qv_num_sms() * 4), 256, 0, s, x, dpooled, B, N, d, gamma, stats, dx, dgamma, dbeta);
  QV_LAUNCH_CHECK();
  return 0;
}

// =============================================================================== cross entropy (H:1373, :1404-1408)
namespace {
// loss = mean_b [ lam * CE_ls(y_a) + (1 - lam) * CE_ls(y_b) ],  CE_ls = (1-eps) * nll + eps * mean_c(-log p_c)
// One warp per row writes the row's loss term into row_loss[row] (no atomics); ce_reduce_kernel then sums the rows in a
// fixed order, so the loss is bitwise reproducible run to run.  Labels outside [0, C) raise bit 0 of *err (torch asserts
// device-side there) and are clamped so that no read leaves the row.  lam_dev (optional) overrides lam: the mixup weight
// of a CUDA-graph-replayed step lives on the device.
__global__ void ce_kernel(const float* __restrict__ logits, const long long* __restrict__ ya,
                          const long long* __restrict__ yb, float lam, const float* __restrict__ lam_dev, int B, int C, float eps,
                          float* __restrict__ row_loss, float* __restrict__ dlogits, int* __restrict__ err) {
  QV_PDL_ENTRY();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  if (lam_dev) lam = *lam_dev;
  const float* l = logits + (long)row * C;
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, l[c]);
  m = warp_max(m);
  float z = 0.f, sl = 0.f;
  for (int c = lane; c < C; c += 32) { z += expf(l[c] - m); sl += l[c]; }
  z = warp_sum(z);
  sl = warp_sum(sl);
  const float lse = m + logf(z);
  long long a64 = ya[row], b64 = yb ? yb[row] : a64;
  if (a64 < 0 || a64 >= C || b64 < 0 || b64 >= C) {
    if (lane == 0 && err) atomicOr(err, 1);
    a64 = a64 < 0 ? 0 : (a64 >= C ? C - 1 : a64);
    b64 = b64 < 0 ? 0 : (b64 >= C ? C - 1 : b64);
  }
  const int a = (int)a64, b2 = (int)b64;
  const float wa = yb ? lam : 1.f, wb = yb ? 1.f - lam : 0.f;
  if (lane == 0) {
    const float nll = wa * (lse - l[a]) + wb * (lse - l[b2]);
    const float smooth = lse - sl / C;
    row_loss[row] = ((1.f - eps) * nll + eps * smooth) / B;
  }
  if (dlogits) {
    for (int c = lane; c < C; c += 32) {
      float t = eps / C;
      if (c == a) t += (1.f - eps) * wa;
      if (c == b2) t += (1.f - eps) * wb;
      dlogits[(long)row * C + c] = (expf(l[c] - lse) - t) / B;
    }
  }
}
// fixed-order sum of row_loss[0 .. B): thread t adds rows t, t + 1024, ... then a fixed shared-memory tree
__global__ void __launch_bounds__(1024) ce_reduce_kernel(const float* __restrict__ row_loss, int B, float* __restrict__ loss) {
  QV_PDL_ENTRY();
  __shared__ float red[1024];
  float a = 0.f;
  for (int i = threadIdx.x; i < B; i += 1024) a += row_loss[i];
  red[threadIdx.x] = a;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = red[0];
}
__global__ void scale_by_scalar_kernel(const float* __restrict__ x, const float* __restrict__ sc, long n, float* __restrict__ y) {
  QV_PDL_ENTRY();
  const float s = *sc;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) y[i] = x[i] * s;
}
}  // namespace

int ce_loss_fwd_bwd(cudaStream_t s, const float* logits, const long long* ya, const long long* yb, float lam, const float* lam_dev,
                    int B, int ncls, float smoothing, float* loss, float* dlogits, float* row_loss, int* err) {
  if (B <= 0) {
    QV_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), s));
    return 0;
  }
  QV_CHECK(row_loss, "cross_entropy: row_loss scratch (B floats) missing");
  qv_launch(ce_kernel, cdiv(B, 4), 128, 0, s, logits, ya, yb, lam, lam_dev, B, ncls, smoothing, row_loss, dlogits, err);
  QV_LAUNCH_CHECK();
  qv_launch(ce_reduce_kernel, 1, 1024, 0, s, row_loss, B, loss);
  QV_LAUNCH_CHECK();
  return 0;
}
int scale_by_scalar(cudaStream_t s, const float* x, const float* scalar_dev, long n, float* y) {
  if (n <= 0) return 0;
  qv_launch(scale_by_scalar_kernel, (int)min((long)qv_num_sms() * 8, (long)cdiv(n, 256)), 256, 0, s, x, scalar_dev, n, y);
  QV_LAUNCH_CHECK();
  return 0;
}
