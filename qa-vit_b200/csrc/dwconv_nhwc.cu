// Depthwise k x k convolutions (k = 3, 5, 7; stride 1, 'same' padding, bias) on channels-last feature maps
// [B, H, W, C] -- the ConvNeXt / LMFAdapter stencils of HQAViT's lateral path (H:722, 811-812; scope row f-1).
//
// One CTA per (image, 64-channel slab): the slab's H x W x 64 tile sits in shared memory (fp32), consecutive lanes own
// consecutive channels (coalesced global traffic, conflict-free shared loads).  A thread computes 8 outputs of one row
// at a time: per kernel row it loads the 8 + k - 1 inputs once into registers and issues 8 * k FMAs on them, so the
// stencil is FMA-bound (k * k FMAs per output, ~4 FMAs per shared load) instead of shared-load-bound.
// HBM-bound in the ideal: algorithmic bytes = B*H*W*C * (in + out).
#include "../../include/qavit_b200.h"
#include "kernels.h"

namespace {

constexpr int CS = 64;     // channels per CTA
constexpr int PG = 4;      // thread groups per channel (CTA = 256 threads)
constexpr int SEG = 8;     // outputs per thread task (one row segment)

template <typename T>
__device__ __forceinline__ void load_tile(float* xs, const T* __restrict__ x, long ldx, long row0, int HW, int c0, int C) {
  // [HW][CS] tile, two channels per thread
  for (int idx = threadIdx.x; idx < HW * (CS / 2); idx += CS * PG) {
    const int p = idx / (CS / 2), cc = (idx % (CS / 2)) * 2;
    float2 v = make_float2(0.f, 0.f);
    if (c0 + cc < C) v = ld2(x + (row0 + p) * ldx + c0 + cc);
    *reinterpret_cast<float2*>(xs + p * CS + cc) = v;
  }
}

// 8 x 8 maps: the next image's tile is fetched into registers while the current one is being computed (a CTA
// otherwise alternates between waiting for its 8 KB tile and computing on it).
constexpr int NPF = 8 * 8 * (CS / 2) / (CS * PG);   // float2 registers per thread for one 8 x 8 x 64 tile
template <typename T>
__device__ __forceinline__ void tile_prefetch(float2* pf, const T* __restrict__ x, long ldx, long row0, int c0, int C) {
#pragma unroll
  for (int i = 0; i < NPF; ++i) {
    const int idx = threadIdx.x + i * CS * PG, p = idx / (CS / 2), cc = (idx % (CS / 2)) * 2;
    pf[i] = (c0 + cc < C) ? ld2(x + (row0 + p) * ldx + c0 + cc) : make_float2(0.f, 0.f);
  }
}
__device__ __forceinline__ void tile_commit(float* xs, const float2* pf) {
#pragma unroll
  for (int i = 0; i < NPF; ++i) {
    const int idx = threadIdx.x + i * CS * PG, p = idx / (CS / 2), cc = (idx % (CS / 2)) * 2;
    *reinterpret_cast<float2*>(xs + p * CS + cc) = pf[i];
  }
}

// y = conv(x) (+ bias) (+ resid) (+ resid2); FLIP: correlate with the flipped kernel (= the input gradient).
// copy (optional): also writes the input tile to copy[row, c] (LMFAdapter's identity branch of the concat).
// WT > 0: the map width is the compile-time constant WT (8 / 16 / 24): a thread owns a whole output row, loads each
// input row once (WT shared loads with immediate offsets) and every out-of-range tap is pruned at compile time.
// WT == 0: generic width, 8-output segments with predicated loads.
template <typename T, int K, bool FLIP, int WT>
__global__ void __launch_bounds__(CS * PG) dw_fwd_kernel(const T* __restrict__ x, int ldx, int B, int H, int W, int C,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         const float* __restrict__ scale,
                                                         T* y, int ldy, const T* resid, int ldr,
                                                         const T* resid2, int ldr2, T* __restrict__ copy, int ldcp) {
  QV_PDL_ENTRY();
  extern __shared__ float xs[];                 // [H*W][CS]
  const int c0 = blockIdx.y * CS, cl = threadIdx.x % CS, pg = threadIdx.x / CS, c = c0 + cl;
  const int HW = H * W;
  const bool act = c < C;
  const float scl = (scale && act) ? scale[c] : 1.f;   // per-channel output scale folded into the taps and the bias
  float wr[K * K];                              // this channel's taps: loaded once, reused for every image of the CTA
#pragma unroll
  for (int t = 0; t < K * K; ++t) wr[t] = act ? w[c * K * K + (FLIP ? K * K - 1 - t : t)] * scl : 0.f;
  const float bs = (bias && act) ? bias[c] * scl : 0.f;
  constexpr bool kPrefetch = (WT == 8);
  float2 pf[NPF];
  if (kPrefetch && (int)blockIdx.x < B) tile_prefetch(pf, x, ldx, (long)blockIdx.x * HW, c0, C);
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const long row0 = (long)b * HW;
    __syncthreads();
    if (kPrefetch) tile_commit(xs, pf);
    else load_tile(xs, x, ldx, row0, HW, c0, C);
    __syncthreads();
    if (kPrefetch && b + (int)gridDim.x < B) tile_prefetch(pf, x, ldx, (long)(b + gridDim.x) * HW, c0, C);
    if (!act) continue;
    if (copy) {
      for (int p = pg; p < HW; p += PG) stf(copy + (row0 + p) * ldcp + c, xs[p * CS + cl]);
    }
    if (WT > 0) {
      for (int py = pg; py < H; py += PG) {
        float acc[WT > 0 ? WT : 1], rv[WT > 0 ? WT : 1];
        const long row = row0 + py * WT;
#pragma unroll
        for (int j = 0; j < WT; ++j) {            // residual loads first: their latency hides behind the stencil
          acc[j] = bs;
          rv[j] = resid ? ldf(resid + (row + j) * ldr + c) : 0.f;
          if (resid2) rv[j] += ldf(resid2 + (row + j) * ldr2 + c);
        }
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
          const int yy = py + ky - K / 2;
          if (yy < 0 || yy >= H) continue;
          const float* xrow = xs + (yy * WT) * CS + cl;
          float xr[WT > 0 ? WT : 1];
#pragma unroll
          for (int i = 0; i < WT; ++i) xr[i] = xrow[i * CS];
#pragma unroll
          for (int kx = 0; kx < K; ++kx)
#pragma unroll
            for (int j = 0; j < WT; ++j)
              if (j + kx - K / 2 >= 0 && j + kx - K / 2 < WT) acc[j] = fmaf(wr[ky * K + kx], xr[j + kx - K / 2], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < WT; ++j) stf(y + (row + j) * ldy + c, acc[j] + rv[j]);
      }
    } else {
      const int segs = (W + SEG - 1) / SEG;
      for (int task = pg; task < H * segs; task += PG) {
        const int py = task / segs, px0 = (task % segs) * SEG;
        float acc[SEG];
#pragma unroll
        for (int j = 0; j < SEG; ++j) acc[j] = bs;
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
          const int yy = py + ky - K / 2;
          if (yy < 0 || yy >= H) continue;
          float xr[SEG + K - 1];
#pragma unroll
          for (int i = 0; i < SEG + K - 1; ++i) {
            const int xx = px0 + i - K / 2;
            xr[i] = (xx >= 0 && xx < W) ? xs[(yy * W + xx) * CS + cl] : 0.f;
          }
#pragma unroll
          for (int kx = 0; kx < K; ++kx)
#pragma unroll
            for (int j = 0; j < SEG; ++j) acc[j] = fmaf(wr[ky * K + kx], xr[j + kx], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < SEG; ++j) {
          if (px0 + j >= W) break;
          const long row = row0 + py * W + px0 + j;
          float v = acc[j];
          if (resid) v += ldf(resid + row * ldr + c);
          if (resid2) v += ldf(resid2 + row * ldr2 + c);
          stf(y + row * ldy + c, v);
        }
      }
    }
  }
}

// dw[c, ky, kx] += sum_{b, p} dy[b, p, c] x[b, p + (ky, kx) - K/2, c];  dbias[c] += sum dy
template <typename T, int K, int WT>
__global__ void __launch_bounds__(CS * PG) dw_wgrad_kernel(const T* __restrict__ x, int ldx, const T* __restrict__ dy,
                                                           int lddy, int B, int H, int W, int C, float* __restrict__ dw,
                                                           float* __restrict__ dbias, const float* __restrict__ w,
                                                           const float* __restrict__ bias, const float* __restrict__ scale,
                                                           float* __restrict__ dscale) {
  QV_PDL_ENTRY();
  extern __shared__ float sm[];
  const int HW = H * W;
  float* xs = sm;                    // [HW][CS]
  float* gs = sm + HW * CS;          // [HW][CS]
  const int c0 = blockIdx.y * CS, cl = threadIdx.x % CS, pg = threadIdx.x / CS;
  float acc[K * K], ab = 0.f;
#pragma unroll
  for (int t = 0; t < K * K; ++t) acc[t] = 0.f;
  constexpr bool kPrefetch = (WT == 8);
  float2 pfx[NPF], pfg[NPF];
  if (kPrefetch && (int)blockIdx.x < B) {
    tile_prefetch(pfx, x, ldx, (long)blockIdx.x * HW, c0, C);
    tile_prefetch(pfg, dy, lddy, (long)blockIdx.x * HW, c0, C);
  }
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    if (kPrefetch) {
      tile_commit(xs, pfx);
      tile_commit(gs, pfg);
    } else {
      load_tile(xs, x, ldx, (long)b * HW, HW, c0, C);
      load_tile(gs, dy, lddy, (long)b * HW, HW, c0, C);
    }
    __syncthreads();
    if (kPrefetch && b + (int)gridDim.x < B) {
      tile_prefetch(pfx, x, ldx, (long)(b + gridDim.x) * HW, c0, C);
      tile_prefetch(pfg, dy, lddy, (long)(b + gridDim.x) * HW, c0, C);
    }
    if (WT > 0) {
      for (int py = pg; py < H; py += PG) {
        float g[WT > 0 ? WT : 1];
        const float* grow = gs + (py * WT) * CS + cl;
#pragma unroll
        for (int j = 0; j < WT; ++j) { g[j] = grow[j * CS]; ab += g[j]; }
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
          const int yy = py + ky - K / 2;
          if (yy < 0 || yy >= H) continue;
          const float* xrow = xs + (yy * WT) * CS + cl;
          float xr[WT > 0 ? WT : 1];
#pragma unroll
          for (int i = 0; i < WT; ++i) xr[i] = xrow[i * CS];
#pragma unroll
          for (int kx = 0; kx < K; ++kx) {
            float a = acc[ky * K + kx];
#pragma unroll
            for (int j = 0; j < WT; ++j)
              if (j + kx - K / 2 >= 0 && j + kx - K / 2 < WT) a = fmaf(g[j], xr[j + kx - K / 2], a);
            acc[ky * K + kx] = a;
          }
        }
      }
    } else {
      const int segs = (W + SEG - 1) / SEG;
      for (int task = pg; task < H * segs; task += PG) {
        const int py = task / segs, px0 = (task % segs) * SEG;
        float g[SEG];
#pragma unroll
        for (int j = 0; j < SEG; ++j) {
          g[j] = (px0 + j < W) ? gs[(py * W + px0 + j) * CS + cl] : 0.f;
          ab += g[j];
        }
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
          const int yy = py + ky - K / 2;
          if (yy < 0 || yy >= H) continue;
          float xr[SEG + K - 1];
#pragma unroll
          for (int i = 0; i < SEG + K - 1; ++i) {
            const int xx = px0 + i - K / 2;
            xr[i] = (xx >= 0 && xx < W) ? xs[(yy * W + xx) * CS + cl] : 0.f;
          }
#pragma unroll
          for (int kx = 0; kx < K; ++kx) {
            float a = acc[ky * K + kx];
#pragma unroll
            for (int j = 0; j < SEG; ++j) a = fmaf(g[j], xr[j + kx], a);
            acc[ky * K + kx] = a;
          }
        }
      }
    }
  }
  // reduce the PG thread groups through shared memory, then one atomic per (channel, tap) per CTA
  __syncthreads();
  float* red = sm;                   // [PG][K*K + 1][CS]
#pragma unroll
  for (int t = 0; t < K * K; ++t) red[(pg * (K * K + 1) + t) * CS + cl] = acc[t];
  red[(pg * (K * K + 1) + K * K) * CS + cl] = ab;
  __syncthreads();
  for (int idx = threadIdx.x; idx < (K * K + 1) * CS; idx += CS * PG) {
    const int t = idx / CS, cc = idx % CS;
    if (c0 + cc >= C) continue;
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < PG; ++q) s += red[(q * (K * K + 1) + t) * CS + cc];
    // y = scale * (conv(x) + bias):  dw = scale * (dy * x), dbias = scale * sum dy, dscale = sum dy * (conv(x) + bias)
    const float scl = scale ? scale[c0 + cc] : 1.f;
    if (t < K * K) {
      atomicAdd(dw + (c0 + cc) * K * K + t, s * scl);
      if (dscale) atomicAdd(dscale + c0 + cc, s * w[(c0 + cc) * K * K + t]);
    } else {
      if (dbias) atomicAdd(dbias + c0 + cc, s * scl);
      if (dscale && bias) atomicAdd(dscale + c0 + cc, s * bias[c0 + cc]);
    }
  }
}

// Banded flavour for maps whose x and dy tiles do not fit shared memory together (24 x 24 maps of the 96 x 96 STL-10 recipe): the
// map is walked in bands of BH output rows; x rows [y0 - K/2, y0 + BH + K/2) and dy rows [y0, y0 + BH) are staged per band.
template <typename T, int K>
__global__ void __launch_bounds__(CS * PG) dw_wgrad_band_kernel(const T* __restrict__ x, int ldx, const T* __restrict__ dy,
                                                                int lddy, int B, int H, int W, int C, float* __restrict__ dw,
                                                                float* __restrict__ dbias, const float* __restrict__ w,
                                                                const float* __restrict__ bias, const float* __restrict__ scale,
                                                                float* __restrict__ dscale, int BH) {
  QV_PDL_ENTRY();
  extern __shared__ float sm[];
  const int HW = H * W, XR = BH + K - 1;
  float* xs = sm;                    // [XR * W][CS]
  float* gs = sm + XR * W * CS;      // [BH * W][CS]
  const int c0 = blockIdx.y * CS, cl = threadIdx.x % CS, pg = threadIdx.x / CS;
  float acc[K * K], ab = 0.f;
#pragma unroll
  for (int t = 0; t < K * K; ++t) acc[t] = 0.f;
  const int segs = (W + SEG - 1) / SEG;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    for (int y0 = 0; y0 < H; y0 += BH) {
      const int y1 = min(H, y0 + BH), ylo = max(0, y0 - K / 2), yhi = min(H, y1 + K / 2);
      __syncthreads();
      load_tile(xs, x, ldx, (long)b * HW + (long)ylo * W, (yhi - ylo) * W, c0, C);
      load_tile(gs, dy, lddy, (long)b * HW + (long)y0 * W, (y1 - y0) * W, c0, C);
      __syncthreads();
      for (int task = pg; task < (y1 - y0) * segs; task += PG) {
        const int py = y0 + task / segs, px0 = (task % segs) * SEG;
        float g[SEG];
#pragma unroll
        for (int j = 0; j < SEG; ++j) {
          g[j] = (px0 + j < W) ? gs[((py - y0) * W + px0 + j) * CS + cl] : 0.f;
          ab += g[j];
        }
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
          const int yy = py + ky - K / 2;
          if (yy < 0 || yy >= H) continue;
          float xr[SEG + K - 1];
#pragma unroll
          for (int i = 0; i < SEG + K - 1; ++i) {
            const int xx = px0 + i - K / 2;
            xr[i] = (xx >= 0 && xx < W) ? xs[((yy - ylo) * W + xx) * CS + cl] : 0.f;
          }
#pragma unroll
          for (int kx = 0; kx < K; ++kx) {
            float a = acc[ky * K + kx];
#pragma unroll
            for (int j = 0; j < SEG; ++j) a = fmaf(g[j], xr[j + kx], a);
            acc[ky * K + kx] = a;
          }
        }
      }
    }
  }
  __syncthreads();
  float* red = sm;                   // [PG][K*K + 1][CS]
#pragma unroll
  for (int t = 0; t < K * K; ++t) red[(pg * (K * K + 1) + t) * CS + cl] = acc[t];
  red[(pg * (K * K + 1) + K * K) * CS + cl] = ab;
  __syncthreads();
  for (int idx = threadIdx.x; idx < (K * K + 1) * CS; idx += CS * PG) {
    const int t = idx / CS, cc = idx % CS;
    if (c0 + cc >= C) continue;
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < PG; ++q) s += red[(q * (K * K + 1) + t) * CS + cc];
    const float scl = scale ? scale[c0 + cc] : 1.f;
    if (t < K * K) {
      atomicAdd(dw + (c0 + cc) * K * K + t, s * scl);
      if (dscale) atomicAdd(dscale + c0 + cc, s * w[(c0 + cc) * K * K + t]);
    } else {
      if (dbias) atomicAdd(dbias + c0 + cc, s * scl);
      if (dscale && bias) atomicAdd(dscale + c0 + cc, s * bias[c0 + cc]);
    }
  }
}

template <typename T, int K, bool F, int WT>
int launch_fwd(cudaStream_t s, const DwP& p, dim3 grid, size_t smem) {
  if (smem > 48 * 1024) QV_CUDA(cudaFuncSetAttribute(dw_fwd_kernel<T, K, F, WT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  qv_launch(dw_fwd_kernel<T, K, F, WT>, grid, CS * PG, smem, s, (const T*)p.x, p.ldx, p.B, p.H, p.W, p.C, p.w, p.bias, p.scale, (T*)p.y, p.ldy,
                                                         (const T*)p.resid, p.ldr, (const T*)p.resid2, p.ldr2, (T*)p.copy, p.ldcp);
  QV_LAUNCH_CHECK();
  return 0;
}
template <typename T, int K, bool F>
int launch_fwd_w(cudaStream_t s, const DwP& p, dim3 grid, size_t smem) {
  switch (p.H == p.W ? p.W : 0) {
    case 4: return launch_fwd<T, K, F, 4>(s, p, grid, smem);
    case 8: return launch_fwd<T, K, F, 8>(s, p, grid, smem);
    case 16: return launch_fwd<T, K, F, 16>(s, p, grid, smem);
    case 24: return launch_fwd<T, K, F, 24>(s, p, grid, smem);
  }
  return launch_fwd<T, K, F, 0>(s, p, grid, smem);
}

template <typename T, int K>
int run_fwd(cudaStream_t s, const DwP& p, bool flip) {
  const size_t smem = (size_t)p.H * p.W * CS * sizeof(float);
  QV_CHECK(smem <= 200 * 1024, "dwconv: %dx%d feature map too large for one shared-memory tile", p.H, p.W);
  const int cch = cdiv(p.C, CS);
  const int occ = max(1, min(4, (int)(200 * 1024 / (smem + 1024))));
  dim3 grid(max(1, min(p.B, qv_num_sms() * occ * 2 / cch)), cch);
  return flip ? launch_fwd_w<T, K, true>(s, p, grid, smem) : launch_fwd_w<T, K, false>(s, p, grid, smem);
}

template <typename T, int K, int WT>
int launch_wgrad(cudaStream_t s, const T* x, int ldx, const T* dy, int lddy, int B, int H, int W, int C, float* dw, float* dbias,
                 dim3 grid, size_t smem, const DwScale& sc) {
  if (smem > 48 * 1024) QV_CUDA(cudaFuncSetAttribute(dw_wgrad_kernel<T, K, WT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  qv_launch(dw_wgrad_kernel<T, K, WT>, grid, CS * PG, smem, s, x, ldx, dy, lddy, B, H, W, C, dw, dbias, sc.w, sc.bias, sc.scale, sc.dscale);
  QV_LAUNCH_CHECK();
  return 0;
}

template <typename T, int K>
int run_wgrad(cudaStream_t s, const T* x, int ldx, const T* dy, int lddy, int B, int H, int W, int C, float* dw, float* dbias,
              const DwScale& sc) {
  const size_t tile = (size_t)2 * H * W * CS * sizeof(float), red = (size_t)PG * (K * K + 1) * CS * sizeof(float);
  if (tile > 200 * 1024) {   // x and dy tiles of the whole map do not fit: banded kernel
    int BH = H;
    auto band_bytes = [&](int bh) { return (size_t)(2 * bh + K - 1) * W * CS * sizeof(float); };
    while (BH > 1 && band_bytes(BH) > 200 * 1024) --BH;
    const size_t bsm = band_bytes(BH) > red ? band_bytes(BH) : red;
    QV_CHECK(bsm <= 200 * 1024, "dwconv wgrad: %dx%d feature map too large", H, W);
    const int cchb = cdiv(C, CS);
    dim3 gridb(max(1, min(B, qv_num_sms() / cchb)), cchb);
    QV_CUDA(cudaFuncSetAttribute(dw_wgrad_band_kernel<T, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm));
    qv_launch(dw_wgrad_band_kernel<T, K>, gridb, CS * PG, bsm, s, x, ldx, dy, lddy, B, H, W, C, dw, dbias, sc.w, sc.bias, sc.scale, sc.dscale, BH);
    QV_LAUNCH_CHECK();
    return 0;
  }
  const size_t smem = tile > red ? tile : red;
  const int cch = cdiv(C, CS);
  const int occ = max(1, min(6, (int)(200 * 1024 / (smem + 1024))));
  dim3 grid(max(1, min(B, qv_num_sms() * occ / cch)), cch);
  switch (H == W ? W : 0) {
    case 4: return launch_wgrad<T, K, 4>(s, x, ldx, dy, lddy, B, H, W, C, dw, dbias, grid, smem, sc);
    case 8: return launch_wgrad<T, K, 8>(s, x, ldx, dy, lddy, B, H, W, C, dw, dbias, grid, smem, sc);
    case 16: return launch_wgrad<T, K, 16>(s, x, ldx, dy, lddy, B, H, W, C, dw, dbias, grid, smem, sc);
    case 24: return launch_wgrad<T, K, 24>(s, x, ldx, dy, lddy, B, H, W, C, dw, dbias, grid, smem, sc);
  }
  return launch_wgrad<T, K, 0>(s, x, ldx, dy, lddy, B, H, W, C, dw, dbias, grid, smem, sc);
}

template <typename T>
int fwd_t(cudaStream_t s, const DwP& p, bool flip) {
  switch (p.K) {
    case 3: return run_fwd<T, 3>(s, p, flip);
    case 5: return run_fwd<T, 5>(s, p, flip);
    case 7: return run_fwd<T, 7>(s, p, flip);
  }
  qv_set_error("dwconv: kernel size %d not supported (3, 5, 7)", p.K);
  return 1;
}
template <typename T>
int wgrad_t(cudaStream_t s, int K, const T* x, int ldx, const T* dy, int lddy, int B, int H, int W, int C, float* dw, float* dbias,
            const DwScale& sc) {
  switch (K) {
    case 3: return run_wgrad<T, 3>(s, x, ldx, dy, lddy, B, H, W, C, dw, dbias, sc);
    case 5: return run_wgrad<T, 5>(s, x, ldx, dy, lddy, B, H, W, C, dw, dbias, sc);
    case 7: return run_wgrad<T, 7>(s, x, ldx, dy, lddy, B, H, W, C, dw, dbias, sc);
  }
  qv_set_error("dwconv: kernel size %d not supported (3, 5, 7)", K);
  return 1;
}

}  // namespace

int dw2d_fwd(cudaStream_t s, int dt, const DwP& p, bool flip) {
  if (p.B <= 0) return 0;
  QV_CHECK(p.C % 2 == 0 && p.ldx % 2 == 0, "dwconv: channel count / row pitch must be even");
  if (dt == QV_BF16 && dwt_ok(p)) return dwt_fwd(s, p, flip);      // 8 x 8 maps: the stencil as a tensor-core product (dwconv_mma.cu)
  return dt == QV_BF16 ? fwd_t<bf16>(s, p, flip) : fwd_t<float>(s, p, flip);
}
int dw2d_wgrad(cudaStream_t s, int dt, int K, const void* x, int ldx, const void* dy, int lddy, int B, int H, int W, int C,
               float* dw, float* dbias, const DwScale& sc) {
  if (B <= 0) return 0;
  QV_CHECK(C % 2 == 0 && ldx % 2 == 0 && lddy % 2 == 0, "dwconv wgrad: channel count / row pitch must be even");
  if (dt == QV_BF16) return wgrad_t<bf16>(s, K, (const bf16*)x, ldx, (const bf16*)dy, lddy, B, H, W, C, dw, dbias, sc);
  return wgrad_t<float>(s, K, (const float*)x, ldx, (const float*)dy, lddy, B, H, W, C, dw, dbias, sc);
}

extern "C" int qavit_dwconv_forward(const void* x, int is_bf16, int B, int H, int W, int C, int K, const float* w,
                                    const float* bias, void* y, void* stream) {
  DwP p{};
  p.x = x; p.ldx = C; p.B = B; p.H = H; p.W = W; p.C = C; p.K = K; p.w = w; p.bias = bias; p.y = y; p.ldy = C;
  return dw2d_fwd((cudaStream_t)stream, is_bf16 ? QV_BF16 : QV_F32, p, false);
}

extern "C" int qavit_dwconv_backward(const void* x, const void* dy, int is_bf16, int B, int H, int W, int C, int K,
                                     const float* w, void* dx, float* dw, float* dbias, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const int dt = is_bf16 ? QV_BF16 : QV_F32;
  if (dx) {
    DwP p{};
    p.x = dy; p.ldx = C; p.B = B; p.H = H; p.W = W; p.C = C; p.K = K; p.w = w; p.y = dx; p.ldy = C;
    QV_TRY(dw2d_fwd(s, dt, p, true));
  }
  return dw2d_wgrad(s, dt, K, x, C, dy, C, B, H, W, C, dw, dbias);
}
