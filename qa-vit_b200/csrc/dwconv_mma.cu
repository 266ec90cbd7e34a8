// Depthwise k x k convolutions on 8 x 8 channels-last maps as tensor-core products (bf16 runs; the ConvNeXt / LMFAdapter stencils
// of HQAViT's lateral path at CIFAR resolution, H:722, 811-812).
//
// On an 8 x 8 map a depthwise 'same' convolution of one channel is a dense 64 x 64 operator T_c (Toeplitz-block matrix of the k x k
// taps, zero where a tap falls outside the map): y_c[p_out, image] = sum_{p_in} T_c[p_out, p_in] x_c[p_in, image].  With the images
// as the N dimension that is an m64 n(images) k64 product per channel on mma.sync.m16n8k16 -- 1.3 x the FLOPs of the stencil (the
// zeros of T_c) at ~30 x the FMA rate, and the CUDA-core kernel (dwconv_nhwc.cu: 33 % of the FP32 peak, MIO- / LG-throttled) was
// the largest non-GEMM family of the step.  A CTA owns 16 channels (one 32 B sector per pixel of the channels-last rows), builds their
// T_c once in shared memory and walks 16-image tiles: pixel rows are transposed into [channel][image][pixel] on the way in, and the
// accumulators are transposed back through shared memory into [image][pixel][channel] rows on the way out (bias, residuals, copy
// applied in that row-wise pass).  Taps are rounded to bf16 like the reference's autocast convolution rounds its weights.
// FLIP (the input gradient) only changes how T_c is filled.  fp32 runs and other map sizes keep the kernels of dwconv_nhwc.cu.
#include "kernels.h"

namespace {

constexpr int HW = 64, CH = 16, NB = 16, TP = 72, NT = 256;
constexpr int XC = NB * TP;                      // bf16 per channel plane of Xt
__device__ __forceinline__ int xt_base(int ch) { return ch * XC + (ch >> 2) * 16; }   // 32 B skew per 4 channels: the four channel
                                                                                     // quads a store instruction touches sit in different banks
constexpr int YW = 9, YI = HW * YW + 1;          // staging pitches in 32-bit words: 9 per (image, pixel) row of 8 words, 577 per image

struct DwtP {
  const bf16* x; int ldx;
  int B, C, K, flip;
  const float* w; const float* bias;
  bf16* y; int ldy;
  const bf16* resid; int ldr;
  const bf16* resid2; int ldr2;
  bf16* copy; int ldcp;
};

__device__ __forceinline__ uint32_t sa(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm4(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
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

__device__ __forceinline__ void tile_prefetch(uint4* pf, const DwtP& p, int b0, int c0) {
  // NB images x 64 pixels x 2 halves (8 channels = 16 B each): 8 vectors per thread
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = threadIdx.x + i * NT, bp = idx >> 1, half = idx & 1;
    pf[i] = make_uint4(0u, 0u, 0u, 0u);
    if (b0 + (bp >> 6) < p.B) pf[i] = *reinterpret_cast<const uint4*>(p.x + ((long)b0 * HW + bp) * p.ldx + c0 + half * 8);
  }
}

// One CTA = 16 channels (a 32 B sector per pixel row), one CTA per SM.  (An 8-channel / two-CTAs-per-SM variant measured 30 % slower:
// 16 B per pixel row halves the sector efficiency of every global access.)
__global__ void __launch_bounds__(NT, 1) dwt_fwd_kernel(DwtP p) {
  QV_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t smraw[];
  bf16* Tm = reinterpret_cast<bf16*>(smraw);               // [CH][64][TP]   T_c[p_out][p_in]
  bf16* Xt = Tm + CH * HW * TP;                             // [CH][NB][TP]   x_c[image][p_in]  (+ skew)
  uint32_t* Ys = reinterpret_cast<uint32_t*>(Xt + CH * XC + 64);   // [NB][64] rows of 8 words (16 channels), pitches YI / YW
  float* wp = reinterpret_cast<float*>(Ys);                 // [CH][15][16] taps centred at (7, 7), zero outside the stencil; only alive
                                                            // while T_c is filled (the staging rows are first written two barriers later)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int c0 = blockIdx.y * CH;
  const int ntiles = (p.B + NB - 1) / NB;

  uint4 pf[8];
  if ((int)blockIdx.x < ntiles) tile_prefetch(pf, p, blockIdx.x * NB, c0);

  // ---- T_c of this CTA's channels: padded tap table first (no branches / global loads in the 32 K-entry fill)
  for (int i = tid; i < CH * 15 * 16; i += NT) {
    const int c = i / 240, dy = (i % 240) / 16 - 7, dx = i % 16 - 7, r = p.K / 2;
    float v = 0.f;
    if (dy >= -r && dy <= r && dx >= -r && dx <= r) {
      int tp = (dy + r) * p.K + dx + r;
      if (p.flip) tp = p.K * p.K - 1 - tp;
      v = p.w[(long)(c0 + c) * p.K * p.K + tp];
    }
    wp[i] = v;
  }
  __syncthreads();
  for (int i = tid; i < CH * HW * (HW / 2); i += NT) {
    const int c = i / (HW * (HW / 2)), po = (i / (HW / 2)) % HW, pi = (i % (HW / 2)) * 2;
    const float* q = wp + c * 240 + ((pi >> 3) - (po >> 3) + 7) * 16 + (pi & 7) - (po & 7) + 7;
    *reinterpret_cast<uint32_t*>(Tm + (c * HW + po) * TP + pi) = pack2(q[0], q[1]);
  }
  const int ch0 = warp * 2;                                 // this warp's channel pair
  const int zband = p.K == 7 ? 2 : 1;                       // blocks = pairs of map rows: |row distance| <= K / 2 reaches 1 (K <= 5) or 2 blocks
  const float bs0 = p.bias ? p.bias[c0 + ch0] : 0.f, bs1 = p.bias ? p.bias[c0 + ch0 + 1] : 0.f;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b0 = tile * NB, nimg = min(NB, p.B - b0);
    __syncthreads();                                        // T_c complete (first pass); the previous tile's Xt / Ys are consumed
    // ---- transpose the tile into [channel][image][pixel]; LMFAdapter's identity branch copies the rows on the way.
    // A thread holds 8 channels of one pixel; lane ^ 2 holds the same channels of the neighbouring pixel.  The pair swaps halves so
    // that each lane ends up with 4 channels x 2 pixels = four 32-bit stores (pixel pair of one channel) instead of eight 16-bit ones.
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = tid + i * NT, bp = idx >> 1, half = idx & 1;
      const int img = bp >> 6, pix = bp & 63, odd = pix & 1;
      if (p.copy && b0 + img < p.B) *reinterpret_cast<uint4*>(p.copy + ((long)b0 * HW + bp) * p.ldcp + c0 + half * 8) = pf[i];
      const uint32_t keep0 = odd ? pf[i].z : pf[i].x, keep1 = odd ? pf[i].w : pf[i].y;
      const uint32_t send0 = odd ? pf[i].x : pf[i].z, send1 = odd ? pf[i].y : pf[i].w;
      const uint32_t got0 = __shfl_xor_sync(0xffffffffu, send0, 2), got1 = __shfl_xor_sync(0xffffffffu, send1, 2);
      const uint32_t ev0 = odd ? got0 : keep0, od0 = odd ? keep0 : got0, ev1 = odd ? got1 : keep1, od1 = odd ? keep1 : got1;
      const int chb = half * 8 + odd * 4;                  // first of this lane's 4 channels
      uint32_t* d = reinterpret_cast<uint32_t*>(Xt + xt_base(chb) + img * TP + (pix & ~1));
      d[0] = __byte_perm(ev0, od0, 0x5410);                // channel chb     : pixels (even, odd)
      d[XC / 2] = __byte_perm(ev0, od0, 0x7632);           // channel chb + 1
      d[XC] = __byte_perm(ev1, od1, 0x5410);               // channel chb + 2
      d[XC + XC / 2] = __byte_perm(ev1, od1, 0x7632);      // channel chb + 3
    }
    __syncthreads();
    if (tile + (int)gridDim.x < ntiles) tile_prefetch(pf, p, (tile + gridDim.x) * NB, c0);

    // ---- y_c[64 x NB] = T_c[64 x 64] x_c[64 x NB] for the warp's two channels
    float acc[2][4][2][4];
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const bf16* Tc = Tm + (ch0 + cc) * HW * TP;
      const bf16* Xc = Xt + xt_base(ch0 + cc);
      uint32_t bfr[2][2][4];                                // [n-tile][k half][b0 b1 of two k-steps]
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int kh = 0; kh < 2; ++kh)
          ldsm4(bfr[nt][kh], sa(Xc + (nt * 8 + (lane & 7)) * TP + kh * 32 + (lane >> 3) * 8));
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) acc[cc][mt][nt][0] = acc[cc][mt][nt][1] = acc[cc][mt][nt][2] = acc[cc][mt][nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if ((ks > mt ? ks - mt : mt - ks) > zband) continue;   // a 16 x 16 block of T_c beyond the stencil's reach is all zeros
          uint32_t a[4];
          const int mat = lane >> 3, r = lane & 7;
          ldsm4(a, sa(Tc + (mt * 16 + r + (mat & 1) * 8) * TP + ks * 16 + (mat >> 1) * 8));
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) mma16816(acc[cc][mt][nt], a, bfr[nt][ks >> 1][(ks & 1) * 2], bfr[nt][ks >> 1][(ks & 1) * 2 + 1]);
        }
      }
    }
    // ---- accumulators -> staging rows [image][pixel][16 channels]: the channel pair of a warp is one 32-bit word
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int pix = mt * 16 + g + (e >> 1) * 8, img = nt * 8 + 2 * t + (e & 1);
          Ys[img * YI + pix * YW + warp] = pack2(acc[0][mt][nt][e] + bs0, acc[1][mt][nt][e] + bs1);
        }
    __syncthreads();
    // ---- row-wise pass: residuals, 16 B stores
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = tid + i * NT, bp = idx >> 1, half = idx & 1, img = bp >> 6, pix = bp & 63;
      if (img >= nimg) break;
      const uint32_t* src = Ys + img * YI + pix * YW + half * 4;
      uint32_t v[4] = {src[0], src[1], src[2], src[3]};
      const long row = (long)b0 * HW + bp;
      if (p.resid || p.resid2) {
        float f[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 q = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v[k]));
          f[2 * k] = q.x; f[2 * k + 1] = q.y;
        }
        if (p.resid) {
          float r8[8];
          load_vec<8>(p.resid + row * p.ldr + c0 + half * 8, r8);
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] += r8[k];
        }
        if (p.resid2) {
          float r8[8];
          load_vec<8>(p.resid2 + row * p.ldr2 + c0 + half * 8, r8);
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] += r8[k];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = pack2(f[2 * k], f[2 * k + 1]);
      }
      *reinterpret_cast<uint4*>(p.y + row * p.ldy + c0 + half * 8) = make_uint4(v[0], v[1], v[2], v[3]);
    }
  }
}

}  // namespace

bool dwt_ok(const DwP& p) {
  static const bool off = getenv("QV_NO_DWT") != nullptr;
  return !off && p.H == 8 && p.W == 8 && p.C % 16 == 0 && p.K == 7 && p.scale == nullptr && p.ldx % 8 == 0 &&
         p.ldy % 8 == 0 && (!p.resid || p.ldr % 8 == 0) && (!p.resid2 || p.ldr2 % 8 == 0) && (!p.copy || p.ldcp % 8 == 0) &&
         ((uintptr_t)p.x & 15) == 0 && ((uintptr_t)p.y & 15) == 0 && (!p.resid || ((uintptr_t)p.resid & 15) == 0) &&
         (!p.resid2 || ((uintptr_t)p.resid2 & 15) == 0) && (!p.copy || ((uintptr_t)p.copy & 15) == 0);
}

int dwt_fwd(cudaStream_t s, const DwP& q, bool flip) {
  DwtP p{};
  p.x = static_cast<const bf16*>(q.x); p.ldx = q.ldx; p.B = q.B; p.C = q.C; p.K = q.K; p.flip = flip ? 1 : 0;
  p.w = q.w; p.bias = q.bias;
  p.y = static_cast<bf16*>(q.y); p.ldy = q.ldy;
  p.resid = static_cast<const bf16*>(q.resid); p.ldr = q.ldr;
  p.resid2 = static_cast<const bf16*>(q.resid2); p.ldr2 = q.ldr2;
  p.copy = static_cast<bf16*>(q.copy); p.ldcp = q.ldcp;
  const size_t smem = ((size_t)CH * HW * TP + (size_t)CH * XC + 64) * 2 + (size_t)NB * YI * 4 + 16;
  QV_CUDA(cudaFuncSetAttribute(dwt_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int groups = q.C / CH, ntiles = cdiv(q.B, NB);
  const int gx = max(1, min(ntiles, qv_num_sms() / groups));
  qv_launch(dwt_fwd_kernel, dim3(gx, groups), NT, smem, s, p);
  QV_LAUNCH_CHECK();
  return 0;
}
