// nn.Dropout / DropPath sites of the quad block that are not inside an attention kernel (H:465, 529, 592, 625 branch
// output dropout; H:654-656 BottleneckMLP; H:710 CCFFFN; H:1082-1083 DropPath) as one pass over the activation, the mask
// regenerated from the Philox snapshot: backward applies the same function to the gradient.  bf16 runs apply these sites
// inside their producers (tcgen05 GEMM epilogue, ln_bwd, gamma_bwd) with the SAME element ids; this pass serves the fp32
// run (SIMT GEMMs), shapes the tcgen05 kernel does not take, pos_drop (qavit_dropout_*), and the per-image DropPath
// scale vector.
#include "kernels.h"

namespace {

template <typename T>
__global__ void __launch_bounds__(256) drop_rows_kernel(T* __restrict__ x, int ldx, long rows, int C, DropP d,
                                                        const float* __restrict__ rowscale, int rows_per_img,
                                                        const float* __restrict__ resid, int ldr, const float* __restrict__ scale,
                                                        float* __restrict__ out, int ldo) {
  QV_PDL_ENTRY();
  const int cpr = C / 8;
  const long total = rows * cpr;
  const bool masked = d.p > 0.f;
  DropState st{};
  if (masked) st = drop_state(d);
  const float sc = scale ? *scale : 1.f;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long row = i / cpr;
    const int c0 = (int)(i % cpr) * 8;
    float v[8], k[8];
    load_vec<8>(x + row * ldx + c0, v);
    const float rs = rowscale ? rowscale[row / rows_per_img] : 1.f;
    if (masked) {
      drop_keep8(st, (unsigned long long)i, k);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= k[j] * rs;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= rs;
    }
    store_vec<8>(x + row * ldx + c0, v);
    if (out) {
      float r[8];
      load_vec<8>(resid + row * ldr + c0, r);
      if (sizeof(T) == 2) {   // the product the next op sees is the stored (rounded) value
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = fmaf(sc, v[j], r[j]);
      store_vec<8>(out + row * ldo + c0, r);
    }
  }
}

__global__ void droppath_scales_kernel(DropP d, int B, float* rs1, float* rs2) {
  QV_PDL_ENTRY();
  const DropState st = drop_state(d);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * B; i += gridDim.x * blockDim.x)
    (i < B ? rs1 : rs2 - B)[i] = d.p > 0.f ? drop_keep1(st, (unsigned long long)i) : 1.f;
}

}  // namespace

int drop_rows(cudaStream_t s, int dt, void* x, int ldx, long rows, int C, const DropP& d, const float* rowscale, int rows_per_img,
              const float* resid, int ldr, const float* scale, float* out, int ldo) {
  if (rows <= 0) return 0;
  QV_CHECK(C % 8 == 0 && ldx % 8 == 0, "drop_rows: C = %d / ld = %d must be multiples of 8", C, ldx);
  QV_CHECK(d.p == 0.f || d.rng, "drop_rows: missing rng snapshot");
  QV_CHECK(!out || (resid && ldr % 4 == 0 && ldo % 4 == 0), "drop_rows: residual form needs resid / aligned pitches");
  const long total = rows * (C / 8);
  const int grid = (int)((total + 255) / 256 < (long)qv_num_sms() * 8 ? (total + 255) / 256 : (long)qv_num_sms() * 8);
  if (dt == QV_F32)
    qv_launch(drop_rows_kernel<float>, grid, 256, 0, s, (float*)x, ldx, rows, C, d, rowscale, rows_per_img, resid, ldr, scale, out, ldo);
  else
    qv_launch(drop_rows_kernel<bf16>, grid, 256, 0, s, (bf16*)x, ldx, rows, C, d, rowscale, rows_per_img, resid, ldr, scale, out, ldo);
  QV_LAUNCH_CHECK();
  return 0;
}

int droppath_scales(cudaStream_t s, const DropP& d, int B, float* rs1, float* rs2) {
  QV_CHECK(d.p == 0.f || d.rng, "droppath_scales: missing rng snapshot");
  qv_launch(droppath_scales_kernel, cdiv(2 * B, 256), 256, 0, s, d, B, rs1, rs2);
  QV_LAUNCH_CHECK();
  return 0;
}
