// tcgen05 / TMEM / TMA GEMMs for sm_100a (bf16 operands, fp32 accumulation in tensor memory).
//
//   tc_gemm_nt : C[M, N] = A[M, K] W[N, K]^T (+ epilogue).  Persistent, warp-specialised: warp 0 = TMA producer,
//                warp 1 = MMA issuer (+ TMEM alloc), warps 2..5 = epilogue (tcgen05.ld -> registers -> global).
//                128 x BLOCK_N output tiles, K streamed in 64-wide (128 B, SWIZZLE_128B) k-blocks through a 4-stage
//                mbarrier ring; two TMEM accumulator buffers so the epilogue of tile i overlaps the MMAs of tile i+1.
//                Used for every forward projection and (with the W^T copy) every dX GEMM.
//   tc_gemm_tn : dW[N, K] += scale * dY[M, N]^T X[M, K].  Both operands are MN-major for the MMA (the contraction
//                runs over rows), loaded by TMA as [64 rows x 64 cols] SWIZZLE_128B boxes and consumed through
//                MN-major UMMA descriptors -- no transposed copies of activations are ever made.  The row range is
//                split over CTAs; each CTA accumulates its slice in TMEM and reduces into dW with fp32 atomics.
//
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix / instruction descriptor" tables.
#include <cuda.h>

#include <stdlib.h>
#include "kernels.h"

namespace {

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (an error the tests report), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// smem (128 B-swizzled box) -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, issued by one thread for the CTA
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// arrives on the mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// The wait names the destination registers as in/out operands so no use of them can be scheduled above it.
__device__ __forceinline__ void tmem_ld_wait16(float* v) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]),
                 "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])
               :
               : "memory");
}

// ---- descriptors
// shared-memory matrix descriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout [61,64)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
// instruction descriptor, kind::f16: D=f32 [4,6)=1 | A=bf16 [7,10)=1 | B=bf16 [10,13)=1 | a_major [15] | b_major [16]
//                                   | N>>3 [17,23) | M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int BM = 128;       // output rows per tile (UMMA M)
constexpr int BK = 64;        // k-block: 64 bf16 = 128 B = one SWIZZLE_128B span
constexpr int STAGES = 4;
constexpr int EPI_WARPS = 12;                  // 3 per TMEM lane quarter: one warp per SMSP cannot hide its own latencies
constexpr int NTHREADS = 64 + EPI_WARPS * 32;   // NT kernel: TMA warp + MMA warp + epilogue warps
constexpr int TN_THREADS = 192;                 // TN kernel: TMA warp + MMA warp + 4 epilogue warps

struct NtArgs {
  int M, N, K;     // C[M, N], contraction K
  int BN;          // columns per tile (multiple of 16, <= 256)
  int n_slices;    // ceil(N / BN)
  int m_tiles;
  uint32_t tmem_cols;
  int stages;      // depth of the TMA ring (4; 3 for the widest tiles so staging still fits; up to 8 A-only stages with w_res)
  int w_res;       // this CTA's [BN x K] weight slice stays resident in shared memory for all its m-tiles (loaded once): the
                   // ring then carries A only.  Re-loading W per tile was 60 % of the L2 -> SM bytes of the block projections
  int tma_store;   // wide epilogue: 64-column groups leave through a TMA store (map_c valid)
  GemmEpi e;
};

// ---- epilogue staging.  tcgen05.ld hands every thread one accumulator ROW; writing rows straight to global memory
// makes each warp store touch 32 different lines with 16 B each (measured: 4x the needed L2 write sectors, 9 % tensor
// pipe).  Instead each epilogue warp stages 32 rows x 32 columns in shared memory and flushes them with row-contiguous,
// sector-complete 16 B accesses; residual inputs come in the same way and the bias sits in shared memory.
constexpr int EPI_GROUP = 32;                    // columns staged per pass
constexpr int EPI_ROWB = 144;                    // staging row pitch in bytes (32 fp32 + 16 pad: conflict-free row writes)
constexpr int EPI_WARP_BYTES = 32 * EPI_ROWB;   // one staging buffer per warp, reused for residual-in, C and C2

__device__ __forceinline__ void stage_store16(uint8_t* rowp, int c, const float* val, bool f32) {
  if (f32) {
    float4* d = reinterpret_cast<float4*>(rowp + c * 64);
#pragma unroll
    for (int j = 0; j < 4; ++j) d[j] = make_float4(val[4 * j], val[4 * j + 1], val[4 * j + 2], val[4 * j + 3]);
  } else {
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      __nv_bfloat162 h = __floats2bfloat162_rn(val[2 * j], val[2 * j + 1]);
      pk[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    uint4* d = reinterpret_cast<uint4*>(rowp + c * 32);
    d[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    d[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
}

// bf16 flavour with a compile-time number of 16 B chunks per row (8 / 4 / 2 for 64 / 32 / 16 staged columns): fully
// unrolled, one pointer bump per pass (the generic loop above spent more issue slots on index math than on the copies)
template <int CPR>
__device__ __forceinline__ void stage_flush_bf16(const uint8_t* st, void* C, long ldc, long row0, int M, int col0, int lane) {
  constexpr int RPP = 32 / CPR;                                // rows covered per pass
  const int r0 = lane / CPR, ch = lane % CPR;
  const int nvalid = (int)min(32L, (long)M - row0);
  uint8_t* gp = static_cast<uint8_t*>(C) + ((row0 + r0) * ldc + col0) * 2 + ch * 16;
  const uint8_t* sp = st + r0 * EPI_ROWB + ch * 16;
  const long gstep = (long)RPP * ldc * 2;
#pragma unroll
  for (int k = 0; k < CPR; ++k) {
    if (r0 + k * RPP < nvalid) *reinterpret_cast<uint4*>(gp) = *reinterpret_cast<const uint4*>(sp);
    gp += gstep;
    sp += RPP * EPI_ROWB;
  }
}

// any element size: CPR = 16 B chunks per staged row (gw * esz / 16), optional fp32 accumulate
template <int CPR, bool ACCUM>
__device__ __forceinline__ void stage_flush_u(const uint8_t* st, void* C, long pitch_bytes, long row0, int M, long col_byte0, int lane) {
  constexpr int RPP = 32 / CPR;
  const int r0 = lane / CPR, ch = lane % CPR;
  const int nvalid = (int)min(32L, (long)M - row0);
  uint8_t* gp = static_cast<uint8_t*>(C) + (row0 + r0) * pitch_bytes + col_byte0 + ch * 16;
  const uint8_t* sp = st + r0 * EPI_ROWB + ch * 16;
#pragma unroll
  for (int k = 0; k < CPR; ++k) {
    if (r0 + k * RPP < nvalid) {
      const uint4 v = *reinterpret_cast<const uint4*>(sp);
      if (ACCUM) {
        float4 o = *reinterpret_cast<float4*>(gp);
        o.x += __uint_as_float(v.x); o.y += __uint_as_float(v.y); o.z += __uint_as_float(v.z); o.w += __uint_as_float(v.w);
        *reinterpret_cast<float4*>(gp) = o;
      } else {
        *reinterpret_cast<uint4*>(gp) = v;
      }
    }
    gp += RPP * pitch_bytes;
    sp += RPP * EPI_ROWB;
  }
}
__device__ __forceinline__ void stage_flush_any(const uint8_t* st, void* C, long ldc, bool f32, bool accum, long row0, int M, int col0,
                                                int gw, int lane) {
  const int esz = f32 ? 4 : 2, cpr = gw * esz / 16;
  const long pitch = ldc * esz, cb = (long)col0 * esz;
  if (accum) {
    if (cpr == 8) stage_flush_u<8, true>(st, C, pitch, row0, M, cb, lane);
    else stage_flush_u<4, true>(st, C, pitch, row0, M, cb, lane);
  } else if (cpr == 8) stage_flush_u<8, false>(st, C, pitch, row0, M, cb, lane);
  else if (cpr == 4) stage_flush_u<4, false>(st, C, pitch, row0, M, cb, lane);
  else stage_flush_u<2, false>(st, C, pitch, row0, M, cb, lane);
}

// fp32 global [32 rows x gw cols] -> staging (row-contiguous reads)
__device__ __forceinline__ void stage_fill_f32(uint8_t* st, const float* R, long ldr, long row0, int M, int col0, int gw, int lane) {
  const int cpr = gw / 4;                                      // 8 or 4
  const int sh = (cpr == 8) ? 3 : 2;
  for (int k = 0; k < cpr; ++k) {
    const int idx = k * 32 + lane, r = idx >> sh, ch = idx & (cpr - 1);
    const long row = row0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < M) v = *reinterpret_cast<const float4*>(R + row * ldr + col0 + ch * 4);
    *reinterpret_cast<float4*>(st + r * EPI_ROWB + ch * 16) = v;
  }
}

// bf16 global [32 rows x gw cols] -> staging (row-contiguous 16 B reads)
__device__ __forceinline__ void stage_fill_bf16(uint8_t* st, const bf16* R, long ldr, long row0, int M, int col0, int gw, int lane) {
  const int cpr = gw / 8;                                      // 4 or 2
  const int sh = (cpr == 4) ? 2 : 1;
  for (int k = 0; k < cpr; ++k) {
    const int idx = k * 32 + lane, r = idx >> sh, ch = idx & (cpr - 1);
    const long row = row0 + r;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (row < M) v = *reinterpret_cast<const uint4*>(R + row * ldr + col0 + ch * 8);
    *reinterpret_cast<uint4*>(st + r * EPI_ROWB + ch * 16) = v;
  }
}
// 16 staged values of this thread's row, chunk c
__device__ __forceinline__ void stage_read16(const uint8_t* my_row, int c, bool f32, float* r) {
  if (f32) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 q = *reinterpret_cast<const float4*>(my_row + (c * 16 + j) * 4);
      r[j] = q.x; r[j + 1] = q.y; r[j + 2] = q.z; r[j + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint4 q = *reinterpret_cast<const uint4*>(my_row + c * 32 + h * 16);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
        r[h * 8 + 2 * i] = f.x; r[h * 8 + 2 * i + 1] = f.y;
      }
    }
  }
}

// Diagnostic build only (make EXTRA=-DQV_GEMM_TRACE): per-phase clock64 / globaltimer stamps of the first and the last CTA of
// an NT launch, read back with qavit_test_gemm_trace_read (tools/gemm_trace.py).
#ifdef QV_GEMM_TRACE
__device__ long long g_trace[128];
__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TR(i) do { if (blockIdx.y == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) { \
  const int b_ = blockIdx.x ? 32 : 0; g_trace[b_ + (i)] = clock64(); g_trace[64 + b_ + (i)] = gtimer(); } } while (0)
#else
#define TR(i) do {} while (0)
#endif

// WIDE: the epilogue flavour is a compile-time choice (two instantiations) so that neither carries the other's registers
template <bool WIDE>
__global__ void __launch_bounds__(NTHREADS, 1)
tc_gemm_nt_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                  const __grid_constant__ CUtensorMap map_c, NtArgs p) {
  pdl_trigger();   // the wait comes after the set-up that does not touch the producer's memory (barriers, descriptors, TMEM)
  if (threadIdx.x == 0) TR(0);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  const int STG = p.stages;
  const int kblocks = (p.K + BK - 1) / BK;
  const int a_bytes = BM * BK * 2, w_bytes = p.BN * BK * 2;
  const int stage_bytes = p.w_res ? a_bytes : a_bytes + w_bytes;
  uint8_t* w_res_base = smem;                                      // [kblocks][BN x 128 B] swizzled boxes (w_res only)
  uint8_t* ring = smem + (p.w_res ? kblocks * w_bytes : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + STG * stage_bytes);
  uint64_t *full = bars, *empty = bars + STG, *tfull = bars + 2 * STG, *tempty = bars + 2 * STG + 2, *wfull = bars + 2 * STG + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STG + 5);
  uint8_t* epi_stage = ring + STG * stage_bytes + 1024;   // 1024-aligned: the TMA-store staging of the wide epilogue is swizzled
  float* bias_s = reinterpret_cast<float*>(epi_stage + EPI_WARPS * EPI_WARP_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * p.BN;

  // producer state (thread 0): the first STG k-blocks are issued BEFORE the CTA-wide set-up barrier -- the ring is empty, so
  // nothing has to be waited for, and the loads fly while TMEM is allocated and the bias is staged
  int pr_s = 0, pr_mt = blockIdx.x, pr_kb = 0;
  uint32_t pr_ph = 0;
  if (warp == 0 && lane == 0) {
    // the loads go out first: they need nothing but their own completion barriers
    for (int s = 0; s < STG; ++s) mbar_init(full + s, 1);
    mbar_init(wfull, 1);
    fence_barrier_init();
    pdl_wait();
    if (p.w_res) {
      mbar_expect_tx(wfull, (uint32_t)(kblocks * w_bytes));
      for (int kb = 0; kb < kblocks; ++kb) tma_load_2d(w_res_base + kb * w_bytes, &map_w, wfull, kb * BK, n0);
    }
    for (int n = 0; n < STG && pr_mt < p.m_tiles; ++n) {
      uint8_t* sa = ring + pr_s * stage_bytes;
      mbar_expect_tx(full + pr_s, (uint32_t)stage_bytes);
      tma_load_2d(sa, &map_a, full + pr_s, pr_kb * BK, pr_mt * BM);
      if (!p.w_res) tma_load_2d(sa + a_bytes, &map_w, full + pr_s, pr_kb * BK, n0);
      if (++pr_s == STG) { pr_s = 0; pr_ph ^= 1; }
      if (++pr_kb == kblocks) { pr_kb = 0; pr_mt += gridDim.x; }
    }
    TR(12);
    for (int s = 0; s < STG; ++s) mbar_init(empty + s, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(tfull + a, 1); mbar_init(tempty + a, EPI_WARPS); }
    fence_barrier_init();
    if (WIDE && p.tma_store) tma_prefetch_desc(&map_c);
    TR(1);
  } else {
    if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
    pdl_wait();
  }
  {   // the bias is staged pre-multiplied by scale_pre so the epilogue is one FFMA per element
    const float sp0 = p.e.scale_pre ? *p.e.scale_pre : 1.f;
    for (int j = threadIdx.x; j < p.BN; j += NTHREADS) bias_s[j] = (p.e.bias && n0 + j < p.N) ? p.e.bias[n0 + j] * sp0 : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 32) TR(2);

  if (warp == 0) {
    // ================= TMA producer
    if (lane == 0) {
      int s = pr_s, kb = pr_kb;
      uint32_t ph = pr_ph;
      for (int mt = pr_mt; mt < p.m_tiles; mt += gridDim.x) {
        for (; kb < kblocks; ++kb) {
          mbar_wait(empty + s, ph ^ 1);
          uint8_t* sa = ring + s * stage_bytes;
          mbar_expect_tx(full + s, (uint32_t)stage_bytes);
          tma_load_2d(sa, &map_a, full + s, kb * BK, mt * BM);
          if (!p.w_res) tma_load_2d(sa + a_bytes, &map_w, full + s, kb * BK, n0);
          if (++s == STG) { s = 0; ph ^= 1; }
        }
        kb = 0;
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BM, p.BN, 0, 0);
      int s = 0, acc = 0;
      uint32_t ph = 0, aph = 0;
      if (p.w_res) mbar_wait(wfull, 0);
      for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x) {
        mbar_wait(tempty + acc, aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full + s, ph);
          if (mt == blockIdx.x && kb == 0) TR(3);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + s * stage_bytes);
          const uint32_t sw = p.w_res ? smem_u32(w_res_base + kb * w_bytes) : sa + a_bytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = make_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t bdesc = make_smem_desc(sw + k * 32, 16, 1024);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty + s);
          if (kb == kblocks - 1) umma_commit(tfull + acc);
          if (++s == STG) { s = 0; ph ^= 1; }
        }
        if (mt == blockIdx.x) TR(4);
        if (mt + gridDim.x >= p.m_tiles) TR(7);
        if (++acc == 2) { acc = 0; aph ^= 1; }
      }
    }
  } else {
    // ================= epilogue warps: TMEM lane quarter = warp % 4, column groups dealt round-robin to the 3 warps
    // of a quarter
    const int quarter = warp & 3, sub = (warp - 2) >> 2;
    const GemmEpi& e = p.e;
    const float s_pre = e.scale_pre ? *e.scale_pre : 1.f;
    const float s_res = e.scale_res ? *e.scale_res : 1.f;
    uint8_t* st = epi_stage + (warp - 2) * EPI_WARP_BYTES;
    uint8_t* my_row = st + lane * EPI_ROWB;
    const bool use_resid = !WIDE && e.C2 && !e.gelu && e.resid;
    const bool use_gmul = !WIDE && !use_resid && e.gmul != nullptr;
    const void* aux = use_resid ? e.resid : e.gmul;
    const long aux_ld = use_resid ? e.ldr : e.ldg;
    const bool aux_f32 = use_resid ? !e.r_bf16 : !e.g_bf16;
    const int ncols = min(p.BN, p.N - n0);
    const bool masked = e.drop.p > 0.f, scaled = e.rowscale != nullptr;   // dropout site / DropPath scale on the output
    DropState dst{};
    if (masked) dst = drop_state(e.drop);
    int ngroups = 0;                                         // column groups of the wide epilogue: 64 | 32 | 16 wide
    for (int g = 0; g < ncols; ++ngroups) { const int rem = ncols - g; g += rem >= 64 ? 64 : (rem >= 32 ? 32 : 16); }
    int acc = 0;
    uint32_t aph = 0;
    for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x) {
      mbar_wait(tfull + acc, aph);
      if (threadIdx.x == 64) { if (mt == blockIdx.x) TR(5); if (mt + gridDim.x >= p.m_tiles) TR(8); }
      tc_fence_after();
      const long row0 = (long)mt * BM + quarter * 32;
      const long my_r = row0 + lane;                                      // this thread's output row
      const float rsc = scaled ? e.rowscale[min(my_r, (long)p.M - 1) / e.rows_per_img] : 1.f;
      if (WIDE) {
        // ---- plain bf16 output (most launches of a step): 64 columns per pass.  A staged row is then 128 B, so the flush
        // reads whole rows per quarter-warp (no bank conflicts; the 32-column passes were 6-way conflicted, ncu) and writes
        // 128 B contiguous per row, with half the passes / warp syncs.
        int slot = 0, gw = 0;
        bool released = false;
        for (int g = 0; g < ncols; g += gw, ++slot) {
          const int rem = ncols - g;
          gw = rem >= 64 ? 64 : (rem >= 32 ? 32 : 16);
          if (slot % (EPI_WARPS / 4) != sub) continue;
          const int nch = gw / 16;
          const bool via_tma = p.tma_store && gw == 64;
          // staging: [32 rows][128 B] swizzled like a SWIZZLE_128B box (16 B chunk j of row r at j ^ (r & 7)) for the TMA
          // store, or the padded row-major buffer of the flush loops
          uint8_t* tb = st + ((warp & 1) ? 512 : 0);              // warp regions are 4608 B apart: this is 1024-aligned
          if (via_tma) {
            if (lane == 0) tma_store_wait_read();                  // the previous store has finished reading the buffer
            __syncwarp();
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {            // two 16-column chunks in flight at a time (register budget)
            if (2 * h >= nch) break;
            float v[2][16];
#pragma unroll
            for (int c = 0; c < 2; ++c)
              if (2 * h + c < nch)
                tmem_ld16_nowait(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * p.BN + g + (2 * h + c) * 16), v[c]);
#pragma unroll
            for (int c = 0; c < 2; ++c)
              if (2 * h + c < nch) {
                const int cc = 2 * h + c;
                tmem_ld_wait16(v[c]);
                const float4* bp = reinterpret_cast<const float4*>(bias_s + g + cc * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float4 b = bp[q];
                  v[c][4 * q] = fmaf(v[c][4 * q], s_pre, b.x); v[c][4 * q + 1] = fmaf(v[c][4 * q + 1], s_pre, b.y);
                  v[c][4 * q + 2] = fmaf(v[c][4 * q + 2], s_pre, b.z); v[c][4 * q + 3] = fmaf(v[c][4 * q + 3], s_pre, b.w);
                }
                if (masked) {
                  const unsigned long long id8 = (unsigned long long)(my_r * p.N + n0 + g + cc * 16) >> 3;
#pragma unroll
                  for (int hh = 0; hh < 2; ++hh) {
                    float k[8];
                    drop_keep8(dst, id8 + hh, k);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[c][hh * 8 + j] *= k[j] * rsc;
                  }
                } else if (scaled) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) v[c][j] *= rsc;
                }
                if (via_tma) {
                  uint32_t pk[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    __nv_bfloat162 hv = __floats2bfloat162_rn(v[c][2 * j], v[c][2 * j + 1]);
                    pk[j] = *reinterpret_cast<uint32_t*>(&hv);
                  }
                  uint8_t* rowp = tb + lane * 128;
                  *reinterpret_cast<uint4*>(rowp + (((2 * cc) ^ (lane & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                  *reinterpret_cast<uint4*>(rowp + (((2 * cc + 1) ^ (lane & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                } else {
                  stage_store16(my_row, cc, v[c], false);
                }
              }
          }
          if (slot + EPI_WARPS / 4 >= ngroups) {   // this warp's last TMEM read of the tile: hand the accumulator back now,
            tc_fence_before();                     // before the stores (the MMA of tile + 2 can start during them)
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + acc);
            released = true;
          }
          if (via_tma) {
            fence_proxy_async();                   // generic-proxy smem writes -> visible to the TMA engine
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&map_c, tb, n0 + g, (int)row0);
              tma_store_commit();
            }
          } else {
            __syncwarp();
            if (gw == 64) stage_flush_bf16<8>(st, e.C, e.ldc, row0, p.M, n0 + g, lane);
            else if (gw == 32) stage_flush_bf16<4>(st, e.C, e.ldc, row0, p.M, n0 + g, lane);
            else stage_flush_bf16<2>(st, e.C, e.ldc, row0, p.M, n0 + g, lane);
            __syncwarp();
          }
        }
        if (!released) {                           // a warp with no column group in this tile still has to arrive
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty + acc);
        }
        if (threadIdx.x == 64) { if (mt == blockIdx.x) TR(6); if (mt + gridDim.x >= p.m_tiles) TR(9); }
        if (++acc == 2) { acc = 0; aph ^= 1; }
        continue;
      }
      for (int g = sub * EPI_GROUP; g < ncols; g += (EPI_WARPS / 4) * EPI_GROUP) {
        const int gw = min(EPI_GROUP, ncols - g), nch = gw / 16;
        float v[2][16], r[2][16];
#pragma unroll
        for (int c = 0; c < 2; ++c)
          if (c < nch) tmem_ld16_nowait(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * p.BN + g + c * 16), v[c]);
        if (use_resid || use_gmul) {
          if (aux_f32) stage_fill_f32(st, static_cast<const float*>(aux), aux_ld, row0, p.M, n0 + g, gw, lane);
          else stage_fill_bf16(st, static_cast<const bf16*>(aux), aux_ld, row0, p.M, n0 + g, gw, lane);
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 2; ++c)
            if (c < nch) stage_read16(my_row, c, aux_f32, r[c]);
          __syncwarp();
        }
#pragma unroll
        for (int c = 0; c < 2; ++c)
          if (c < nch) {
            tmem_ld_wait16(v[c]);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[c][j] = fmaf(v[c][j], s_pre, bias_s[g + c * 16 + j]);   // bias staged pre-scaled
            if (use_gmul) {
              if (e.gmul_raw) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[c][j] *= r[c][j];
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[c][j] *= gelu_grad_fast_f(r[c][j]);
              }
            }
            if ((masked || scaled) && !e.gelu) {   // ids of drop_rows on the contiguous [M, N] output
              float k[16];
              if (masked) {
                const unsigned long long id8 = (unsigned long long)(my_r * p.N + n0 + g + c * 16) >> 3;
                drop_keep8(dst, id8, k);
                drop_keep8(dst, id8 + 1, k + 8);
              }
#pragma unroll
              for (int j = 0; j < 16; ++j) v[c][j] *= (masked ? k[j] : 1.f) * rsc;
            }
          }
        const bool dual_dg = e.gelu && e.gelu_dgrad && e.C && e.C2;   // C = gelu'(pre), C2 = gelu(pre): one tanh for both
        if (dual_dg) {
#pragma unroll
          for (int c = 0; c < 2; ++c)
            if (c < nch) {
#pragma unroll
              for (int j = 0; j < 16; ++j) gelu_both_fast_f(v[c][j], v[c][j], r[c][j]);   // v <- gelu, r <- gelu'
              stage_store16(my_row, c, r[c], e.c_f32 != 0);
            }
          __syncwarp();
          stage_flush_any(st, e.C, e.ldc, e.c_f32 != 0, false, row0, p.M, n0 + g, gw, lane);
          __syncwarp();
        } else if (e.C) {
#pragma unroll
          for (int c = 0; c < 2; ++c)
            if (c < nch) stage_store16(my_row, c, v[c], e.c_f32 != 0);
          __syncwarp();
          stage_flush_any(st, e.C, e.ldc, e.c_f32 != 0, e.c_accum != 0, row0, p.M, n0 + g, gw, lane);
          __syncwarp();
        }
        if (e.C2) {
#pragma unroll
          for (int c = 0; c < 2; ++c)
            if (c < nch) {
              if (e.gelu) {
                if (!dual_dg) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) v[c][j] = gelu_fast_f(v[c][j]);
                }
                if (masked) {                      // nn.Dropout after the activation (H:654): only C2 is dropped
                  float k[16];
                  const unsigned long long id8 = (unsigned long long)(my_r * p.N + n0 + g + c * 16) >> 3;
                  drop_keep8(dst, id8, k);
                  drop_keep8(dst, id8 + 1, k + 8);
#pragma unroll
                  for (int j = 0; j < 16; ++j) v[c][j] *= k[j];
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[c][j] = s_res * v[c][j] + (use_resid ? r[c][j] : 0.f);
              }
              stage_store16(my_row, c, v[c], e.c2_f32 != 0);
            }
          __syncwarp();
          stage_flush_any(st, e.C2, e.ldc2, e.c2_f32 != 0, false, row0, p.M, n0 + g, gw, lane);
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + acc);
      if (++acc == 2) { acc = 0; aph ^= 1; }
    }
  }
  if (WIDE && warp >= 2 && lane == 0) tma_store_wait_all();   // outstanding TMA stores of this warp
  if (threadIdx.x == 64) TR(10);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
  if (threadIdx.x == 0) TR(11);
}

// ------------------------------------------------------------------------------------------------ dW kernel
struct TnArgs {
  int M, N, K;         // dW[N, K], contraction over M rows
  int KD;              // K rounded up to 16 (UMMA N)
  int kboxes;          // ceil(K / 64)
  int mblocks;         // ceil(M / 64)
  int mb_per_cta;
  uint32_t tmem_cols;
  float* dW;
  const float* scale;
  int transposed;      // write dW[k * ldw + n] instead of dW[n * K + k]
  int ldw;
  float* db;           // optional bias gradient: column sums of dY, accumulated by the otherwise idle epilogue warps
  int db_mode;         // 1: dY is the A operand (columns n0 .. n0+127 of this CTA); 2: dY is the B operand (CTAs with n0 == 0)
  int db_cols;         // number of dY columns
  TnSegs segs;         // segs.n > 0: rows of the product are scattered to per-segment dW / db (non-transposed flavour, db_mode 1)
};
__device__ __forceinline__ int tn_seg_of(const TnSegs& g, int n) {
  int i = 0;
  while (i + 1 < g.n && n >= g.end[i]) ++i;
  return i;
}

__global__ void __launch_bounds__(TN_THREADS, 1)
tc_gemm_tn_kernel(const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_x, TnArgs p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int BOX = 64 * 64 * 2;                    // one [64 rows x 128 B] box
  const int a_bytes = 2 * BOX, b_bytes = p.kboxes * BOX, stage_bytes = a_bytes + b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  uint64_t *full = bars, *empty = bars + STAGES, *tfull = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BM;
  const int mb0 = blockIdx.y * p.mb_per_cta;
  const int mb1 = min(p.mblocks, mb0 + p.mb_per_cta);
  const int nmb = mb1 - mb0;

  const bool do_db = (p.db != nullptr || p.segs.n > 0) && (p.db_mode == 1 || blockIdx.x == 0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_y);
    tma_prefetch_desc(&map_x);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, do_db ? 5 : 1); }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (nmb > 0) {
    if (warp == 0) {
      if (lane == 0) {
        int s = 0;
        uint32_t ph = 0;
        for (int mb = mb0; mb < mb1; ++mb) {
          mbar_wait(empty + s, ph ^ 1);
          uint8_t* sa = smem + s * stage_bytes;
          mbar_expect_tx(full + s, (uint32_t)stage_bytes);
          tma_load_2d(sa, &map_y, full + s, n0, mb * 64);
          tma_load_2d(sa + BOX, &map_y, full + s, n0 + 64, mb * 64);
          for (int kx = 0; kx < p.kboxes; ++kx) tma_load_2d(sa + a_bytes + kx * BOX, &map_x, full + s, kx * 64, mb * 64);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = make_idesc(BM, p.KD, 1, 1);
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < nmb; ++i) {
          mbar_wait(full + s, ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * stage_bytes);
          const uint32_t sb = sa + a_bytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {   // 64 contraction rows = 4 x UMMA_K(16); 16 rows x 128 B = 2048 B apart
            const uint64_t adesc = make_smem_desc(sa + k * 2048, BOX, 1024);
            const uint64_t bdesc = make_smem_desc(sb + k * 2048, BOX, 1024);
            umma_bf16(tmem_base, adesc, bdesc, idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty + s);
          if (i == nmb - 1) umma_commit(tfull);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    } else {
      const int quarter = warp & 3;
      const float sc = p.scale ? *p.scale : 1.f;
      if (do_db) {
        // bias gradient while the MMAs run.  Work item = (column pair, row half): one 32-bit shared load covers two
        // adjacent bf16 columns, 32 rows per item with four independent accumulator chains.
        // Box layout (SWIZZLE_128B): row r at r * 128 B, 16 B chunk j stored at chunk j ^ (r & 7).
        const int t = threadIdx.x - 64;
        const int ncol = p.db_mode == 1 ? 128 : p.kboxes * 64;
        const int nitems = ncol;                       // (ncol / 2 pairs) x 2 row halves
        float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};    // [item slot][column of the pair]
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < nmb; ++i) {
          mbar_wait(full + s, ph);
          const uint8_t* tile = smem + s * stage_bytes + (p.db_mode == 1 ? 0 : a_bytes);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int item = t + h * 128;
            if (item < nitems) {
              const int pair = item % (ncol / 2), r0 = (item / (ncol / 2)) * 32;
              const int col = pair * 2;
              const uint8_t* box = tile + (col >> 6) * BOX;
              const int j = (col & 63) >> 3, o = (col & 7) * 2;
              float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int rr = 0; rr < 32; ++rr) {
                const int r = r0 + rr;
                const uint32_t w = *reinterpret_cast<const uint32_t*>(box + r * 128 + ((j ^ (r & 7)) << 4) + o);
                a0[rr & 3] += __uint_as_float(w << 16);
                a1[rr & 3] += __uint_as_float(w & 0xFFFF0000u);
              }
              acc[h][0] += (a0[0] + a0[1]) + (a0[2] + a0[3]);
              acc[h][1] += (a1[0] + a1[1]) + (a1[2] + a1[3]);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(empty + s);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        const int cbase = p.db_mode == 1 ? n0 : 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int item = t + h * 128;
          if (item < nitems) {
            const int col = cbase + (item % (ncol / 2)) * 2;
            if (p.segs.n > 0) {
#pragma unroll
              for (int e2 = 0; e2 < 2; ++e2)
                if (col + e2 < p.db_cols) {
                  const int sg = tn_seg_of(p.segs, col + e2);
                  if (p.segs.db[sg]) atomicAdd(p.segs.db[sg] + (col + e2 - (sg ? p.segs.end[sg - 1] : 0)), acc[h][e2] * sc);
                }
            } else {
              if (col < p.db_cols) atomicAdd(p.db + col, acc[h][0] * sc);
              if (col + 1 < p.db_cols) atomicAdd(p.db + col + 1, acc[h][1] * sc);
            }
          }
        }
      }
      mbar_wait(tfull, 0);
      tc_fence_after();
      const int n = n0 + quarter * 32 + lane;
      for (int c0 = 0; c0 < p.KD; c0 += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
        if (p.transposed) {
          if (n < p.N) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < p.K) atomicAdd(p.dW + (long)(c0 + j) * p.ldw + n, v[j] * sc);   // lanes = consecutive n: coalesced
          }
        } else if (n < p.N) {
          float* dst = p.dW + (long)n * p.K + c0;
          if (p.segs.n > 0) {
            const int sg = tn_seg_of(p.segs, n);
            dst = p.segs.dW[sg] + (long)(n - (sg ? p.segs.end[sg - 1] : 0)) * p.K + c0;
          }
          if (c0 + 16 <= p.K && (p.K & 3) == 0) {      // 16 B vector reductions: 4x fewer RED instructions
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(v[j] * sc), "f"(v[j + 1] * sc),
                           "f"(v[j + 2] * sc), "f"(v[j + 3] * sc)
                           : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < p.K) atomicAdd(dst + j, v[j] * sc);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 row-major [rows, cols] with row stride ld (elements); box = [box_rows x 64 cols], SWIZZLE_128B, zero OOB fill
int make_map(CUtensorMap* map, const bf16* base, long rows, long cols, long ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  QV_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  QV_CHECK(((uintptr_t)base & 15) == 0 && (ld * 2) % 16 == 0, "TMA operand must be 16 B aligned (ptr %p, ld %ld)", (const void*)base, ld);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  QV_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%ld cols=%ld ld=%ld box_rows=%d", (int)r, rows, cols, ld, box_rows);
  return 0;
}

uint32_t pow2_cols(int n) {
  uint32_t c = 32;
  while ((int)c < n) c <<= 1;
  return c;
}

int pick_bn(int N) {
  // largest multiple of 16 that is <= 256 and splits N into equal-ish slices
  if (N <= 256) return ((N + 15) / 16) * 16;
  for (int s = 2; s <= 8; ++s) {
    if (N % s == 0 && (N / s) % 16 == 0 && N / s <= 256) return N / s;
  }
  return 256;
}

}  // namespace

#ifdef QV_GEMM_TRACE
extern "C" int qavit_test_gemm_trace_read(long long* host128) {
  return cudaMemcpyFromSymbol(host128, g_trace, sizeof(long long) * 128) == cudaSuccess ? 0 : 1;
}
#endif

bool tc_shape_ok_nt(int M, int N, int K, int lda) {
  return M >= 1 && N >= 16 && N % 16 == 0 && K >= 8 && K % 8 == 0 && lda % 8 == 0;
}
bool tc_shape_ok_tn(int M, int N, int K, int ldy, int ldx) {
  return M >= 1 && N >= 8 && K >= 8 && K <= 256 && N % 8 == 0 && K % 8 == 0 && ldy % 8 == 0 && ldx % 8 == 0;
}

int tc_gemm_nt(cudaStream_t s, const bf16* A, int lda, int M, int N, int K, const bf16* Wb, const GemmEpi& e) {
  if (M <= 0) return 0;
  QV_CHECK(tc_shape_ok_nt(M, N, K, lda), "tc_gemm_nt: unsupported shape M=%d N=%d K=%d lda=%d", M, N, K, lda);
  QV_CHECK(!(e.c_accum && !e.c_f32), "gemm: accumulate needs an fp32 C");
  NtArgs p{};
  p.M = M; p.N = N; p.K = K;
  p.BN = pick_bn(N);
  p.n_slices = cdiv(N, p.BN);
  p.m_tiles = cdiv(M, BM);
  p.tmem_cols = pow2_cols(2 * p.BN);
  p.e = e;
  QV_CHECK(p.tmem_cols <= 512, "tc_gemm_nt: BN=%d needs %u TMEM columns", p.BN, p.tmem_cols);
  CUtensorMap ma, mw;
  QV_TRY(make_map(&ma, A, M, K, lda, BM));
  QV_TRY(make_map(&mw, Wb, N, K, K, p.BN));
  QV_CHECK((!e.C || (e.ldc % (e.c_f32 ? 4 : 8) == 0 && ((uintptr_t)e.C & 15) == 0)) &&
               (!e.C2 || (e.ldc2 % (e.c2_f32 ? 4 : 8) == 0 && ((uintptr_t)e.C2 & 15) == 0)) &&
               (!e.resid || (e.ldr % (e.r_bf16 ? 8 : 4) == 0 && ((uintptr_t)e.resid & 15) == 0)) &&
               (!e.gmul || (e.ldg % (e.g_bf16 ? 8 : 4) == 0 && ((uintptr_t)e.gmul & 15) == 0)),
           "tc_gemm_nt: outputs / residual must be 16 B aligned with 16 B-multiple row pitch");
  p.stages = p.BN > 208 ? 3 : STAGES;
  const size_t fixed = 1024 + 1024 + EPI_WARPS * EPI_WARP_BYTES + 1024;      // alignment slack, barriers, epilogue staging, bias
  size_t smem = fixed + (size_t)p.stages * (BM * BK * 2 + p.BN * BK * 2);
  {
    // resident weight slice: worth it when a CTA streams more than one m-tile and the slice leaves room for >= 4 A stages
    static const int force_res = [] { const char* v = getenv("QV_W_RES"); return v ? atoi(v) : -1; }();
    const size_t wres = (size_t)cdiv(K, BK) * p.BN * BK * 2;
    const size_t cap = 227 * 1024;
    const int gx0 = max(1, min(p.m_tiles, qv_num_sms() / p.n_slices));
    if (force_res != 0 && wres + fixed + 4 * (size_t)(BM * BK * 2) <= cap && p.m_tiles > gx0) {
      p.w_res = 1;
      p.stages = (int)min((size_t)8, (cap - fixed - wres) / (size_t)(BM * BK * 2));
      smem = fixed + wres + (size_t)p.stages * (BM * BK * 2);
    }
  }
  QV_CHECK(smem <= 227 * 1024, "tc_gemm_nt: BN=%d needs %zu B of shared memory", p.BN, smem);
  // plain bf16 output -> the 64-column epilogue
  const bool wide = e.C && !e.C2 && !e.c_f32 && !e.c_accum && !e.gmul;
  auto kern = wide ? tc_gemm_nt_kernel<true> : tc_gemm_nt_kernel<false>;
  CUtensorMap mc = ma;
  // TMA-store epilogue only where a CTA streams enough tiles to amortise the store drain at kernel end (measured: the
  // 12-tile qkv projection gains 6 %, the 4-tile N = 192 projections lose 6 %); QV_TMA_STORE=0/1 forces it for A/B runs
  static const int force = [] { const char* v = getenv("QV_TMA_STORE"); return v ? atoi(v) : -1; }();
  const int tiles_per_cta = p.m_tiles * p.n_slices / max(1, min(p.m_tiles * p.n_slices, qv_num_sms()));
  p.tma_store = wide && N >= 64 && (force >= 0 ? force != 0 : tiles_per_cta >= 8);
  if (p.tma_store) QV_TRY(make_map(&mc, static_cast<const bf16*>(e.C), M, N, e.ldc, 32));   // box = 32 rows x 64 columns
  QV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int gx = max(1, min(p.m_tiles, qv_num_sms() / p.n_slices));
  qv_launch(kern, dim3(gx, p.n_slices), NTHREADS, smem, s, ma, mw, mc, p);
  QV_LAUNCH_CHECK();
  return 0;
}

static int tc_gemm_tn_impl(cudaStream_t s, const bf16* dY, int ldy, const bf16* X, int ldx, int M, int N, int K, float* dW,
                           const float* scale, int transposed, int ldw, float* db, int db_mode, int db_cols, const TnSegs* segs);
int tc_gemm_tn(cudaStream_t s, const bf16* dY, int ldy, const bf16* X, int ldx, int M, int N, int K, float* dW,
               const float* scale, int transposed, int ldw, float* db, int db_mode, int db_cols) {
  return tc_gemm_tn_impl(s, dY, ldy, X, ldx, M, N, K, dW, scale, transposed, ldw, db, db_mode, db_cols, nullptr);
}
int tc_gemm_tn_seg(cudaStream_t s, const bf16* dY, int ldy, const bf16* X, int ldx, int M, int N, int K, const TnSegs& segs) {
  QV_CHECK(segs.n >= 1 && segs.n <= 4 && segs.end[segs.n - 1] == N, "tc_gemm_tn_seg: segments must cover the %d rows", N);
  return tc_gemm_tn_impl(s, dY, ldy, X, ldx, M, N, K, segs.dW[0], nullptr, 0, 0, nullptr, 1, N, &segs);
}
static int tc_gemm_tn_impl(cudaStream_t s, const bf16* dY, int ldy, const bf16* X, int ldx, int M, int N, int K, float* dW,
                           const float* scale, int transposed, int ldw, float* db, int db_mode, int db_cols, const TnSegs* segs) {
  if (M <= 0) return 0;
  QV_CHECK(tc_shape_ok_tn(M, N, K, ldy, ldx), "tc_gemm_tn: unsupported shape M=%d N=%d K=%d", M, N, K);
  TnArgs p{};
  p.M = M; p.N = N; p.K = K;
  p.KD = ((K + 15) / 16) * 16;
  p.kboxes = cdiv(K, 64);
  p.mblocks = cdiv(M, 64);
  p.tmem_cols = pow2_cols(p.KD);
  p.dW = dW;
  p.scale = scale;
  p.transposed = transposed;
  p.ldw = ldw;
  p.db = db;
  p.db_mode = db_mode;
  p.db_cols = db_cols;
  if (segs) p.segs = *segs;
  const int n_tiles = cdiv(N, BM);
  // Every CTA ends with 128 x K fp32 atomics, so a CTA must own enough rows to amortise them: >= 16 m-blocks
  // (1024 rows) each, and no more CTAs than SMs.
  int splits = max(1, min(cdiv(p.mblocks, 16), qv_num_sms() / n_tiles));
  p.mb_per_cta = cdiv(p.mblocks, splits);
  splits = cdiv(p.mblocks, p.mb_per_cta);
  CUtensorMap my, mx;
  QV_TRY(make_map(&my, dY, M, N, ldy, 64));
  QV_TRY(make_map(&mx, X, M, K, ldx, 64));
  const size_t smem = 1024 + (size_t)STAGES * (2 + p.kboxes) * (64 * 64 * 2) + 256;
  QV_CHECK(smem <= 227 * 1024, "tc_gemm_tn: K=%d needs %zu B smem", K, smem);
  QV_CUDA(cudaFuncSetAttribute(tc_gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  qv_launch(tc_gemm_tn_kernel, dim3(n_tiles, splits), TN_THREADS, smem, s, my, mx, p);
  QV_LAUNCH_CHECK();
  return 0;
}
