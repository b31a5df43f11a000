// bf16 tensor-core (tcgen05 / TMEM) path for the per-node projections of the FeaSt conv and the FC heads.
//
// Building block: D[128 x N] (fp32, TMEM) += A[128 x 64] . B[N x 64]^T with bf16 operands staged in shared memory
// by the CTA's own threads in the canonical K-major SWIZZLE_128B layout (rows of 128 B, 8-row groups of 1 KB,
// 16-byte chunks XOR-ed with row%8), issued by one thread as four K=16 tcgen05.mma, completion tracked with
// tcgen05.commit -> mbarrier, accumulator read back with tcgen05.ld (32 lanes x 32 columns per warp).
// Operands come from fp32 global memory and are rounded to bf16 on the way into shared memory, so no bf16 copy of
// the activations ever exists in HBM.  No TMA: the A operand is produced by threads (converted / aggregated).
#include <cuda.h>
#include <stdlib.h>

#include "tc.cuh"

namespace geobi {
namespace tc {

// ------------------------------------------------------------------------------ weight preparation
// Bq[n, k] (bf16, row stride kpad, zero padded).  mode 0: plain W[n, k];  mode 1: FeaSt lin.weight -> W_flat,
// Bq[o, h*C_in + c] = W[(h*C_out + o)*C_in + c].
// Bq holds two planes: hi at [0, N*kpad), lo (residual) at [N*kpad, 2*N*kpad).
__global__ void prep_weight_kernel(const float* __restrict__ W, int N, int K, int kpad, int mode, int c_in, __nv_bfloat16* __restrict__ Bq) {
  const int total = N * kpad;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int n = t / kpad, k = t - n * kpad;
    float v = 0.f;
    if (k < K) {
      if (mode == 0) v = W[(int64_t)n * K + k];
      else {
        const int h = k / c_in, c = k - h * c_in;
        v = W[(int64_t)(h * N + n) * c_in + c];
      }
    }
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    Bq[t] = hi;
    Bq[total + t] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

int prep_weight(const float* W, int N, int K, int kpad, int mode, int c_in, __nv_bfloat16* Bq, cudaStream_t st) {
  prep_weight_kernel<<<64, 256, 0, st>>>(W, N, K, kpad, mode, c_in, Bq);
  GEOBI_LAUNCH_OK("prep_weight");
  return GEOBI_OK;
}

// ------------------------------------------------------------------------------ out = act(A . Bq^T + bias)
// A and Bq are bf16 planes (hi, and lo = residual when PASSES == 3; plane strides a_plane / NT*kpad elements), row stride kpad.
// One CTA = 128 rows x NT columns, 128 threads.  STAGES-deep cp.async ring: 16-byte chunks go straight from global memory into
// the swizzled operand tiles (no registers, no conversion), so the kernel streams A at HBM rate while one thread issues the
// MMAs of the stage that has landed.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;   // src-size 0 -> 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NT, int PASSES, int STAGES>
__global__ void __launch_bounds__(128) tc_gemm_kernel(const __nv_bfloat16* __restrict__ A, int64_t a_plane, int64_t M, int kpad,
                                                      const __nv_bfloat16* __restrict__ Bq, const float* __restrict__ bias, float slope,
                                                      float* __restrict__ out, int64_t ldo) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t mbar[STAGES];
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr int SPLIT = PASSES == 3 ? 2 : 1;
  constexpr int A_BYTES = BM * 128, B_BYTES = NT * 128;
  constexpr int STAGE = SPLIT * (A_BYTES + B_BYTES);   // [A_hi | A_lo | B_hi | B_lo]
  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int64_t b_plane = (int64_t)NT * kpad;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&mbar[s], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, NT < 32 ? 32 : NT);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  constexpr uint32_t idesc = make_idesc(BM, NT);
  const int KB = kpad / BK;

  auto issue_stage = [&](int kb) {
    const uint32_t st = base + (uint32_t)((kb % STAGES) * STAGE);
    // A: 128 rows x 8 chunks (x planes)
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int idx = it * 128 + tid;
      const int r = idx >> 3, ch = idx & 7;
      const int64_t m = m0 + r;
      const bool ok = m < M;
      const __nv_bfloat16* src = A + (ok ? m : 0) * (int64_t)kpad + (int64_t)kb * BK + ch * 8;
      cp_async16(st + sw128_off(r, ch), src, ok);
      if (PASSES == 3) cp_async16(st + A_BYTES + sw128_off(r, ch), src + a_plane, ok);
    }
    for (int idx = tid; idx < NT * 8; idx += 128) {
      const int r = idx >> 3, ch = idx & 7;
      const __nv_bfloat16* src = Bq + (int64_t)r * kpad + (int64_t)kb * BK + ch * 8;
      cp_async16(st + SPLIT * A_BYTES + sw128_off(r, ch), src, true);
      if (PASSES == 3) cp_async16(st + SPLIT * A_BYTES + B_BYTES + sw128_off(r, ch), src + b_plane, true);
    }
  };

  for (int kb = 0; kb < STAGES - 1; ++kb) {
    if (kb < KB) issue_stage(kb);
    cp_async_commit();
  }
  for (int kb = 0; kb < KB; ++kb) {
    const int s = kb % STAGES;
    cp_async_wait<STAGES - 2>();   // this thread's copies of stage kb have landed
    fence_proxy_async();           // ... and are visible to the tensor core's async proxy
    tc_fence_before();
    __syncthreads();               // ... for every thread's copies
    if (tid < 32 && elect_one()) {
      tc_fence_after();
      const uint32_t st = base + (uint32_t)(s * STAGE);
      const uint64_t ah = make_desc(st), bh = make_desc(st + SPLIT * A_BYTES);
#pragma unroll
      for (int k16 = 0; k16 < BK / 16; ++k16) mma_f16(tmem_d, ah + 2 * k16, bh + 2 * k16, idesc, (kb | k16) ? 1u : 0u);
      if (PASSES == 3) {
        const uint64_t al = make_desc(st + A_BYTES), bl = make_desc(st + SPLIT * A_BYTES + B_BYTES);
#pragma unroll
        for (int k16 = 0; k16 < BK / 16; ++k16) mma_f16(tmem_d, ah + 2 * k16, bl + 2 * k16, idesc, 1u);
#pragma unroll
        for (int k16 = 0; k16 < BK / 16; ++k16) mma_f16(tmem_d, al + 2 * k16, bh + 2 * k16, idesc, 1u);
      }
      mma_commit(&mbar[s]);
    }
    // refill the ring: stage kb+STAGES-1 reuses the buffer read by the MMAs of stage kb-1
    const int nk = kb + STAGES - 1;
    if (nk < KB) {
      if (kb >= 1) {
        mbar_wait(&mbar[(kb - 1) % STAGES], (uint32_t)(((kb - 1) / STAGES) & 1));
        tc_fence_after();
      }
      issue_stage(nk);
    }
    cp_async_commit();
  }
  const int last = KB - 1;
  mbar_wait(&mbar[last % STAGES], (uint32_t)((last / STAGES) & 1));
  tc_fence_after();

  // epilogue: thread = output row (TMEM lane); 32 columns at a time
  const int64_t m = m0 + tid;
  const uint32_t lane_addr = tmem_d + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
  for (int c0 = 0; c0 < NT; c0 += 32) {
    float v[32];
    tmem_ld32(lane_addr + (uint32_t)c0, v);
    if (m < M) {
      float* o = out + m * ldo + c0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 r;
        r.x = v[j] + bias[c0 + j];
        r.y = v[j + 1] + bias[c0 + j + 1];
        r.z = v[j + 2] + bias[c0 + j + 2];
        r.w = v[j + 3] + bias[c0 + j + 3];
        r.x = r.x > 0.f ? r.x : r.x * slope;
        r.y = r.y > 0.f ? r.y : r.y * slope;
        r.z = r.z > 0.f ? r.z : r.z * slope;
        r.w = r.w > 0.f ? r.w : r.w * slope;
        *reinterpret_cast<float4*>(o + j) = r;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, NT < 32 ? 32 : NT);
}

// ------------------------------------------------------------------------------ same GEMM, operands moved by TMA
// One elected thread of warp 0 streams the operand tiles with cp.async.bulk.tensor (SWIZZLE_128B tensor maps: the boxes
// land directly in the K-major UMMA layout), one elected thread of warp 1 issues the MMAs; full[] / empty[] mbarriers
// form the ring.  Against the cp.async version this removes ~24 address computations + copies per thread per stage
// (the 128-thread CTA was spending its issue slots on them) and the per-stage __syncthreads.
template <int NT, int PASSES, int STAGES>
__global__ void __launch_bounds__(128) tc_gemm_tma_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                                                          const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                                                          int64_t M, int kpad, const float* __restrict__ bias, float slope,
                                                          float* __restrict__ out, int64_t ldo) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[STAGES];
  __shared__ __align__(8) uint64_t empty[STAGES];
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr int SPLIT = PASSES == 3 ? 2 : 1;
  constexpr int A_BYTES = BM * 128, B_BYTES = NT * 128;
  constexpr int STAGE = SPLIT * (A_BYTES + B_BYTES);   // [A_hi | A_lo | B_hi | B_lo]
  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t m0 = (int64_t)blockIdx.x * BM;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, NT < 32 ? 32 : NT);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  constexpr uint32_t idesc = make_idesc(BM, NT);
  const int KB = kpad / BK;

  if (warp == 0) {
    if (elect_one()) {                                   // ---- producer
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        if (kb >= STAGES) mbar_wait(&empty[s], (uint32_t)(((kb / STAGES) - 1) & 1));   // the MMAs that read this slot are done
        const uint32_t st = base + (uint32_t)(s * STAGE);
        mbar_expect_tx(&full[s], (uint32_t)STAGE);
        tma_load_2d(st, &tm_a_hi, kb * BK, (int)m0, &full[s]);
        if (PASSES == 3) tma_load_2d(st + A_BYTES, &tm_a_lo, kb * BK, (int)m0, &full[s]);
        tma_load_2d(st + SPLIT * A_BYTES, &tm_b_hi, kb * BK, 0, &full[s]);
        if (PASSES == 3) tma_load_2d(st + SPLIT * A_BYTES + B_BYTES, &tm_b_lo, kb * BK, 0, &full[s]);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {                                   // ---- MMA issue
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&full[s], (uint32_t)((kb / STAGES) & 1));
        tc_fence_after();
        const uint32_t st = base + (uint32_t)(s * STAGE);
        const uint64_t ah = make_desc(st), bh = make_desc(st + SPLIT * A_BYTES);
#pragma unroll
        for (int k16 = 0; k16 < BK / 16; ++k16) mma_f16(tmem_d, ah + 2 * k16, bh + 2 * k16, idesc, (kb | k16) ? 1u : 0u);
        if (PASSES == 3) {
          const uint64_t al = make_desc(st + A_BYTES), bl = make_desc(st + SPLIT * A_BYTES + B_BYTES);
#pragma unroll
          for (int k16 = 0; k16 < BK / 16; ++k16) mma_f16(tmem_d, ah + 2 * k16, bl + 2 * k16, idesc, 1u);
#pragma unroll
          for (int k16 = 0; k16 < BK / 16; ++k16) mma_f16(tmem_d, al + 2 * k16, bh + 2 * k16, idesc, 1u);
        }
        mma_commit(&empty[s]);
      }
      mma_commit(&done);
    }
    __syncwarp();
  }
  mbar_wait(&done, 0u);
  tc_fence_after();

  // epilogue: thread = output row (TMEM lane), 32 columns at a time.  The 32 x 32 block of a warp leaves through a 4 KB staging
  // area in the (now dead) operand ring - rows of 128 B, 16-byte chunks XOR-ed with row % 8 - so that 8 lanes write one full
  // 128-byte line of a row; straight from the row-per-thread layout every store instruction touched 32 different lines (the
  // wide-N, short-K products of the backward pass, dZ = g . W_flat, ran at 2.1 TB/s of output).
  const int lane = tid & 31;
  const uint32_t lane_addr = tmem_d + ((uint32_t)(warp * 32) << 16);
  uint8_t* stg = smem_raw + (base - smem_u32(smem_raw)) + warp * 4096;
#pragma unroll 1
  for (int c0 = 0; c0 < NT; c0 += 32) {
    float v[32];
    tmem_ld32(lane_addr + (uint32_t)c0, v);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 r;
      r.x = v[j] + bias[c0 + j];
      r.y = v[j + 1] + bias[c0 + j + 1];
      r.z = v[j + 2] + bias[c0 + j + 2];
      r.w = v[j + 3] + bias[c0 + j + 3];
      r.x = r.x > 0.f ? r.x : r.x * slope;
      r.y = r.y > 0.f ? r.y : r.y * slope;
      r.z = r.z > 0.f ? r.z : r.z * slope;
      r.w = r.w > 0.f ? r.w : r.w * slope;
      *reinterpret_cast<float4*>(stg + lane * 128 + (((j >> 2) ^ (lane & 7)) << 4)) = r;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rr = 4 * i + (lane >> 3), ch = lane & 7;
      const int64_t mr = m0 + warp * 32 + rr;
      if (mr < M) *reinterpret_cast<float4*>(out + mr * ldo + c0 + 4 * ch) = *reinterpret_cast<const float4*>(stg + rr * 128 + ((ch ^ (rr & 7)) << 4));
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, NT < 32 ? 32 : NT);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libgeobi links cudart statically, not libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static const EncodeTiledFn fn = []() -> EncodeTiledFn {      // thread-safe: a `tried` flag set before the pointer handed a second thread nullptr
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    EncodeTiledFn f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      f = reinterpret_cast<EncodeTiledFn>(p);
    cudaGetLastError();
    return f;
  }();
  return fn;
}
// bf16 [rows, kpad] row-major, box = 64 columns (128 B, the swizzle span) x box_rows
bool make_tmap(CUtensorMap* m, const __nv_bfloat16* g, int64_t rows, int kpad, int box_rows) {   // also used by feast_bwd_tc.cu
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn || rows <= 0) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)kpad * sizeof(__nv_bfloat16)};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t es[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(g), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NT, int PASSES>
static int launch_gemm(const __nv_bfloat16* A, int64_t a_plane, int64_t M, int kpad, const __nv_bfloat16* Bq, const float* bias, float slope,
                       float* out, int64_t ldo, cudaStream_t st) {
  constexpr int STAGE_BYTES = (PASSES == 3 ? 2 : 1) * (BM * 128 + NT * 128);
  // as many stages as fit under the 227 KB per-CTA limit (one CTA per SM: the bytes in flight are what hides HBM latency)
  // two CTAs per SM when at least two stages fit in half the shared memory: one CTA's prologue / TMEM drain then overlaps
  // the other's main loop (A/B on the bench: NT=64 530 -> 501 us, NT=32 (K = 128 only) 159 -> 97 us per step)
  constexpr int FIT2 = (110 * 1024) / STAGE_BYTES;
  constexpr int FIT = FIT2 >= 2 ? FIT2 : (220 * 1024) / STAGE_BYTES;
  constexpr int STAGES = FIT > 6 ? 6 : (FIT < 2 ? 2 : FIT);
  const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024;
  if (getenv("GEOBI_GEMM_CPASYNC") == nullptr && M > 0 && (reinterpret_cast<uintptr_t>(A) & 127) == 0 && (reinterpret_cast<uintptr_t>(Bq) & 127) == 0 &&
      (a_plane * 2) % 128 == 0) {
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    const bool split = PASSES == 3;
    bool ok = make_tmap(&ta_hi, A, M, kpad, BM) && make_tmap(&tb_hi, Bq, NT, kpad, NT);
    ok = ok && make_tmap(&ta_lo, split ? A + a_plane : A, M, kpad, BM) && make_tmap(&tb_lo, split ? Bq + (int64_t)NT * kpad : Bq, NT, kpad, NT);
    if (ok) {
      GEOBI_CUDA_OK(cudaFuncSetAttribute(tc_gemm_tma_kernel<NT, PASSES, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      // a product with fewer K blocks than ring slots (dZ = g . W_flat: K = C_out, one or two blocks) only ever touches the first
      // kpad / 64 slots: asking for just those lets two CTAs share an SM where the full ring allows one (NT = 256: 197 -> 99 KB)
      const int kblocks = kpad / BK;
      const size_t smem_used = (size_t)(kblocks < STAGES ? kblocks : STAGES) * STAGE_BYTES + 1024;
      tc_gemm_tma_kernel<NT, PASSES, STAGES><<<(unsigned)cdiv(M, BM), 128, smem_used, st>>>(ta_hi, ta_lo, tb_hi, tb_lo, M, kpad, bias, slope, out, ldo);
      GEOBI_LAUNCH_OK("tc_gemm_tma");
      return GEOBI_OK;
    }
  }
  GEOBI_CUDA_OK(cudaFuncSetAttribute(tc_gemm_kernel<NT, PASSES, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_gemm_kernel<NT, PASSES, STAGES><<<(unsigned)cdiv(M, BM), 128, smem, st>>>(A, a_plane, M, kpad, Bq, bias, slope, out, ldo);
  GEOBI_LAUNCH_OK("tc_gemm");
  return GEOBI_OK;
}

int gemm_dispatch(const __nv_bfloat16* A, int64_t a_plane, int64_t M, int kpad, const __nv_bfloat16* Bq, int N, const float* bias, float slope,
                  float* out, int64_t ldo, int passes, cudaStream_t st) {
#define GEOBI_TC_CASE(NT)                                                                                  \
  case NT:                                                                                                 \
    return passes == 3 ? launch_gemm<NT, 3>(A, a_plane, M, kpad, Bq, bias, slope, out, ldo, st)            \
                       : launch_gemm<NT, 1>(A, a_plane, M, kpad, Bq, bias, slope, out, ldo, st);
  switch (N) {
    GEOBI_TC_CASE(32)
    GEOBI_TC_CASE(64)
    GEOBI_TC_CASE(128)
    GEOBI_TC_CASE(256)
  }
#undef GEOBI_TC_CASE
  set_error("tc gemm: N must be 32, 64, 128 or 256 (got %d)", N);
  return GEOBI_ERR_INVALID;
}

// fp32 [M, lda] -> bf16 planes hi | lo with row stride kpad (zero padded)
__global__ void split_rows_kernel(const float* __restrict__ A, int64_t lda, int64_t M, int K, int kpad, __nv_bfloat16* __restrict__ out) {
  const int64_t total = M * kpad;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = t / kpad;
    const int k = (int)(t - m * kpad);
    const float v = k < K ? A[m * lda + k] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    out[t] = hi;
    out[total + t] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

// ------------------------------------------------------------------------------ fused FC head on tensor cores
// y = W2 . lrelu(W1 . f + b1) + b2 with c_in = 32: GEMM1 [128 nodes x 32] x [32 x 256-chunk] runs on tcgen05 (split bf16, 3 passes,
// hi|lo packed side by side in one 128-byte swizzle row), the 256 hidden activations of a chunk are consumed straight out of TMEM:
// +b1, leaky_relu, dotted with W2 (<= 4 outputs) on CUDA cores.  The [N,1024] hidden never reaches shared memory or HBM.
// 256 threads: warps w and w+4 share a TMEM lane quadrant and split each chunk's columns.
constexpr int FC_CHUNK = 128;         // hidden units per MMA chunk: two row tiles x 128 columns = 256 TMEM columns per CTA (two CTAs per SM)
// W1 [hidden, 32] fp32 -> [hidden, 64] bf16 rows: hi(W1[r, :]) | lo(W1[r, :]) - one 128-byte swizzle row per hidden unit, so a
// 256-row TMA box is a ready B-operand tile (k 0..31 = hi, 32..63 = lo).  Done once per call instead of by every CTA per chunk.
__global__ void fc_w1_split_kernel(const float* __restrict__ W1, int hidden, __nv_bfloat16* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= hidden * 32) return;
  const int r = t >> 5, c = t & 31;
  const float v = W1[t];
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  out[r * 64 + c] = hi;
  out[r * 64 + 32 + c] = __float2bfloat16_rn(v - __bfloat162float(hi));
}
__device__ __forceinline__ unsigned long long pack_u2(uint32_t lo, uint32_t hi) {
  unsigned long long d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// Two 128-row tiles per CTA share every read of the {b1, W2} table: the epilogue is bound by those broadcast LDS.128 (ncu round 2:
// 64 shared-memory wavefronts per node with one tile, data pipe 70 % + 15 % busy), so each table entry now serves two rows per thread.
constexpr int FC_TILES = 2;
__global__ void __launch_bounds__(256) fc_head_tc_kernel(const float* __restrict__ f, int64_t ldf, int64_t N,
                                                         const __grid_constant__ CUtensorMap tm_w1,
                                                         const float* __restrict__ b1, int hidden, const float* __restrict__ W2,
                                                         const float* __restrict__ b2, int CO, int epilogue, const float* __restrict__ res,
                                                         int64_t ldres, const float* __restrict__ res2, int64_t ldres2,
                                                         float* __restrict__ out, int64_t ldo) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ __align__(8) uint64_t w_full[2];          // W1 chunk landed in b_t[0 / 1] (TMA, byte-counted)
  __shared__ uint32_t tmem_slot;
  __shared__ float part[FC_TILES * 128][4];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* a_t = sm;                                   // FC_TILES x [128 rows x 128 B]: k 0..31 = hi, 32..63 = lo
  uint8_t* b_t = sm + FC_TILES * BM * 128;             // 2 x [FC_CHUNK rows x 128 B]: same packing for the W1 chunks (double buffered)
  // [hidden / 2][2] float4: {b1[j], b1[j+1], w2_0[j], w2_0[j+1]}, {w2_1[j], w2_1[j+1], w2_2[j], w2_2[j+1]} - pairs of hidden units
  // so that the epilogue runs on packed f32x2 instructions
  float4* tab = reinterpret_cast<float4*>(sm + FC_TILES * BM * 128 + 2 * FC_CHUNK * 128);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.x * (FC_TILES * BM);

  const int n_chunks = hidden / FC_CHUNK;
  if (tid == 0) {
    mbar_init(&mbar, 1);
    mbar_init(&w_full[0], 1);
    mbar_init(&w_full[1], 1);
    fence_mbar_init();
    for (int c = 0; c < 2 && c < n_chunks; ++c) {      // the first two weight chunks start streaming right away
      mbar_expect_tx(&w_full[c], FC_CHUNK * 128);
      tma_load_2d(smem_u32(b_t + c * FC_CHUNK * 128), &tm_w1, 0, c * FC_CHUNK, &w_full[c]);
    }
  }
  if (warp == 0) tmem_alloc(&tmem_slot, FC_TILES * FC_CHUNK);
  for (int j2 = tid; j2 < hidden / 2; j2 += 256) {
    const int j = 2 * j2;
    tab[2 * j2] = make_float4(b1[j], b1[j + 1], W2[j], W2[j + 1]);
    tab[2 * j2 + 1] = make_float4(CO > 1 ? W2[hidden + j] : 0.f, CO > 1 ? W2[hidden + j + 1] : 0.f, CO > 2 ? W2[2 * hidden + j] : 0.f,
                                  CO > 2 ? W2[2 * hidden + j + 1] : 0.f);
  }
  // A tiles: FC_TILES x 128 rows x 8 float4
  for (int idx = tid; idx < FC_TILES * BM * 8; idx += 256) {
    const int r = idx >> 3, c4 = idx & 7;
    const int64_t m = m0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < N) v = *reinterpret_cast<const float4*>(f + m * ldf + c4 * 4);
    uint2 hi, lo;
    split_bf16x4(v, hi, lo);
    uint8_t* at = a_t + (r >> 7) * (BM * 128);
    *reinterpret_cast<uint2*>(at + sw128_off(r & 127, c4 >> 1) + (c4 & 1) * 8) = hi;
    *reinterpret_cast<uint2*>(at + sw128_off(r & 127, 4 + (c4 >> 1)) + (c4 & 1) * 8) = lo;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  constexpr uint32_t idesc = make_idesc(BM, FC_CHUNK);
  const int row = (warp & 3) * 32 + lane;
  const int chalf = warp >> 2;                         // which half of the chunk's columns this warp consumes
  constexpr int HC = FC_CHUNK / 2;                     // columns per warp and chunk
  const uint32_t lane_addr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(chalf * HC);
  unsigned long long yp[FC_TILES][3];                  // (even, odd) hidden-unit partial sums of the three outputs, per row tile
#pragma unroll
  for (int t = 0; t < FC_TILES; ++t) yp[t][0] = yp[t][1] = yp[t][2] = 0ull;

  for (int ch = 0; ch < n_chunks; ++ch) {
    if (ch == 0) fence_proxy_async();                  // the A tiles were written with generic stores
    tc_fence_before();
    __syncthreads();                                   // every warp has finished reading TMEM of the previous chunk
    if (tid < 32 && elect_one()) {
      mbar_wait(&w_full[ch & 1], (uint32_t)((ch >> 1) & 1));
      tc_fence_after();
      const uint64_t bd = make_desc(smem_u32(b_t + (ch & 1) * FC_CHUNK * 128));
#pragma unroll
      for (int t = 0; t < FC_TILES; ++t) {
        const uint64_t ad = make_desc(smem_u32(a_t + t * (BM * 128)));
        const uint32_t acc = tmem_d + (uint32_t)(t * FC_CHUNK);
        mma_f16(acc, ad + 0, bd + 0, idesc, 0u);       // hi . hi   (k 0..15)
        mma_f16(acc, ad + 2, bd + 2, idesc, 1u);       // hi . hi   (k 16..31)
        mma_f16(acc, ad + 0, bd + 4, idesc, 1u);       // hi . lo
        mma_f16(acc, ad + 2, bd + 6, idesc, 1u);
        mma_f16(acc, ad + 4, bd + 0, idesc, 1u);       // lo . hi
        mma_f16(acc, ad + 6, bd + 2, idesc, 1u);
      }
      mma_commit(&mbar);
    }
    mbar_wait(&mbar, (uint32_t)(ch & 1));
    tc_fence_after();
    if (tid == 0 && ch + 2 < n_chunks) {               // this chunk's MMAs are done with b_t[ch & 1]: refill it with chunk ch + 2
      mbar_expect_tx(&w_full[ch & 1], FC_CHUNK * 128);
      tma_load_2d(smem_u32(b_t + (ch & 1) * FC_CHUNK * 128), &tm_w1, 0, (ch + 2) * FC_CHUNK, &w_full[ch & 1]);
    }
    const ulonglong2* t2 = reinterpret_cast<const ulonglong2*>(tab) + ch * FC_CHUNK + chalf * HC;   // 2 entries per column pair
    // per 16 columns: both tiles' accumulator values are read from TMEM (the next 16 columns' reads are in flight meanwhile), and a
    // column pair costs 2 LDS.128 for BOTH rows + per row add2 + mul2 + 2 max + 3 fma2
    uint32_t va[FC_TILES][16], vb[FC_TILES][16];
    auto consume = [&](const uint32_t (&v)[FC_TILES][16], int col0) {
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const ulonglong2 ta = t2[col0 + j], tb = t2[col0 + j + 1];
#pragma unroll
        for (int t = 0; t < FC_TILES; ++t) {
          unsigned long long h = add2(pack_u2(v[t][j], v[t][j + 1]), ta.x);          // + b1
          const unsigned long long hs = mul2(h, 0x3e4ccccd3e4ccccdull);              // 0.2f, 0.2f
          h = pack_u2(__float_as_uint(fmaxf(__uint_as_float((unsigned)h), __uint_as_float((unsigned)hs))),
                      __float_as_uint(fmaxf(__uint_as_float((unsigned)(h >> 32)), __uint_as_float((unsigned)(hs >> 32)))));
          yp[t][0] = fma2(ta.y, h, yp[t][0]);
          yp[t][1] = fma2(tb.x, h, yp[t][1]);
          yp[t][2] = fma2(tb.y, h, yp[t][2]);
        }
      }
    };
    auto issue = [&](uint32_t (&v)[FC_TILES][16], int col0) {
#pragma unroll
      for (int t = 0; t < FC_TILES; ++t) tmem_ld16_issue(lane_addr + (uint32_t)(t * FC_CHUNK + col0), v[t]);
    };
    issue(va, 0);
#pragma unroll
    for (int q = 0; q < HC / 32; ++q) {
      tmem_ld_wait();
      issue(vb, 32 * q + 16);
      consume(va, 32 * q);
      tmem_ld_wait();
      if (q + 1 < HC / 32) issue(va, 32 * q + 32);
      consume(vb, 32 * q + 16);
    }
    tc_fence_before();
  }
  float y0[FC_TILES], y1[FC_TILES], y2[FC_TILES];
#pragma unroll
  for (int t = 0; t < FC_TILES; ++t) {
    y0[t] = __uint_as_float((unsigned)yp[t][0]) + __uint_as_float((unsigned)(yp[t][0] >> 32));
    y1[t] = __uint_as_float((unsigned)yp[t][1]) + __uint_as_float((unsigned)(yp[t][1] >> 32));
    y2[t] = __uint_as_float((unsigned)yp[t][2]) + __uint_as_float((unsigned)(yp[t][2] >> 32));
    if (chalf == 1) {
      part[t * 128 + row][0] = y0[t];
      part[t * 128 + row][1] = y1[t];
      part[t * 128 + row][2] = y2[t];
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, FC_TILES * FC_CHUNK);
  if (chalf != 0) return;
#pragma unroll
  for (int t = 0; t < FC_TILES; ++t) {
    const int64_t n = m0 + t * 128 + row;
    if (n >= N) continue;
    const float* pt = part[t * 128 + row];
    float y[3] = {y0[t] + pt[0] + b2[0], CO > 1 ? y1[t] + pt[1] + b2[1] : 0.f, CO > 2 ? y2[t] + pt[2] + b2[2] : 0.f};
    int co = CO;
    if (epilogue == 2) {
      if (CO == 1) {
        const float s0 = y[0];
        for (int c = 0; c < 3; ++c) y[c] = s0 * res2[n * ldres2 + c];
        co = 3;
      } else {
        for (int c = 0; c < co; ++c) y[c] *= res2[n * ldres2 + c];
      }
    }
    if (epilogue == 1 || epilogue == 2)
      for (int c = 0; c < co; ++c) y[c] += res[n * ldres + c];
    if (epilogue == 3) {
      float s2 = 0.f;
      for (int c = 0; c < co; ++c) s2 += y[c] * y[c];
      const float d = fmaxf(sqrtf(s2), 1e-12f);
      for (int c = 0; c < co; ++c) y[c] /= d;
    }
    for (int c = 0; c < co; ++c) out[n * ldo + c] = y[c];
  }
}

}  // namespace tc

// ------------------------------------------------------------------------------ FeaSt forward, bf16 projection
// defined in feast.cu
int feast_project_and_aggregate(const float* x, int64_t ldx, int64_t N, int c_in, const int32_t* rowptr, const int32_t* nbr,
                                const int32_t* row_map, int64_t n_src, const float* U, const float* c, double* P, void* Z, int64_t ldz,
                                int out_mode, cudaStream_t st);
bool feast_aggregate_fills_padding(int c_in, int64_t ldx, int64_t ldz, const float* x, bool row_map);

struct TcWs {
  double* P;
  __nv_bfloat16* Z;   // hi | lo planes, [N, kpad] each
  __nv_bfloat16* Bq;
};
template <class C>
static void carve_tc(C& c, int64_t N, int c_in, int c_out, TcWs* out) {
  const int kpad = (int)(cdiv(tc::H * c_in, tc::BK) * tc::BK);
  double* P = c.template take<double>((size_t)N * tc::H);
  __nv_bfloat16* Z = c.template take<__nv_bfloat16>((size_t)2 * N * kpad);
  __nv_bfloat16* Bq = c.template take<__nv_bfloat16>((size_t)2 * c_out * kpad);
  if (out) *out = TcWs{P, Z, Bq};
}
struct NullCarverT {
  Sizer s;
  template <typename T>
  T* take(size_t n) { s.take<T>(n); return nullptr; }
};

size_t feast_fwd_tc_ws_bytes(int64_t N, int c_in, int c_out) {
  NullCarverT c;
  carve_tc(c, N, c_in, c_out, nullptr);
  return c.s.total();
}

int feast_fwd_tc(const float* x, int64_t ldx, int64_t N, int c_in, const int32_t* rowptr, const int32_t* nbr, const int32_t* row_map,
                 int64_t n_src, const float* W, const float* U, const float* c, const float* bias, int c_out, float act_slope, float* out,
                 int64_t ldo, int passes, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!ws || ws_bytes < feast_fwd_tc_ws_bytes(N, c_in, c_out)) {
    set_error("feast_fwd (bf16): workspace too small");
    return GEOBI_ERR_WORKSPACE;
  }
  GEOBI_REQUIRE(ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "feast_fwd (bf16): out rows must be 16-byte aligned");
  Carver cv(ws, ws_bytes);
  TcWs Wk;
  carve_tc(cv, N, c_in, c_out, &Wk);
  const int K = tc::H * c_in;
  const int kpad = (int)(cdiv(K, tc::BK) * tc::BK);
  if (kpad != K && !feast_aggregate_fills_padding(c_in, ldx, kpad, x, row_map != nullptr)) GEOBI_CUDA_OK(cudaMemsetAsync(Wk.Z, 0, sizeof(__nv_bfloat16) * (size_t)(passes == 3 ? 2 : 1) * N * kpad, st));
  tc::prep_weight_kernel<<<64, 256, 0, st>>>(W, c_out, K, kpad, 1, c_in, Wk.Bq);
  int rc = feast_project_and_aggregate(x, ldx, N, c_in, rowptr, nbr, row_map, n_src, U, c, Wk.P, Wk.Z, kpad, passes == 3 ? 2 : 1, st);
  if (rc) return rc;
  return tc::gemm_dispatch(Wk.Z, (int64_t)N * kpad, N, kpad, Wk.Bq, c_out, bias, act_slope, out, ldo, passes, st);
}

size_t fc_head_tc_ws_bytes(int hidden) { return align256((size_t)hidden * 64 * sizeof(__nv_bfloat16)) + 256; }

int fc_head_fwd_tc(const float* f, int64_t ldf, int64_t n, int c_in, const float* W1, const float* b1, int hidden, const float* W2,
                   const float* b2, int c_out, int epilogue, const float* res, int64_t ldres, const float* res2, int64_t ldres2, float* out,
                   int64_t ldo, void* ws, size_t ws_bytes, cudaStream_t st) {
  GEOBI_REQUIRE(c_in == 32, "fc_head_fwd (tensor core): c_in must be 32 (got %d)", c_in);
  GEOBI_REQUIRE(hidden % tc::FC_CHUNK == 0 && hidden <= 4096, "fc_head_fwd (tensor core): hidden must be a multiple of 128, <= 4096");
  GEOBI_REQUIRE(c_out <= 3, "fc_head_fwd (tensor core): c_out must be <= 3 (got %d)", c_out);
  GEOBI_REQUIRE(ldf % 4 == 0 && (reinterpret_cast<uintptr_t>(f) & 15) == 0, "fc_head_fwd (tensor core): feature rows must be 16-byte aligned");
  if (!ws || ws_bytes < fc_head_tc_ws_bytes(hidden) || (reinterpret_cast<uintptr_t>(ws) & 127) != 0) {
    set_error("fc_head_fwd (tensor core): workspace missing, too small or not 128-byte aligned");
    return GEOBI_ERR_WORKSPACE;
  }
  if (n == 0) return GEOBI_OK;
  __nv_bfloat16* w1q = static_cast<__nv_bfloat16*>(ws);
  tc::fc_w1_split_kernel<<<(unsigned)cdiv((int64_t)hidden * 32, 256), 256, 0, st>>>(W1, hidden, w1q);
  CUtensorMap tm;
  if (!tc::make_tmap(&tm, w1q, hidden, 64, tc::FC_CHUNK)) {
    set_error("fc_head_fwd (tensor core): cuTensorMapEncodeTiled unavailable");
    return GEOBI_ERR_CUDA;
  }
  const size_t smem = (size_t)tc::FC_TILES * tc::BM * 128 + 2 * tc::FC_CHUNK * 128 + (size_t)hidden * 16 + 1024;
  GEOBI_CUDA_OK(cudaFuncSetAttribute(tc::fc_head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc::fc_head_tc_kernel<<<(unsigned)cdiv(n, tc::FC_TILES * tc::BM), 256, smem, st>>>(f, ldf, n, tm, b1, hidden, W2, b2, c_out, epilogue, res, ldres, res2,
                                                                     ldres2, out, ldo);
  GEOBI_LAUNCH_OK("fc_head_tc");
  return GEOBI_OK;
}
}  // namespace geobi

// ------------------------------------------------------------------------------ public: tensor-core linear layer
using namespace geobi;

extern "C" size_t geobi_linear_tc_ws_bytes(int64_t m, int k, int n) {
  const size_t kpad = (size_t)(cdiv(k, tc::BK) * tc::BK);
  return align256((size_t)2 * n * kpad * 2) + align256((size_t)2 * (size_t)m * kpad * 2) + 512;
}

extern "C" int geobi_linear_tc(const float* A, int64_t lda, int64_t M, int K, const float* W, int N, const float* bias, float act_slope,
                               float* out, int64_t ldo, int precision, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GEOBI_REQUIRE(A && W && bias && out && M >= 0 && K > 0, "linear_tc: bad arguments");
  GEOBI_REQUIRE(precision == GEOBI_PREC_BF16 || precision == GEOBI_PREC_BF16X3, "linear_tc: precision must be BF16 or BF16X3");
  GEOBI_REQUIRE(ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "linear_tc: out rows must be 16-byte aligned");
  if (!ws || ws_bytes < geobi_linear_tc_ws_bytes(M, K, N)) { set_error("linear_tc: workspace too small"); return GEOBI_ERR_WORKSPACE; }
  if (M == 0) return GEOBI_OK;
  const int kpad = (int)(cdiv(K, tc::BK) * tc::BK);
  Carver cv(ws, ws_bytes);
  __nv_bfloat16* Bq = cv.take<__nv_bfloat16>((size_t)2 * N * kpad);
  __nv_bfloat16* Aq = cv.take<__nv_bfloat16>((size_t)2 * M * kpad);
  tc::prep_weight_kernel<<<64, 256, 0, st>>>(W, N, K, kpad, 0, 0, Bq);
  tc::split_rows_kernel<<<(unsigned)(cdiv(M * kpad, 256) > 148 * 32 ? 148 * 32 : cdiv(M * kpad, 256)), 256, 0, st>>>(A, lda, M, K, kpad, Aq);
  GEOBI_LAUNCH_OK("linear_tc prep");
  return tc::gemm_dispatch(Aq, (int64_t)M * kpad, M, kpad, Bq, N, bias, act_slope, out, ldo, precision == GEOBI_PREC_BF16X3 ? 3 : 1, st);
}
