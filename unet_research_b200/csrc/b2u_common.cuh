// Shared device helpers for the b2u kernels: error plumbing, PTX wrappers for mbarrier / TMA /
// tcgen05 (sm_100a), small vector load/store helpers.  No reference code is used here.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#ifdef __cplusplus
#include <utility>
#endif

#include "../../include/b2u.h"

// ----------------------------------------------------------------------------- host errors
void b2u_set_error(const char* fmt, ...);
#define B2U_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      b2u_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return B2U_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)
#define B2U_REQUIRE(cond, ...)                                                            \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      b2u_set_error(__VA_ARGS__);                                                         \
      return B2U_ERR_ARG;                                                                 \
    }                                                                                     \
  } while (0)
#define B2U_LAUNCH_CHECK()                                                                \
  do {                                                                                    \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      b2u_set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return B2U_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

int b2u_num_sms();
// Opt a kernel into > 48 KB of dynamic shared memory once per (instantiation, device): function attributes are per
// device, a process-wide flag would leave the second GPU of a multi-device process at the default limit.
#define B2U_SET_MAX_SMEM_ONCE(kernel, bytes)                                                          \
  do {                                                                                                \
    static unsigned long long _done = 0;                                                              \
    int _dev = 0;                                                                                     \
    B2U_CHECK_CUDA(cudaGetDevice(&_dev));                                                             \
    if (!((_done >> (_dev & 63)) & 1ull)) {                                                           \
      B2U_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)); \
      _done |= 1ull << (_dev & 63);                                                                   \
    }                                                                                                 \
  } while (0)
int b2u_pdl_enabled();      // 1 unless the environment sets B2U_PDL=0

#ifdef __CUDACC__
// Launch `kernel` with the programmatic-stream-serialization attribute (see pdl_wait / pdl_trigger below).  Every
// kernel launched through this helper calls pdl_wait() before its first dependent global-memory access.
template <typename... KArgs, typename... Args>
inline cudaError_t b2u_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = b2u_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#define B2U_PDL_LAUNCH(kernel, grid, block, smem, st, ...) (void)b2u_launch_pdl((kernel), (grid), (block), (smem), (st), __VA_ARGS__)
#endif

// ----------------------------------------------------------------------------- device helpers
#ifdef __CUDACC__
namespace b2u {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- explicit shared-space 16-byte accesses (a pointer re-derived through an integer cast makes nvcc emit GENERIC
// LD.E / ST.E for shared memory; these stay LDS.128 / STS.128)
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must surface as a trapped kernel (launch error), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("b2u: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}

// ---- TMA (cp.async.bulk.tensor), tile mode, completion on an mbarrier
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

// ---- TMA stores (shared -> global, bulk async-group completion).  The box is clipped at the tensor bounds, so ragged
// border tiles need no per-row predicate; the source rows must be visible to the async proxy (fence_proxy_async)
// before the store is issued, and the staging buffer may be rewritten once `bulk_wait_read<N>` has returned.
__device__ __forceinline__ void tma_store_4d(const void* tmap, uint32_t src_smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_5d(const void* tmap, uint32_t src_smem, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// ---- programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may
// be scheduled while its predecessor in the stream is still running; `pdl_wait()` (griddepcontrol.wait) blocks until the
// predecessor grid has completed and its memory is visible, so everything a kernel does BEFORE the wait (barrier
// init, TMEM allocation, descriptor prefetch, index arithmetic) and its launch latency overlap the predecessor's
// tail.  `pdl_trigger()` (griddepcontrol.launch_dependents) lets the successor be scheduled once every CTA of this grid
// has issued it or exited.  Both are no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16/fp16 inputs, kind::tf32 fp32 inputs.
template <bool kTf32>
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  if constexpr (kTf32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// Arrives on the mbarrier once every previously issued tcgen05.mma of this thread has completed
// (implies tcgen05.fence::before_thread_sync).
// Warp-converged issue: the WHOLE warp executes these and the instruction itself is predicated on `leader` (one lane,
// chosen once with elect_one()).  Issuing from inside `if (lane == 0)` makes the region divergent, and ptxas then wraps
// every tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall to move its operands into uniform registers:
// ~9 dependent instructions = ~100 cycles per MMA, more than an M128 x N64 x K16 MMA takes on the tensor pipe.
template <bool kTf32>
__device__ __forceinline__ void umma_ss_conv(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate, uint32_t leader) {
  if constexpr (kTf32) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 e, %5, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 e, %5, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
        : "memory");
  }
}
// Four K steps (4 x 32 bytes along K: one 128-byte swizzle row) in ONE asm statement.  Only the low word of a shared
// memory descriptor moves (start address, 16-byte units); the high words are per-kernel constants, so the advance is a
// 32-bit add and ptxas needs one register -> uniform-register move per descriptor instead of two per MMA.
template <bool kTf32>
__device__ __forceinline__ void umma_ss_conv4(uint32_t tmem_d, uint32_t adesc_lo, uint32_t adesc_hi, uint32_t bdesc_lo,
                                              uint32_t bdesc_hi, uint32_t idesc, uint32_t accumulate_first, uint32_t leader) {
#define B2U_MMA4(KIND)                                                                                              \
  asm volatile(                                                                                                     \
      "{\n\t.reg .pred p, e, t;\n\t.reg .b32 x, y;\n\t.reg .b64 a, b;\n\t"                                        \
      "setp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 e, %5, 0;\n\tsetp.eq.b32 t, 0, 0;\n\t"                                \
      "mov.b64 a, {%1, %6};\n\tmov.b64 b, {%2, %7};\n\t"                                                          \
      "@e tcgen05.mma.cta_group::1.kind::" KIND " [%0], a, b, %3, p;\n\t"                                           \
      "add.u32 x, %1, 2;\n\tadd.u32 y, %2, 2;\n\tmov.b64 a, {x, %6};\n\tmov.b64 b, {y, %7};\n\t"                 \
      "@e tcgen05.mma.cta_group::1.kind::" KIND " [%0], a, b, %3, t;\n\t"                                           \
      "add.u32 x, %1, 4;\n\tadd.u32 y, %2, 4;\n\tmov.b64 a, {x, %6};\n\tmov.b64 b, {y, %7};\n\t"                 \
      "@e tcgen05.mma.cta_group::1.kind::" KIND " [%0], a, b, %3, t;\n\t"                                           \
      "add.u32 x, %1, 6;\n\tadd.u32 y, %2, 6;\n\tmov.b64 a, {x, %6};\n\tmov.b64 b, {y, %7};\n\t"                 \
      "@e tcgen05.mma.cta_group::1.kind::" KIND " [%0], a, b, %3, t;\n\t}\n"                                       \
      ::"r"(tmem_d), "r"(adesc_lo), "r"(bdesc_lo), "r"(idesc), "r"(accumulate_first), "r"(leader), "r"(adesc_hi),    \
        "r"(bdesc_hi)                                                                                               \
      : "memory")
  if constexpr (kTf32) B2U_MMA4("tf32");
  else B2U_MMA4("f16");
#undef B2U_MMA4
}
template <bool kTf32>
__device__ __forceinline__ void umma_ss_conv4(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate_first, uint32_t leader) {
  umma_ss_conv4<kTf32>(tmem_d, static_cast<uint32_t>(adesc), static_cast<uint32_t>(adesc >> 32), static_cast<uint32_t>(bdesc),
                       static_cast<uint32_t>(bdesc >> 32), idesc, accumulate_first, leader);
}
// TMEM base address as a value ptxas can prove warp-uniform (it is read from shared memory, i.e. per thread): without
// this every tcgen05.mma gets its own predicated R2UR.BROADCAST of the accumulator address.
__device__ __forceinline__ uint32_t warp_uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ void umma_commit_conv(uint64_t* bar, uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred e;\n\tsetp.ne.b32 e, %1, 0;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n"
      ::"r"(smem_u32(bar)), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets row (lane base + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows of 128 B, 8-row
// swizzle atoms 1024 B apart (SBO), LBO unused (1), descriptor version 1 (sm_100), layout type 2.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Same, with an explicit stride between 8-row groups (shifted views of a halo patch use 10 rows = 1280 B) and a
// start address that only needs 128 B (one row) alignment: the swizzle is a function of the absolute address.
__device__ __forceinline__ uint64_t umma_desc_k_sw128_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Instruction descriptor: fp32 accumulate, A/B both K-major, format 1 = bf16 (kind::f16) or 2 = tf32.
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n, int ab_format) {
  return (1u << 4) | (static_cast<uint32_t>(ab_format) << 7) | (static_cast<uint32_t>(ab_format) << 10) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// ---- misc
__device__ __forceinline__ float round_tf32(float x) {   // round-to-nearest (ties away) to the 10-bit tf32 mantissa
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// Storage conversion of an MMA operand: bf16 round-to-nearest-even, or fp32 pre-rounded to tf32 (the
// tensor core would otherwise truncate the low 13 mantissa bits, a biased error twice as large).
template <typename T> __device__ __forceinline__ T to_operand(float x);
template <> __device__ __forceinline__ __nv_bfloat16 to_operand<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }
template <> __device__ __forceinline__ float to_operand<float>(float x) { return round_tf32(x); }
// fp16 operands (kind::f16 with the F16 format): 11-bit mantissa at bf16 speed; values are saturated to the finite
// range instead of overflowing to inf (post-GroupNorm activations and kaiming-scale weights are O(1))
__device__ __forceinline__ float sat_f16(float x) { return fminf(fmaxf(x, -65504.f), 65504.f); }
// two floats -> packed fp16x2, round-to-nearest-even, saturated to +-65504 in ONE instruction (F2FP.SATFINITE): the
// explicit fmin / fmax pair per element cost the fp16 mode 64 extra instructions per 32-column epilogue chunk
__device__ __forceinline__ uint32_t pack_half2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <> __device__ __forceinline__ __half to_operand<__half>(float x) {
  const uint32_t r = pack_half2_sat(x, 0.f);
  return __ushort_as_half(static_cast<unsigned short>(r & 0xFFFFu));
}

// Storage format of a tensor-core kernel: 0 = bf16, 1 = fp32 storage / tf32 math, 2 = fp16 (== the B2U_* dtype codes)
template <int F> struct FmtTraits;
template <> struct FmtTraits<0> { using T = __nv_bfloat16; static constexpr int kIdescFmt = 1; };
template <> struct FmtTraits<1> { using T = float; static constexpr int kIdescFmt = 2; };
template <> struct FmtTraits<2> { using T = __half; static constexpr int kIdescFmt = 0; };

// 32 consecutive accumulator columns of one row -> global memory in the storage type
template <typename T> __device__ __forceinline__ void store_chunk32(T* dst, const float (&x)[32]);
template <> __device__ __forceinline__ void store_chunk32<float>(float* dst, const float (&x)[32]) {
  float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int i = 0; i < 8; ++i) d4[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
}
template <> __device__ __forceinline__ void store_chunk32<__nv_bfloat16>(__nv_bfloat16* dst, const float (&x)[32]) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 a = __floats2bfloat162_rn(x[8 * i], x[8 * i + 1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(x[8 * i + 2], x[8 * i + 3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(x[8 * i + 4], x[8 * i + 5]);
    __nv_bfloat162 d = __floats2bfloat162_rn(x[8 * i + 6], x[8 * i + 7]);
    uint4 v;
    v.x = *reinterpret_cast<uint32_t*>(&a);
    v.y = *reinterpret_cast<uint32_t*>(&b);
    v.z = *reinterpret_cast<uint32_t*>(&c);
    v.w = *reinterpret_cast<uint32_t*>(&d);
    d4[i] = v;
  }
}
template <> __device__ __forceinline__ void store_chunk32<__half>(__half* dst, const float (&x)[32]) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 v;
    v.x = pack_half2_sat(x[8 * i], x[8 * i + 1]);
    v.y = pack_half2_sat(x[8 * i + 2], x[8 * i + 3]);
    v.z = pack_half2_sat(x[8 * i + 4], x[8 * i + 5]);
    v.w = pack_half2_sat(x[8 * i + 6], x[8 * i + 7]);
    d4[i] = v;
  }
}

// 8 consecutive accumulator columns -> one 16-byte vector of the 16-bit storage type (same rounding as store_chunk32)
template <typename T> __device__ __forceinline__ uint4 pack8(const float* x);
template <> __device__ __forceinline__ uint4 pack8<__nv_bfloat16>(const float* x) {
  __nv_bfloat162 a = __floats2bfloat162_rn(x[0], x[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(x[2], x[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(x[4], x[5]);
  __nv_bfloat162 d = __floats2bfloat162_rn(x[6], x[7]);
  uint4 v;
  v.x = *reinterpret_cast<uint32_t*>(&a);
  v.y = *reinterpret_cast<uint32_t*>(&b);
  v.z = *reinterpret_cast<uint32_t*>(&c);
  v.w = *reinterpret_cast<uint32_t*>(&d);
  return v;
}
template <> __device__ __forceinline__ uint4 pack8<__half>(const float* x) {
  uint4 v;
  v.x = pack_half2_sat(x[0], x[1]);
  v.y = pack_half2_sat(x[2], x[3]);
  v.z = pack_half2_sat(x[4], x[5]);
  v.w = pack_half2_sat(x[6], x[7]);
  return v;
}
template <> __device__ __forceinline__ uint4 pack8<float>(const float* x) { return make_uint4(0u, 0u, 0u, 0u); }   // never staged

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum NV per-lane values across the 32 lanes of a warp with NV-1 (+ up to 4) shuffles.  On return
// v[j] (j < max(NV/32,1)) holds the total of value index (lane*NV)/32 + j (replicated when NV < 32).
template <int NV>
__device__ __forceinline__ void warp_transpose_reduce(float (&v)[NV], int lane) {
  int cur = NV;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    if (cur > 1) {
      const int half = cur >> 1;
      const bool up = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < NV / 2; ++i) {
        if (i < half) {
          float send = up ? v[i] : v[i + half];
          float keep = up ? v[i + half] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      cur = half;
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
    }
  }
}

template <typename T> struct Vec8;   // 8 consecutive channels
template <> struct Vec8<__nv_bfloat16> {
  uint4 raw;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { raw = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ __forceinline__ void to_float(float (&f)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  __device__ __forceinline__ void from_float(const float (&f)[8]) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  }
};
template <> struct Vec8<__half> {
  uint4 raw;
  __device__ __forceinline__ void load(const __half* p) { raw = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__half* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ __forceinline__ void to_float(float (&f)[8]) const {
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __half22float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  __device__ __forceinline__ void from_float(const float (&f)[8]) {
    raw.x = pack_half2_sat(f[0], f[1]);
    raw.y = pack_half2_sat(f[2], f[3]);
    raw.z = pack_half2_sat(f[4], f[5]);
    raw.w = pack_half2_sat(f[6], f[7]);
  }
};
template <> struct Vec8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = a;
    *reinterpret_cast<float4*>(p + 4) = b;
  }
  __device__ __forceinline__ void to_float(float (&f)[8]) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  __device__ __forceinline__ void from_float(const float (&f)[8]) {
    a = make_float4(f[0], f[1], f[2], f[3]);
    b = make_float4(f[4], f[5], f[6], f[7]);
  }
};

}  // namespace b2u
#endif  // __CUDACC__
