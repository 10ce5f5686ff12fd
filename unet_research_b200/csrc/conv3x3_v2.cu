// conv3x3 v2: persistent, halo-patch-reuse implicit GEMM (sm_100a, tcgen05 + TMEM + TMA).
//
// v1 (conv_gemm.cu) re-fetches a 16 KB activation tile for each of the 9 filter taps: the shallow layers are
// L2->SM bandwidth bound (profiles/r01_conv_ncu.md).  v2 changes three things:
//
//  1. HALO PATCH REUSE.  A pixel tile is 8 wide x 16 tall.  Per 64-channel chunk ONE TMA box {128 B, 10, 18}
//     brings the (16+2) x (8+2) input patch into shared memory (180 rows x 128 B, 128B swizzle; TMA zero fill =
//     conv padding).  The 9 taps are 9 UMMA descriptors over that patch: start address + (r*10 + s) rows,
//     8-row-group stride (SBO) = 10 rows = 1280 B.  A K-major SWIZZLE_128B descriptor may start at any 128 B
//     row because the swizzle is a function of the absolute shared-memory address (hardware experiment
//     tests/exp_umma_shift.cu, profiles/r01_exp_umma_shift.log).  Activation traffic drops from 9x to 1.41x.
//  2. WEIGHT-STAGE SHARING.  A work item is MT (1 or 2) pixel tiles x BLOCK_N channels: each (tap, chunk)
//     weight stage feeds MT accumulators, halving weight traffic per FLOP for MT = 2.
//  3. PERSISTENT CTAs, double-buffered TMEM.  grid = min(items, SMs * resident CTAs); a CTA loops over items
//     (n-tile-major order, so concurrently running CTAs share weight tiles in L2); when 2*MT*BLOCK_N <= 512
//     columns the epilogue of item j overlaps the main loop of item j+1.
//
// Warp roles (224 threads): warp 0 = activation-patch TMA producer, warp 6 = weight TMA producer,
// warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue (TMEM -> GroupNorm partials -> bf16/fp32 -> global).
#include "b2u_common.cuh"
#include "conv_host.cuh"

#include <stdlib.h>

namespace b2u {

// Division by a launch-invariant divisor as one IMAD.HI: q = (n * mul) >> 32 with mul = floor(2^32 / d) + 1, exact for
// n * d < 2^32 (the host checks it).  The transform warps locate a patch with three of them instead of three ~30
// instruction software divisions.
__device__ __forceinline__ int fd_div(int n, uint32_t mul) {      // mul == 0 encodes d == 1
  return mul ? static_cast<int>(__umulhi(static_cast<uint32_t>(n), mul)) : n;
}
static inline uint32_t fd_make(int d) { return d <= 1 ? 0u : static_cast<uint32_t>((1ull << 32) / static_cast<unsigned>(d)) + 1u; }

struct ConvV2Params {
  int n, h, w;                // image grid
  int tiles_w, tiles_h;       // 8 x 16 tiles per image
  int total_tiles;            // n * tiles_w * tiles_h
  int num_mgroups;            // ceil(total_tiles / MT)
  int num_items;              // num_mgroups * n_tiles
  int kc_chunks;              // Cin / (128 B of channels)
  int cout;
  int sa, sb;                 // pipeline depths of the patch ring and the weight ring
  int sgs_log2;               // statistics sub-group size (log2), -1 = none
  int resident_b;             // 1: the whole filter (9 * kc_chunks stages, one N tile) stays in shared memory for the CTA's lifetime
  void* y;
  float* partials;            // [n][tiles_per_image][cout/sgs][2]
  // fused A-operand prologue (kPro kernels): x is the RAW output of the producing conv / pool; the transform warps rewrite
  // every TMA-landed patch in shared memory as act = [relu]((a * x + b) * keep) before the MMA warp may read it
  const float2* pro_coef;     // [n][cin] (a, b) of the producer's GroupNorm (DropBlock rescale folded in)
  const uint8_t* pro_mask;    // NHWC keep bits of the producer's DropBlock site, [n][h][w][cin / 8] bytes, or nullptr
  int pro_relu;
  int x_shared;               // 1: every image reads the activation tensor of image 0 (Monte-Carlo: shared first conv)
  int cin;
  uint32_t fd_mgroups, fd_tpi, fd_tw;   // fd_make(num_mgroups), fd_make(tiles per image), fd_make(tiles_w)
  int combine_stats;          // MT = 2 and an even number of tiles per image: one statistics reduction per item
  int stage_off;              // byte offset of the four 2 KB staging slots inside the aligned shared memory
};

// TMA-store epilogue (16-bit formats): tiles up to this BLOCK_N stage their output rows in shared memory and write them
// with TMA stores (256 = every 16-bit kernel, 0 = never).  Compile-time so that the unused path costs no registers: A/B
// runs build two libraries (python tests/build_variant.py stg -DB2U_V2_TMA_MAX_BN=0) and select one with B2U_LIB.
// Measured stand-alone at batch 10, fp16 (profiles/r02d_exp_tma_store.log): 64->64 281 -> 241 us, 128->128 200 -> 175,
// 128->256 111 -> 96, 256->256 184 -> 170, 512->512 176 -> 169, 1024->1024 243 -> 220.
#ifndef B2U_V2_TMA_MAX_BN
#define B2U_V2_TMA_MAX_BN 256
#endif
constexpr int kPatchRows = 180;                      // (16 + 2) * (8 + 2)
constexpr int kPatchBytes = kPatchRows * 128;        // 23040
constexpr int kPatchStride = 23552;                  // rounded up to the 1024 B swizzle-atom alignment
constexpr int kV2Threads = 224;
constexpr int kProWarps = 8;                         // transform warps of the fused-prologue kernels (warps 7 .. 14), two groups
constexpr int kV2ThreadsPro = kV2Threads + 32 * kProWarps;

// ---- fused prologue: one 16-byte vector (8 channels of one patch pixel) through GroupNorm affine, DropBlock, ReLU.
// Same arithmetic, in the same order and precision, as gn_apply_kernel (fp32 fma -> mask -> relu -> storage rounding), so
// the MMA consumes bit-identical operands whether the apply ran as its own pass or here.
template <int kFmt>
__device__ __forceinline__ uint4 pro_transform(uint4 raw, const float (&a)[8], const float (&b)[8], uint32_t mbits, float lo_clamp) {
  using T = typename FmtTraits<kFmt>::T;
  Vec8<T> v;
  v.raw = raw;
  float f[8];
  v.to_float(f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float t = fmaf(f[i], a[i], b[i]);
    t = ((mbits >> i) & 1u) ? t : 0.f;
    f[i] = fmaxf(t, lo_clamp);
  }
  v.from_float(f);
  return v.raw;
}

template <int NV>
__device__ __forceinline__ void v2_epilogue_stats(const float (&x)[32], bool valid, int lane, float* scratch) {
  constexpr int NSG = NV / 2;
  constexpr int SGS = 32 / NSG;
  float v[NV];
#pragma unroll
  for (int j = 0; j < NSG; ++j) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int i = 0; i < SGS; ++i) {
      float t = x[j * SGS + i];
      s += t;
      q += t * t;
    }
    v[2 * j] = valid ? s : 0.f;
    v[2 * j + 1] = valid ? q : 0.f;
  }
  warp_transpose_reduce<NV>(v, lane);
  constexpr int LPV = 32 / NV;
  if (lane % LPV == 0) scratch[lane / LPV] = v[0];
}

// MT = 2 items whose two tiles lie in the same image: the per-thread sub-group sums of BOTH tiles are added before the
// 31-shuffle transposed reduction, which then runs once per item and chunk instead of once per tile and chunk (it is the
// critical path of the Cout = 64 / 128 layers: ~6200 cycles of epilogue per item next to 2304 of MMA).  The combined
// sum goes to the statistics row of the first tile, the second tile's row is written as zero: gn_finalize adds all
// rows of an image, so the row granularity is free.
template <int NV>
__device__ __forceinline__ void v2_partial_sums(const float (&x)[32], bool valid, float (&v)[NV], bool accumulate) {
  constexpr int NSG = NV / 2;
  constexpr int SGS = 32 / NSG;
#pragma unroll
  for (int j = 0; j < NSG; ++j) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int i = 0; i < SGS; ++i) {
      const float t = x[j * SGS + i];
      s += t;
      q += t * t;
    }
    s = valid ? s : 0.f;
    q = valid ? q : 0.f;
    v[2 * j] = accumulate ? v[2 * j] + s : s;
    v[2 * j + 1] = accumulate ? v[2 * j + 1] + q : q;
  }
}
// One 32-column accumulator chunk of this warp's 32 pixels (4 rows x 8 pixels of the tile) -> global memory.
// Per-thread path: a thread writes its own 64-byte run; a warp-wide STG.128 then touches 32 different 128-byte lines.
// TMA path (16-bit formats): the warp stages its 32 x 64-byte runs in its 2 KB slot (64-byte swizzle: physical 16-byte
// chunk = i ^ ((row >> 1) & 3), conflict-free for STS.128) and lane 0 stores the box {32 ch, 8 w, 4 h, 1 n}; the tensor
// bounds clip ragged tiles and tiles past the end.  Single-buffered: the wait for the previous store's shared-memory
// read sits behind the TMEM load and the statistics arithmetic of this chunk.
template <typename OutT>
__device__ __forceinline__ void v2_store_chunk(const CUtensorMap* tmY, uint32_t slot, int lane, const float (&x)[32], int c0, int w0, int h0,
                                               int img) {
  if (lane == 0) bulk_wait_read<0>();
  __syncwarp();
  const uint32_t row = slot + static_cast<uint32_t>(lane * 64);
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) sts128(row + static_cast<uint32_t>((i ^ sw) << 4), pack8<OutT>(x + 8 * i));
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) {
    tma_store_4d(tmY, slot, c0, w0, h0, img);
    bulk_commit();
  }
}

template <int NV, typename St0, typename St1>
__device__ __forceinline__ void v2_chunk_pair(uint32_t taddr0, uint32_t taddr1, bool valid0, bool valid1, St0&& store0, St1&& store1, int lane,
                                              float* scratch) {
  float v[NV];
  {
    uint32_t rr[32];
    tmem_ld_32x32(taddr0, rr);
    tmem_ld_wait();
    float x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(rr[i]);
    v2_partial_sums<NV>(x, valid0, v, false);
    store0(x);
  }
  {
    uint32_t rr[32];
    tmem_ld_32x32(taddr1, rr);
    tmem_ld_wait();
    float x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(rr[i]);
    v2_partial_sums<NV>(x, valid1, v, true);
    store1(x);
  }
  warp_transpose_reduce<NV>(v, lane);
  constexpr int LPV = 32 / NV;
  if (lane % LPV == 0) scratch[lane / LPV] = v[0];
}

template <int BLOCK_N, int MT, int kFmt, bool kPro>
// kPro kernels launch 480 threads but are compiled for a 680-thread bound = 96 registers per thread: 46 k registers per
// CTA.  A CTA that owns 59 k of the SM's 64 k registers (352 x 168) evicts the co-resident DropBlock mask-build blocks
// (8 k registers each) that the Monte-Carlo step overlaps with the forward -- measured: +0.5 ms per step.
__global__ void __launch_bounds__(kPro ? 680 : 416, 1)   // register caps: 80 (pro) / 152 (plain), see above
conv3x3_v2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmY,
                  const ConvV2Params p) {
  constexpr bool kTf32 = kFmt == 1;
  constexpr bool kTmaSt = !kTf32 && BLOCK_N <= B2U_V2_TMA_MAX_BN;
  static_assert(!(kPro && kTf32), "the fused prologue is built for the 16-bit storage formats");
  using OutT = typename FmtTraits<kFmt>::T;
  constexpr int kBBytes = BLOCK_N * 128;
  constexpr int kKElems = kTf32 ? 32 : 64;
  constexpr int kUmmaK = kTf32 ? 8 : 16;
  constexpr uint32_t kIdesc = umma_idesc(128, BLOCK_N, FmtTraits<kFmt>::kIdescFmt);
  constexpr int kAccCols = MT * BLOCK_N;                       // TMEM columns of one accumulator buffer
  constexpr int kNumBuf = (2 * kAccCols <= 512) ? 2 : 1;
  constexpr int kTmemCols = (kNumBuf * kAccCols <= 64) ? 64 : (kNumBuf * kAccCols <= 128) ? 128 : (kNumBuf * kAccCols <= 256) ? 256 : 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int SA = p.sa, SB = p.sb;
  uint8_t* smA = smem;                                         // [SA][MT][kPatchStride]
  uint8_t* smB = smem + SA * MT * kPatchStride;                // [SB][kBBytes]
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smB + SB * kBBytes);
  uint64_t* a_empty = a_full + SA;
  uint64_t* b_full = a_empty + SA;
  uint64_t* b_empty = b_full + SB;
  uint64_t* t_full = b_empty + SB;                             // [2]
  uint64_t* t_empty = t_full + 2;                              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  float* stat_scratch = reinterpret_cast<float*>(tmem_slot + 2);   // [4 warps][128]
  uint64_t* a_ready = reinterpret_cast<uint64_t*>(stat_scratch + 4 * 128);   // [SA] (kPro): patch transformed, MMA may read

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_per_image = p.tiles_w * p.tiles_h;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (kTmaSt) tma_prefetch_desc(&tmY);
    for (int s = 0; s < SA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 4); }
    if (kPro)
      for (int s = 0; s < SA; ++s) mbar_init(&a_ready[s], 4 * MT);     // one arrive per warp that transformed a patch of the stage
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: barrier init, TMEM allocation and descriptor prefetch above overlapped the predecessor's tail; from here on the
  // kernel touches tensors the predecessor wrote (or still reads)
  pdl_wait();

  // All three single-thread role loops below keep (stage, phase) pairs that advance by compare-and-wrap:
  // no runtime division / modulo on the issue path (a single thread issues every TMA / MMA of the CTA, so
  // its instruction count per filter tap bounds the tensor pipe at small BLOCK_N).
  if (warp == 0) {
    // ===================== activation-patch producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;                                          // first pass over a fresh barrier: wait(1) returns
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        if (item + static_cast<int>(gridDim.x) >= p.num_items) pdl_trigger();   // last item of this CTA: let the successor launch
        const int mg = item % p.num_mgroups;
        int cw[MT], ch[MT], cn[MT];
#pragma unroll
        for (int t = 0; t < MT; ++t) {
          const int tile = mg * MT + t;                        // tiles past the end are fully out of bounds -> zeros
          const int img = tile / tiles_per_image;
          const int r = tile - img * tiles_per_image;
          const int ty = r / p.tiles_w;
          cn[t] = p.x_shared ? 0 : img;
          ch[t] = ty * 16 - 1;
          cw[t] = (r - ty * p.tiles_w) * 8 - 1;
        }
        for (int kc = 0; kc < p.kc_chunks; ++kc) {
          mbar_wait(&a_empty[s], ph);
          mbar_arrive_expect_tx(&a_full[s], MT * kPatchBytes);
#pragma unroll
          for (int t = 0; t < MT; ++t)
            tma_load_4d(smA + (s * MT + t) * kPatchStride, &tmA, &a_full[s], kc * kKElems, cw[t], ch[t], cn[t]);
          if (++s == SA) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 6) {
    // ===================== weight producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        // resident filter (Cout = BLOCK_N, 9 * kc_chunks = SB stages): loaded once, by the CTA's first item -- every
        // later item would fetch the same 9 * kc_chunks * BLOCK_N * 128 bytes again (0.98 GB per 64->64 launch)
        if (p.resident_b && item != static_cast<int>(blockIdx.x)) break;
        const int n0 = (item / p.num_mgroups) * BLOCK_N;
        for (int kc = 0; kc < p.kc_chunks; ++kc) {
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&b_empty[s], ph);
            mbar_arrive_expect_tx(&b_full[s], kBBytes);
            tma_load_3d(smB + s * kBBytes, &tmB, &b_full[s], kc * kKElems, n0, tap);
            if (++s == SB) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // the whole warp runs the loop (converged control flow); one elected lane issues (see umma_ss_conv)
    {
      const uint32_t leader = elect_one() ? 1u : 0u;
      int sa = 0, sb = 0, buf = 0;
      uint32_t pa = 0, pb = 0, pt = 1;                           // pt: parity to wait on t_empty[buf]
      const uint64_t adesc0 = umma_desc_k_sw128_sbo(smem_u32(smA), 1280);
      const uint64_t bdesc0 = umma_desc_k_sw128(smem_u32(smB));
      // only the low descriptor word (start address in 16-byte units) changes between MMAs
      const uint32_t a_lo0 = static_cast<uint32_t>(adesc0), a_hi = static_cast<uint32_t>(adesc0 >> 32);
      const uint32_t b_lo0 = static_cast<uint32_t>(bdesc0), b_hi = static_cast<uint32_t>(bdesc0 >> 32);
      const uint32_t tmem_u = warp_uniform(tmem_base);
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        mbar_wait(&t_empty[buf], pt);                            // epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t acc0 = tmem_u + buf * kAccCols;
        for (int kc = 0; kc < p.kc_chunks; ++kc) {
          mbar_wait(kPro ? &a_ready[sa] : &a_full[sa], pa);
          const uint32_t a_stage = a_lo0 + static_cast<uint32_t>((sa * MT * kPatchStride) >> 4);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            if (!p.resident_b || item == static_cast<int>(blockIdx.x)) mbar_wait(&b_full[sb], pb);
            tc_fence_after();
            const uint32_t b_lo = b_lo0 + static_cast<uint32_t>((sb * kBBytes) >> 4);
            constexpr int kRowsPerGroup = 10;
            const int row_off = (tap / 3) * kRowsPerGroup + (tap % 3);     // compile-time: shifted view of the patch
#pragma unroll
            for (int t = 0; t < MT; ++t) {
              const uint32_t a_lo = a_stage + static_cast<uint32_t>((t * kPatchStride + row_off * 128) >> 4);
              static_assert(kKElems / kUmmaK == 4, "one swizzle row = four K steps");
              const uint32_t accumulate = tap == 0 ? (kc != 0 ? 1u : 0u) : 1u;
              umma_ss_conv4<kTf32>(acc0 + t * BLOCK_N, a_lo, a_hi, b_lo, b_hi, kIdesc, accumulate, leader);
            }
            if (!p.resident_b) umma_commit_conv(&b_empty[sb], leader);   // resident stages are never recycled
            if (++sb == SB) { sb = 0; pb ^= 1; }
          }
          umma_commit_conv(&a_empty[sa], leader);
          if (++sa == SA) { sa = 0; pa ^= 1; }
        }
        umma_commit_conv(&t_full[buf], leader);
        if (++buf == kNumBuf) { buf = 0; pt ^= 1; }
      }
    }
    __syncwarp();
  } else if (kPro && warp >= 7) {
    // ===================== prologue transform (warps 7 .. 10) =====================
    // Thread tt owns the 16-byte channel chunk j = tt & 7 of patch rows r0 + 16 k (k = 0 .. 11): its 8 (a, b) pairs live
    // in registers for a whole patch, a quarter warp touches one full 128-byte row per access (conflict-free under the
    // 128-byte swizzle: physical chunk = j ^ (r & 7)).  Pixels outside the image keep TMA's zero fill: the conv's zero
    // padding applies to the ACTIVATED tensor (reference utils_unet.py:166-182: Conv -> GroupNorm -> DropBlock -> ReLU,
    // the next Conv2d pads its input), so the affine must not touch them.
    if constexpr (kPro) {
      // Two groups of four warps take turns: with MT = 2 group g owns patch t = g of EVERY stage, with MT = 1 the
      // stages of parity g.  A single warp per scheduler ran this role latency-bound (IPC 0.25, ~3000 cycles per patch
      // against 1152-2304 of MMA per patch on the two shallow levels); the second warp per scheduler hides it.
      const int grp = (threadIdx.x - kV2Threads) >> 7;           // 0 / 1
      const int tt = (threadIdx.x - kV2Threads) & 127;
      const int j = tt & 7;
      const int r0 = tt >> 3;                                    // 0 .. 15
      const int cvs = p.cin >> 3;                                // 16-byte vectors (= mask bytes) per pixel
      const float lo_clamp = p.pro_relu ? 0.f : -3.0e38f;        // storage conversion saturates fp16 on its own
      const uint32_t rows_mask = r0 < 4 ? 0xFFFu : 0x7FFu;       // row r0 + 176 exists only for r0 < 4 (180 rows)
      // Mask bytes and in-image flags of a patch come straight from global memory and are fetched ONE PATCH (of this
      // group) AHEAD -- they stream from HBM; the 8 coefficient pairs (L2 hits) are fetched when the patch starts.
      struct Meta {
        uint32_t mb[12];                                         // consumed one patch later, never touched before
        uint32_t inb;                                            // bit k: patch row r0 + 16 k is a real pixel
        int img;
      };
      auto fetch = [&](int item, int kc, int t, Meta& m) {
        const int mg = item - fd_div(item, p.fd_mgroups) * p.num_mgroups;
        const int tile = mg * MT + t;
        const bool tile_ok = item < p.num_items && tile < p.total_tiles;
        const int img = tile_ok ? fd_div(tile, p.fd_tpi) : 0;
        const int r = tile_ok ? tile - img * tiles_per_image : 0;
        const int ty = fd_div(r, p.fd_tw);
        const int h0 = ty * 16 - 1, w0 = (r - ty * p.tiles_w) * 8 - 1;
        m.img = img;
        // interior patches (92 % at 592x576): every row is a real pixel, no per-row tests
        uint32_t inb = rows_mask;
        if (!(h0 >= 0 && h0 + 17 < p.h && w0 >= 0 && w0 + 9 < p.w)) {       // warp-uniform
#pragma unroll
          for (int k = 0; k < 12; ++k) {
            const int row = r0 + 16 * k;
            const int py = (row * 205) >> 11, px = row - py * 10;          // row / 10, row % 10 for row < 192
            const int hh = h0 + py, ww = w0 + px;
            if (!(hh >= 0 && hh < p.h && ww >= 0 && ww < p.w)) inb &= ~(1u << k);
          }
        }
        if (!tile_ok) inb = 0;
        m.inb = inb;
        if (p.pro_mask) {
          const uint8_t* mbase = p.pro_mask + (static_cast<size_t>(img) * p.h * p.w + static_cast<long>(h0) * p.w + w0) * cvs + kc * 8 + j;
#pragma unroll
          for (int k = 0; k < 12; ++k) {
            const int row = r0 + 16 * k;
            const int py = (row * 205) >> 11, px = row - py * 10;
            m.mb[k] = 0xFFu;
            if ((inb >> k) & 1u) m.mb[k] = __ldg(mbase + (py * p.w + px) * cvs);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 12; ++k) m.mb[k] = 0xFFu;
        }
      };
      const uint32_t smA_u32 = smem_u32(smA);
      // this group's patch sequence: (stage counter sc, item, kc, t)
      const int t_mine = MT == 2 ? grp : 0;
      const int stage_step = MT == 2 ? 1 : 2;
      int c_item = blockIdx.x, c_kc = 0, c_sc = 0;               // cursor of the patch `nxt` describes
      auto advance_stages = [&](int n) {
        for (int i = 0; i < n; ++i) {
          ++c_sc;
          if (++c_kc == p.kc_chunks) { c_kc = 0; c_item += gridDim.x; }
        }
      };
      if (MT == 1 && grp == 1) advance_stages(1);
      Meta cur, nxt;
      fetch(c_item, c_kc, t_mine, nxt);
      while (c_item < p.num_items) {
        cur = nxt;
        const int kc = c_kc, sc = c_sc;
        advance_stages(stage_step);
        fetch(c_item, c_kc, t_mine, nxt);                         // next patch's mask bytes in flight during this transform
        float ca[8], cb[8];
        {
          const float4* cf = reinterpret_cast<const float4*>(p.pro_coef + static_cast<size_t>(cur.img) * p.cin + kc * 64 + j * 8);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 c4 = __ldg(cf + i);
            ca[2 * i] = c4.x; cb[2 * i] = c4.y; ca[2 * i + 1] = c4.z; cb[2 * i + 1] = c4.w;
          }
        }
        const int s = sc % p.sa;
        const uint32_t ph = static_cast<uint32_t>(sc / p.sa) & 1u;
        mbar_wait(&a_full[s], ph);                               // TMA bytes of all MT patches of this stage have landed
        // row r0 + 16 k: (row & 7) = (r0 & 7) for every k, so the swizzled chunk offset is a per-thread constant
        const uint32_t patch = smA_u32 + static_cast<uint32_t>((s * MT + t_mine) * kPatchStride + r0 * 128 + ((j ^ (r0 & 7)) << 4));
        // two halves of six rows: loads first (unconditional: rows outside the image hold TMA's zeros; only row
        // 176 + r0 may not exist), then the arithmetic, then predicated stores
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint4 v[6];
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const int k = half * 6 + i;
            v[i] = (k < 11 || r0 < 4) ? lds128(patch + k * 2048) : make_uint4(0u, 0u, 0u, 0u);
          }
#pragma unroll
          for (int i = 0; i < 6; ++i) v[i] = pro_transform<kFmt>(v[i], ca, cb, cur.mb[half * 6 + i], lo_clamp);
#pragma unroll
          for (int i = 0; i < 6; ++i)
            if ((cur.inb >> (half * 6 + i)) & 1u) sts128(patch + (half * 6 + i) * 2048, v[i]);
        }
        // generic-proxy writes -> visible to the tensor core's async-proxy reads, then one arrive per warp
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[s]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int dh = row >> 3, dw = row & 7;
    float* my_scratch = stat_scratch + q * 128;
    const int et = threadIdx.x - 64;                             // 0..127
    const uint32_t my_slot = smem_u32(smem + p.stage_off) + static_cast<uint32_t>(q * 2048);
    int buf = 0;
    uint32_t pf = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int n0 = (item / p.num_mgroups) * BLOCK_N;
      const int mg = item % p.num_mgroups;
      mbar_wait(&t_full[buf], pf);
      tc_fence_after();
      if (MT == 2 && p.combine_stats && p.sgs_log2 >= 0) {
        // both tiles exist and belong to one image (tiles per image is even: checked by the host)
        const int tile0 = mg * 2;
        const int img = tile0 / tiles_per_image;
        const int ra = tile0 - img * tiles_per_image, rb = ra + 1;
        const int tya = ra / p.tiles_w, txa = ra - tya * p.tiles_w;
        const int tyb = rb / p.tiles_w, txb = rb - tyb * p.tiles_w;
        const int ha = tya * 16 + dh, wa = txa * 8 + dw, hb = tyb * 16 + dh, wb = txb * 8 + dw;
        const bool va = ha < p.h && wa < p.w, vb = hb < p.h && wb < p.w;
        OutT* ya = reinterpret_cast<OutT*>(p.y) + ((static_cast<size_t>(img) * p.h + ha) * p.w + wa) * p.cout + n0;
        OutT* yb = reinterpret_cast<OutT*>(p.y) + ((static_cast<size_t>(img) * p.h + hb) * p.w + wb) * p.cout + n0;
        const uint32_t tb0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kAccCols;
#pragma unroll 1
        for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
          const uint32_t t0 = tb0 + chunk * 32, t1 = tb0 + BLOCK_N + chunk * 32;
          auto st_a = [&](const float (&x)[32]) {
            if constexpr (kTmaSt) v2_store_chunk<OutT>(&tmY, my_slot, lane, x, n0 + chunk * 32, txa * 8, tya * 16 + q * 4, img);
            else if (va) store_chunk32<OutT>(ya + chunk * 32, x);
          };
          auto st_b = [&](const float (&x)[32]) {
            if constexpr (kTmaSt) v2_store_chunk<OutT>(&tmY, my_slot, lane, x, n0 + chunk * 32, txb * 8, tyb * 16 + q * 4, img);
            else if (vb) store_chunk32<OutT>(yb + chunk * 32, x);
          };
          switch (p.sgs_log2) {
            case 1: v2_chunk_pair<32>(t0, t1, va, vb, st_a, st_b, lane, my_scratch + chunk * 32); break;
            case 2: v2_chunk_pair<16>(t0, t1, va, vb, st_a, st_b, lane, my_scratch + chunk * 16); break;
            case 3: v2_chunk_pair<8>(t0, t1, va, vb, st_a, st_b, lane, my_scratch + chunk * 8); break;
            case 4: v2_chunk_pair<4>(t0, t1, va, vb, st_a, st_b, lane, my_scratch + chunk * 4); break;
            default: v2_chunk_pair<2>(t0, t1, va, vb, st_a, st_b, lane, my_scratch + chunk * 2); break;
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");           // all four quarter-tile partials are in smem
        const int nslots = (BLOCK_N * 2) >> p.sgs_log2;
        if (et < nslots) {
          const float sum = stat_scratch[et] + stat_scratch[128 + et] + stat_scratch[256 + et] + stat_scratch[384 + et];
          const int slots_per_row = (p.cout * 2) >> p.sgs_log2;
          float* prow = p.partials + (static_cast<size_t>(img) * tiles_per_image + ra) * slots_per_row + ((n0 * 2) >> p.sgs_log2) + et;
          prow[0] = sum;
          prow[slots_per_row] = 0.f;                              // the second tile's row
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");           // scratch may be overwritten by the next item
      } else
#pragma unroll 1
      for (int t = 0; t < MT; ++t) {
        const int tile = mg * MT + t;
        const bool tile_ok = tile < p.total_tiles;
        const int img = tile / tiles_per_image;
        const int r = tile - img * tiles_per_image;
        const int ty = r / p.tiles_w, tx = r - ty * p.tiles_w;
        const int hh = ty * 16 + dh, ww = tx * 8 + dw;
        const bool valid = tile_ok && hh < p.h && ww < p.w;
        OutT* yrow = reinterpret_cast<OutT*>(p.y) + ((static_cast<size_t>(img) * p.h + hh) * p.w + ww) * p.cout + n0;
#pragma unroll 1
        for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
          uint32_t rr[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kAccCols + t * BLOCK_N + chunk * 32, rr);
          tmem_ld_wait();
          float x[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(rr[i]);
          if (p.sgs_log2 >= 0) {
            switch (p.sgs_log2) {
              case 1: v2_epilogue_stats<32>(x, valid, lane, my_scratch + chunk * 32); break;
              case 2: v2_epilogue_stats<16>(x, valid, lane, my_scratch + chunk * 16); break;
              case 3: v2_epilogue_stats<8>(x, valid, lane, my_scratch + chunk * 8); break;
              case 4: v2_epilogue_stats<4>(x, valid, lane, my_scratch + chunk * 4); break;
              default: v2_epilogue_stats<2>(x, valid, lane, my_scratch + chunk * 2); break;
            }
          }
          if constexpr (kTmaSt) {
            v2_store_chunk<OutT>(&tmY, my_slot, lane, x, n0 + chunk * 32, tx * 8, ty * 16 + q * 4, img);   // a tile past the end has img == n: clipped
          } else if (valid) {
            store_chunk32<OutT>(yrow + chunk * 32, x);
          }
        }
        if (p.sgs_log2 >= 0) {
          asm volatile("bar.sync 1, 128;" ::: "memory");           // all four quarter-tile partials are in smem
          const int nslots = (BLOCK_N * 2) >> p.sgs_log2;
          if (et < nslots && tile_ok) {
            const float s = stat_scratch[et] + stat_scratch[128 + et] + stat_scratch[256 + et] + stat_scratch[384 + et];
            const int slots_per_row = (p.cout * 2) >> p.sgs_log2;
            p.partials[(static_cast<size_t>(img) * tiles_per_image + r) * slots_per_row + ((n0 * 2) >> p.sgs_log2) + et] = s;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");           // scratch may be overwritten by the next tile
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[buf]);
      if (++buf == kNumBuf) { buf = 0; pf ^= 1; }
    }
  }
  if (kTmaSt && warp >= 2 && warp <= 5 && lane == 0) bulk_wait<0>();   // staging slots read, stores complete, before the CTA exits
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ----------------------------------------------------------------------------- host side
struct V2Plan {
  int block_n, mt, sa, sb, tiles_w, tiles_h, sgs, resident_b, tma_store, stage_off;
  size_t smem;
};

constexpr int kV2StoreSlots = 4 * 2048;             // one 2 KB staging slot per epilogue warp

static int v2_make_plan(const b2u_conv_desc* d, V2Plan* pl, bool pro = false) {
  pl->tiles_w = (d->w + 7) / 8;
  pl->tiles_h = (d->h + 15) / 16;
  int bn = d->cout % 256 == 0 ? 256 : (d->cout % 128 == 0 ? 128 : 64);
  if (d->reserved[0] == 64 || d->reserved[0] == 128 || d->reserved[0] == 256) {
    B2U_REQUIRE(d->cout % d->reserved[0] == 0, "BLOCK_N override %d does not divide cout", d->reserved[0]);
    bn = d->reserved[0];
  }
  // measured on B200 (tests/gpu_diag.py convbench): BLOCK_N 256 wants the double-buffered accumulator (MT 1),
  // narrower tiles want the shared weight stage (MT 2)
  int mt = bn == 256 ? 1 : 2;
  // one 64-channel chunk of K and a 128-wide tile (the 64 -> 128 layer): the second pixel tile only lengthens the item
  // (217 -> 162 us fused, 76 -> 66 us plain at 296x288; tests/exp_conv_plan.py)
  if (bn == 128 && d->cin * (d->dtype == B2U_F32 ? 2 : 1) <= 64) mt = 1;
  if (d->reserved[3] == 1 || d->reserved[3] == 2) mt = d->reserved[3];
  // UNDER-FILLED GRIDS (batch-1 training / single-image inference at the deep levels: 15 pixel tiles at 37x36 make 30-60
  // items for 148 SMs, and every CTA walks its whole K loop alone: 50 us for a layer whose tensor time is 18 us).  When
  // the default tile leaves SMs idle, pick the narrower BLOCK_N with the lowest waves x columns-per-item x inefficiency
  // (narrow tiles re-read the activation patch per N tile: 1.1 / 1.5 for 128 / 64 against 256 at MT 1, 1.35 for 64 against
  // 128 at MT 2).  MT is NOT changed: the statistics epilogue pairs the two tiles of an MT = 2 item before its fp32
  // reduction, so a batch-dependent MT would make an image's result depend on the batch it is computed in (outputs and
  // per-chunk statistics do not depend on BLOCK_N: same K order per output element, 32-column reduction chunks).
  if (d->reserved[0] == 0 && d->reserved[3] == 0) {
    const int sms = b2u_num_sms();
    const long tiles = static_cast<long>(d->n) * pl->tiles_w * pl->tiles_h;
    auto items = [&](int b) { return ((tiles + mt - 1) / mt) * (d->cout / b); };
    if (items(bn) < sms) {
      auto ineff = [&](int b) { return mt == 1 ? (b == 256 ? 1.0 : (b == 128 ? 1.1 : 1.5)) : (b == 64 ? 1.35 : 1.0); };
      auto cost = [&](int b) { return static_cast<double>((items(b) + sms - 1) / sms) * b * ineff(b); };
      double best = cost(bn);
      for (int b = bn / 2; b >= 64; b /= 2) {
        if (d->cout % b != 0) continue;
        const double c = cost(b);
        if (c < 0.85 * best) { best = c; bn = b; }
      }
    }
  }
  pl->block_n = bn;
  pl->mt = mt;
  pl->tma_store = (d->dtype != B2U_F32 && bn <= B2U_V2_TMA_MAX_BN) ? 1 : 0;      // the kernels' kTmaSt
  // pipeline depths: fill ~200 KB; the weight ring gets at least 3 stages, the patch ring at least 2
  const size_t budget = 222 * 1024 - 4096 - (pl->tma_store ? kV2StoreSlots : 0);
  int sa = 2, sb = 3;
  auto bytes = [&](int a, int b) { return static_cast<size_t>(a) * mt * kPatchStride + static_cast<size_t>(b) * bn * 128; };
  if (pro) {
    // fused prologue: a patch stage lives through TMA -> transform -> MMA, one more hop than the plain kernel: the patch
    // ring gets its third (and, if it fits, fourth) stage before the weight ring grows past four
    while (bytes(sa + 1, sb) <= budget && sa < 3) ++sa;
    while (bytes(sa, sb + 1) <= budget && sb < 4) ++sb;
    while (bytes(sa + 1, sb) <= budget && sa < 4) ++sa;
    while (bytes(sa, sb + 1) <= budget && sb < 9) ++sb;
  } else {
    while (bytes(sa, sb + 1) <= budget && sb < 6) ++sb;
    while (bytes(sa + 1, sb) <= budget && sa < 3) ++sa;
    while (bytes(sa, sb + 1) <= budget && sb < 9) ++sb;
  }
  if (d->reserved[1] >= 2 && d->reserved[1] <= 12) sb = d->reserved[1];
  // whole filter resident: one N tile (every item uses the same weights) and all 9 * (Cin / 64-or-32) stages fit
  {
    const int kc = d->cin / (d->dtype == B2U_F32 ? 32 : 64);
    pl->resident_b = (d->cout == bn && 9 * kc <= 12 && bytes(2, 9 * kc) <= budget && d->reserved[1] == 0) ? 1 : 0;
    if (pl->resident_b) {
      sb = 9 * kc;
      sa = 2;
      while (bytes(sa + 1, sb) <= budget && sa < (pro ? 4 : 3)) ++sa;
    }
  }
  B2U_REQUIRE(bytes(sa, sb) <= budget, "conv3x3 v2: pipeline does not fit shared memory (BLOCK_N %d MT %d)", bn, mt);
  pl->sa = sa;
  pl->sb = sb;
  pl->smem = bytes(sa, sb) + 1024 + (2 * sa + 2 * sb + 4) * 8 + 16 + 4 * 128 * 4 + sa * 8;
  pl->stage_off = 0;
  if (pl->tma_store) {
    pl->stage_off = static_cast<int>((pl->smem - 1024 + 1023) & ~static_cast<size_t>(1023));   // offset from the ALIGNED base
    pl->smem = static_cast<size_t>(pl->stage_off) + kV2StoreSlots + 1024;
  }
  pl->sgs = conv_stat_subgroup(d->cout, d->num_groups);
  return B2U_OK;
}

template <int BN, int MT, int TF, bool PRO>
static int v2_launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ty, const ConvV2Params& gp, int grid, size_t smem,
                     cudaStream_t st) {
  B2U_SET_MAX_SMEM_ONCE((conv3x3_v2_kernel<BN, MT, TF, PRO>), 227 * 1024);
  B2U_PDL_LAUNCH((conv3x3_v2_kernel<BN, MT, TF, PRO>), grid, PRO ? kV2ThreadsPro : kV2Threads, smem, st, ta, tb, ty, gp);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

int conv3x3_v2_stat_layout(const b2u_conv_desc* d, int* rows_per_image, int* subgroup_size) {
  int rc = conv_validate_desc(d);
  if (rc) return rc;
  V2Plan pl;
  rc = v2_make_plan(d, &pl);
  if (rc) return rc;
  if (rows_per_image) *rows_per_image = pl.tiles_w * pl.tiles_h;
  if (subgroup_size) *subgroup_size = pl.sgs;
  return B2U_OK;
}

int conv3x3_v2_run(const void* x, const void* wpacked, void* y, float* partials, const b2u_conv_desc* d, void* stream,
                   const V2Prologue* pro) {
  int rc = conv_validate_desc(d);
  if (rc) return rc;
  V2Plan pl;
  rc = v2_make_plan(d, &pl, pro != nullptr);
  if (rc) return rc;
  B2U_REQUIRE(x && wpacked && y, "null tensor pointer");
  B2U_REQUIRE(d->num_groups == 0 || partials != nullptr, "partials required when num_groups > 0");
  const int es = d->dtype == B2U_F32 ? 4 : 2;
  const int ke = 128 / es;
  CUtensorMap ta, tb;
  {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(d->x_cstride), static_cast<cuuint64_t>(d->w), static_cast<cuuint64_t>(d->h),
                          static_cast<cuuint64_t>(d->n)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(d->x_cstride) * es, static_cast<cuuint64_t>(d->w) * d->x_cstride * es,
                             static_cast<cuuint64_t>(d->h) * d->w * d->x_cstride * es};
    cuuint32_t box[4] = {static_cast<cuuint32_t>(ke), 10, 18, 1};
    rc = conv_encode_map(&ta, d->dtype, 4, x, dims, strides, box);
    if (rc) return rc;
  }
  {
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(d->cin), static_cast<cuuint64_t>(d->cout), 9};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(d->cin) * es, static_cast<cuuint64_t>(d->cout) * d->cin * es};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(ke), static_cast<cuuint32_t>(pl.block_n), 1};
    rc = conv_encode_map(&tb, d->dtype, 3, wpacked, dims, strides, box);
    if (rc) return rc;
  }
  CUtensorMap ty = ta;
  if (pl.tma_store) {
    // output [n][h][w][cout]: an epilogue warp stores 32 channels of 4 rows x 8 pixels (64-byte runs, 64-byte swizzle)
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(d->cout), static_cast<cuuint64_t>(d->w), static_cast<cuuint64_t>(d->h),
                          static_cast<cuuint64_t>(d->n)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(d->cout) * es, static_cast<cuuint64_t>(d->w) * d->cout * es,
                             static_cast<cuuint64_t>(d->h) * d->w * d->cout * es};
    cuuint32_t box[4] = {32, 8, 4, 1};
    rc = conv_encode_map(&ty, d->dtype, 4, y, dims, strides, box, 64);
    if (rc) return rc;
  }
  ConvV2Params gp;
  gp.n = d->n; gp.h = d->h; gp.w = d->w;
  gp.tiles_w = pl.tiles_w; gp.tiles_h = pl.tiles_h;
  gp.total_tiles = d->n * pl.tiles_w * pl.tiles_h;
  gp.num_mgroups = (gp.total_tiles + pl.mt - 1) / pl.mt;
  gp.num_items = gp.num_mgroups * (d->cout / pl.block_n);
  gp.kc_chunks = d->cin / ke;
  gp.cout = d->cout;
  gp.sa = pl.sa; gp.sb = pl.sb;
  gp.sgs_log2 = pl.sgs > 0 ? conv_ilog2(pl.sgs) : -1;
  gp.resident_b = pl.resident_b;
  gp.y = y;
  gp.partials = partials;
  gp.stage_off = pl.stage_off;
  gp.pro_coef = pro ? reinterpret_cast<const float2*>(pro->coef) : nullptr;
  gp.pro_mask = pro ? reinterpret_cast<const uint8_t*>(pro->mask) : nullptr;
  gp.pro_relu = pro ? pro->relu : 0;
  gp.x_shared = pro ? pro->x_shared : 0;
  gp.cin = d->cin;
  gp.fd_mgroups = fd_make(gp.num_mgroups);
  gp.fd_tpi = fd_make(pl.tiles_w * pl.tiles_h);
  gp.fd_tw = fd_make(pl.tiles_w);
  {
    static int env_combine = -1;
    if (env_combine < 0) {
      const char* e = getenv("B2U_COMBINE_STATS");
      env_combine = (e && e[0] == '0') ? 0 : 1;
    }
    // even tile count per image: an item's two tiles never straddle images and the pairing is the same for every
    // image whatever the batch (per-image results stay independent of the batch composition, bit for bit)
    gp.combine_stats = (env_combine && pl.mt == 2 && (pl.tiles_w * pl.tiles_h) % 2 == 0) ? 1 : 0;
  }
  if (pro) {
    B2U_REQUIRE(static_cast<double>(gp.num_items) * gp.num_mgroups < 4.0e9 && static_cast<double>(gp.total_tiles + 2) * pl.tiles_w * pl.tiles_h < 4.0e9,
                "fused prologue: tile count out of range for the fast division");
    B2U_REQUIRE(static_cast<double>(d->h) * d->w * (d->cin / 8) < 2.0e9, "fused prologue: image too large for 32-bit mask offsets");
    B2U_REQUIRE(d->dtype != B2U_F32, "the fused conv prologue supports the 16-bit storage formats (bf16 / fp16)");
    B2U_REQUIRE(pro->coef != nullptr, "fused prologue: coefficient pointer is NULL");
    B2U_REQUIRE(d->x_cstride == d->cin, "fused prologue: the input tensor must be dense in channels");
  }
  int grid = b2u_num_sms();
  if (grid > gp.num_items) grid = gp.num_items;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define B2U_V2_CASE(BN, MTV)                                                                        \
  if (pl.block_n == BN && pl.mt == MTV) {                                                            \
    if (pro) {                                                                                       \
      if (d->dtype == B2U_F16) return v2_launch<BN, MTV, 2, true>(ta, tb, ty, gp, grid, pl.smem, st);    \
      return v2_launch<BN, MTV, 0, true>(ta, tb, ty, gp, grid, pl.smem, st);                             \
    }                                                                                                \
    if (d->dtype == B2U_F32) return v2_launch<BN, MTV, 1, false>(ta, tb, ty, gp, grid, pl.smem, st);     \
    if (d->dtype == B2U_F16) return v2_launch<BN, MTV, 2, false>(ta, tb, ty, gp, grid, pl.smem, st);     \
    return v2_launch<BN, MTV, 0, false>(ta, tb, ty, gp, grid, pl.smem, st);                              \
  }
  B2U_V2_CASE(64, 1) B2U_V2_CASE(64, 2) B2U_V2_CASE(128, 1) B2U_V2_CASE(128, 2) B2U_V2_CASE(256, 1) B2U_V2_CASE(256, 2)
#undef B2U_V2_CASE
  b2u_set_error("conv3x3 v2: no kernel for BLOCK_N %d MT %d", pl.block_n, pl.mt);
  return B2U_ERR_UNSUPPORTED;
}

}  // namespace b2u
