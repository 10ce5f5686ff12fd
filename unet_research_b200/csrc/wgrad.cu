// Weight gradients on the tensor cores (sm_100a tcgen05): dW = G^T (x) X with K = pixels.
//
//   conv3x3: dW[tap][cg][cx] = sum_p G[p][cg] * X[p + tap][cx]     1x1: dW[cg][cx] = sum_p G[p][cg] * X[p][cx]
//
// Both operands come straight from the NHWC tensors (no transposition pass): a TMA box {64 channels, 8, 16} lands as
// 128 rows (pixels = K) x 128 B (64 channels = M or N), which is the canonical MN-major SWIZZLE_128B UMMA layout
// (8-pixel K atoms of 1024 B, 64-channel MN blocks).  The 9 taps reuse ONE (16+2)x(8+2) halo patch of X through
// shifted descriptors (start row (2*ks + r)*10 + s, K-atom stride 10 rows), exactly like the forward v2 kernel.
// A CTA owns a 128 x 64 (cg x cx) block for up to 8 taps (8 x 64 = 512 TMEM columns; the accumulators live for the
// whole CTA, there is no per-tile epilogue) and a slice of the pixel tiles (split-K); partial sums are written as
// fp32 and reduced in a fixed order by wgrad_reduce_kernel, which also permutes into the PyTorch weight layout.
//
// TAP PAIRS.  The B operand of an MN-major descriptor is a sequence of 64-element MN blocks LBO bytes apart.  Two taps
// of the same patch are two shifted views of it, a constant number of patch rows apart -- so ONE N = 128 MMA whose
// second MN block starts `distance` bytes after the first computes two taps at once: pairs (0,1), (3,4), (6,7)
// (one patch row = 128 B apart) and (2,5) (one patch line = 10 rows = 1280 B apart).  Half the MMAs, and the A operand
// (the gradient tile, 4 KB per K step) is read from shared memory once per pair instead of once per tap: an N = 64 MMA
// needs 192 B of operands per tensor cycle against the 128 B/cycle the SM delivers.
#include "b2u_common.cuh"
#include "conv_host.cuh"

#include <stdlib.h>

namespace b2u {

struct WgradParams {
  int n, h, w;
  int tiles_w, tiles_h, total_tiles;
  int tiles_per_slice, slices;
  int cg, cx;
  int taps;                  // 9 or 1
  int xchunks;               // cx / 64
  int stages;
  int tma_store;             // 1: the epilogue stages 32 x 32 fp32 blocks in the (by then idle) pipeline stages and writes them with TMA stores
  float* ws;                 // [slices][taps][cg][cx]
};

constexpr int kWgThreads = 192;
constexpr int kGBytes = 2 * 128 * 128;            // two 64-channel boxes of 128 pixels
constexpr int kXPatchStride = 23552;

// MN-major SWIZZLE_128B descriptor: lbo = byte distance between 64-element MN blocks, sbo = between 8-row K atoms
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

#ifndef B2U_WGRAD_PAIR_TAPS
#define B2U_WGRAD_PAIR_TAPS 1
#endif
constexpr bool kPairTaps = B2U_WGRAD_PAIR_TAPS != 0;

template <int TAPS>   // 9: 3x3 (tap groups of 8 + 1 over blockIdx.z), 1: plain
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
             const WgradParams p) {
  // bf16 A/B both MN-major (bits 15, 16), fp32 accumulate, M = 128, N = 64
  constexpr uint32_t kIdesc = umma_idesc(128, 64, 1) | (1u << 15) | (1u << 16);
  constexpr uint32_t kIdescPair = umma_idesc(128, 128, 1) | (1u << 15) | (1u << 16);
  constexpr int kXBytes = TAPS == 9 ? 180 * 128 : 128 * 128;
  constexpr int kXStride = TAPS == 9 ? kXPatchStride : 128 * 128;
  constexpr int kStageBytes = kGBytes + kXStride;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = p.stages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + stages * kStageBytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* done_bar = empty_bar + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x;
  const int mblk = blockIdx.y / p.xchunks, xc = blockIdx.y - mblk * p.xchunks;
  const int m0 = mblk * 128, c0 = xc * 64;
  const int tap_begin = TAPS == 9 ? (blockIdx.z == 0 ? 0 : 8) : 0;
  const int tap_count = TAPS == 9 ? (blockIdx.z == 0 ? 8 : 1) : 1;
  const int t_begin = slice * p.tiles_per_slice;
  const int t_end = min(t_begin + p.tiles_per_slice, p.total_tiles);
  const int tiles_per_image = p.tiles_w * p.tiles_h;
  constexpr int kTmemCols = TAPS == 9 ? 512 : 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmG);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(done_bar, 1);
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
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;
      for (int tile = t_begin; tile < t_end; ++tile) {
        const int img = tile / tiles_per_image;
        const int r = tile - img * tiles_per_image;
        const int ty = r / p.tiles_w, tx = r - ty * p.tiles_w;
        mbar_wait(&empty_bar[s], ph);
        mbar_arrive_expect_tx(&full_bar[s], kGBytes + kXBytes);
        uint8_t* st = smem + s * kStageBytes;
        tma_load_4d(st, &tmG, &full_bar[s], m0, tx * 8, ty * 16, img);
        tma_load_4d(st + 128 * 128, &tmG, &full_bar[s], m0 + 64, tx * 8, ty * 16, img);   // zero-filled when cg == 64
        if (TAPS == 9) tma_load_4d(st + kGBytes, &tmX, &full_bar[s], c0, tx * 8 - 1, ty * 16 - 1, img);
        else tma_load_4d(st + kGBytes, &tmX, &full_bar[s], c0, tx * 8, ty * 16, img);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // the whole warp runs the loop (converged control flow); one elected lane issues (see umma_ss_conv)
    {
      const uint32_t leader = elect_one() ? 1u : 0u;
      const uint32_t tmem_u = warp_uniform(tmem_base);
      int s = 0;
      uint32_t ph = 0;
      bool first = true;
      for (int tile = t_begin; tile < t_end; ++tile) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t g_addr = smem_u32(smem + s * kStageBytes);
        const uint32_t x_addr = g_addr + kGBytes;
        if (TAPS == 9 && blockIdx.z == 0 && kPairTaps) {
#pragma unroll
          for (int pr = 0; pr < 4; ++pr) {
            // pair pr = taps (a, b): TMEM columns [pr * 128, +64) = tap a, [pr * 128 + 64, +64) = tap b
            const int ta = pr < 3 ? 3 * pr : 2;
            const int dist = pr < 3 ? 128 : 1280;                  // bytes between the two views
            const int row_off = (ta / 3) * 10 + (ta % 3);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint64_t adesc = umma_desc_mn_sw128(g_addr + ks * 2048, 128 * 128, 1024);
              const uint64_t bdesc = umma_desc_mn_sw128(x_addr + (row_off + ks * 20) * 128, dist, 1280);
              umma_ss_conv<false>(tmem_u + pr * 128, adesc, bdesc, kIdescPair, (first && ks == 0) ? 0u : 1u, leader);
            }
          }
        } else
        for (int t = 0; t < tap_count; ++t) {
          const int tap = tap_begin + t;
          const int row_off = TAPS == 9 ? (tap / 3) * 10 + (tap % 3) : 0;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {                      // 16 pixels (two tile rows) per MMA
            const uint64_t adesc = umma_desc_mn_sw128(g_addr + ks * 2048, 128 * 128, 1024);
            const uint64_t bdesc = TAPS == 9 ? umma_desc_mn_sw128(x_addr + (row_off + ks * 20) * 128, 128 * 128, 1280)
                                             : umma_desc_mn_sw128(x_addr + ks * 2048, 128 * 128, 1024);
            umma_ss_conv<false>(tmem_u + t * 64, adesc, bdesc, kIdesc, (first && ks == 0) ? 0u : 1u, leader);
          }
        }
        first = false;
        umma_commit_conv(&empty_bar[s], leader);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
      umma_commit_conv(done_bar, leader);
    }
    __syncwarp();
  } else {
    // epilogue: once per CTA, fp32 partials out
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const bool valid = (m0 + row) < p.cg && t_begin < t_end;
    // TMA-store path: a thread owns one accumulator ROW (one gradient channel), i.e. a 128-byte run of fp32 per 32
    // columns, and neighbouring threads own rows cx * 4 bytes apart -- a warp-wide STG.128 touched 32 lines with 16
    // bytes each, 16 of them per tap and thread, and this epilogue is fully exposed (the accumulators live for the whole
    // CTA).  Every warp now stages its 32 rows x 128 bytes in one of two 4 KB slots carved out of the pipeline stages (all
    // MMAs have completed: done_bar) under the 128-byte swizzle and writes the block with one TMA store.
    const bool tma_st = p.tma_store != 0 && t_begin < t_end;
    const uint32_t slot0 = smem_u32(smem) + static_cast<uint32_t>(q * 8192);
    int nstore = 0;
    for (int t = 0; t < tap_count; ++t) {
      // paired layout: TMEM column block t holds tap [0, 1, 3, 4, 6, 7, 2, 5][t]
      const int tap = (TAPS == 9 && blockIdx.z == 0 && kPairTaps) ? (t < 6 ? 3 * (t >> 1) + (t & 1) : (t == 6 ? 2 : 5)) : tap_begin + t;
      float* dst = p.ws + ((static_cast<size_t>(slice) * p.taps + tap) * p.cg + (m0 + row)) * p.cx + c0;
#pragma unroll 1
      for (int chunk = 0; chunk < 2; ++chunk) {
        uint32_t rr[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * 64 + chunk * 32, rr);
        tmem_ld_wait();
        if (tma_st) {
          const uint32_t slot = slot0 + static_cast<uint32_t>((nstore & 1) * 4096);
          if (lane == 0) bulk_wait_read<1>();                 // the store issued two blocks ago has read this slot
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i)
            sts128(slot + static_cast<uint32_t>(lane * 128 + ((i ^ (lane & 7)) << 4)), make_uint4(rr[4 * i], rr[4 * i + 1], rr[4 * i + 2], rr[4 * i + 3]));
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&tmW)), "r"(slot), "r"(c0 + chunk * 32),
                           "r"((slice * p.taps + tap) * p.cg + m0 + q * 32)
                         : "memory");
            bulk_commit();
          }
          ++nstore;
        } else if (valid) {
          float4* d4 = reinterpret_cast<float4*>(dst + chunk * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            d4[i] = make_float4(__uint_as_float(rr[4 * i]), __uint_as_float(rr[4 * i + 1]), __uint_as_float(rr[4 * i + 2]),
                                __uint_as_float(rr[4 * i + 3]));
        }
      }
    }
    if (tma_st && lane == 0) bulk_wait<0>();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// Cg = 64 (the three weight gradients of the full-resolution level, the most expensive ones at batch 1): with the
// gradient tile as the M = 128 operand half of every MMA multiplies zero rows, and the ninth tap needs its own CTA group.
// Here the roles are SWAPPED: A = two taps of the activation patch stacked along M (two 64-channel MN blocks LBO = the
// distance between the two shifted views apart), B = the 64 gradient channels (N = 64).  Five MMAs per K step cover the
// nine taps -- pairs (0,1), (3,4), (6,7), (2,5) and (7,8), the last one recomputing tap 7 in rows it then ignores --
// in 5 x 64 = 320 TMEM columns: one CTA owns all nine taps, M is full, the gradient tile is a single 16 KB box.
// D[row = (tap of the pair, cx), column = cg]; the epilogue writes ws[slice][tap][cg][cx] with lanes along cx (coalesced).
constexpr int kSwapStageBytes = 128 * 128 + kXPatchStride;       // one gradient box + one halo patch
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_swap64_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX, const WgradParams p) {
  constexpr uint32_t kIdesc = umma_idesc(128, 64, 1) | (1u << 15) | (1u << 16);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = p.stages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + stages * kSwapStageBytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* done_bar = empty_bar + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x;
  const int c0 = blockIdx.y * 64;
  const int t_begin = slice * p.tiles_per_slice;
  const int t_end = min(t_begin + p.tiles_per_slice, p.total_tiles);
  const int tiles_per_image = p.tiles_w * p.tiles_h;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmG);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;
      for (int tile = t_begin; tile < t_end; ++tile) {
        const int img = tile / tiles_per_image;
        const int r = tile - img * tiles_per_image;
        const int ty = r / p.tiles_w, tx = r - ty * p.tiles_w;
        mbar_wait(&empty_bar[s], ph);
        mbar_arrive_expect_tx(&full_bar[s], 128 * 128 + 180 * 128);
        uint8_t* st = smem + s * kSwapStageBytes;
        tma_load_4d(st, &tmG, &full_bar[s], 0, tx * 8, ty * 16, img);
        tma_load_4d(st + 128 * 128, &tmX, &full_bar[s], c0, tx * 8 - 1, ty * 16 - 1, img);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t tmem_u = warp_uniform(tmem_base);
    int s = 0;
    uint32_t ph = 0;
    bool first = true;
    for (int tile = t_begin; tile < t_end; ++tile) {
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      const uint32_t g_addr = smem_u32(smem + s * kSwapStageBytes);
      const uint32_t x_addr = g_addr + 128 * 128;
#pragma unroll
      for (int pr = 0; pr < 5; ++pr) {
        const int ta = pr < 3 ? 3 * pr : (pr == 3 ? 2 : 7);       // first tap of the pair
        const int dist = pr == 3 ? 1280 : 128;                     // bytes between the two views (second tap: ta + 1, or ta + 3)
        const int row_off = (ta / 3) * 10 + (ta % 3);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t adesc = umma_desc_mn_sw128(x_addr + (row_off + ks * 20) * 128, dist, 1280);
          const uint64_t bdesc = umma_desc_mn_sw128(g_addr + ks * 2048, 128 * 128, 1024);
          umma_ss_conv<false>(tmem_u + pr * 64, adesc, bdesc, kIdesc, (first && ks == 0) ? 0u : 1u, leader);
        }
      }
      first = false;
      umma_commit_conv(&empty_bar[s], leader);
      if (++s == stages) { s = 0; ph ^= 1; }
    }
    umma_commit_conv(done_bar, leader);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int half = row >> 6, cxl = row & 63;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    if (t_begin < t_end) {
#pragma unroll 1
      for (int pr = 0; pr < 5; ++pr) {
        const int ta = pr < 3 ? 3 * pr : (pr == 3 ? 2 : 7);
        const int tap = half ? (pr == 3 ? 5 : ta + 1) : ta;
        const bool write = !(pr == 4 && half == 0);               // tap 7 was already written by pair 2
        float* dst = p.ws + ((static_cast<size_t>(slice) * 9 + tap) * 64) * p.cx + c0 + cxl;
#pragma unroll 1
        for (int chunk = 0; chunk < 2; ++chunk) {
          uint32_t rr[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + pr * 64 + chunk * 32, rr);
          tmem_ld_wait();
          if (write) {
#pragma unroll
            for (int i = 0; i < 32; ++i) dst[static_cast<size_t>(chunk * 32 + i) * p.cx] = __uint_as_float(rr[i]);
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dw = sum over slices of ws, permuted to the PyTorch layout.
// layout 0: Conv2d [cg][cx][3][3] (taps 9) ; layout 1: ConvTranspose2d [cx][cout][2][2] with cg = 4*cout, tap-major (taps 1)
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int slices, int taps, int cg, int cx,
                                    int layout, int empty_slices_from) {
  pdl_wait();       // predecessor grid complete + visible (no-op without the PDL launch attribute)
  pdl_trigger();    // the successor may be scheduled once every CTA got here
  const long total = static_cast<long>(taps) * cg * cx;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % cx);
    const int g = static_cast<int>((i / cx) % cg);
    const int tap = static_cast<int>(i / (static_cast<long>(cx) * cg));
    float acc = 0.f;
    for (int s = 0; s < slices && s < empty_slices_from; ++s) acc += ws[static_cast<long>(s) * total + i];
    long dst;
    if (layout == 0) {
      dst = (static_cast<long>(g) * cx + x) * taps + tap;
    } else {
      const int cout = cg / 4;
      const int t4 = g / cout, co = g - t4 * cout;
      dst = (static_cast<long>(x) * cout + co) * 4 + t4;
    }
    dw[dst] = acc;
  }
}

// layout 0 with 9 taps: one thread per (cg, cx) pair sums the slices of all nine taps (reads coalesced along cx)
// and writes its nine taps as one 36-byte run, so a warp writes 1152 contiguous bytes of the PyTorch layout.
__global__ void __launch_bounds__(1024) wgrad_reduce9_kernel(const float* __restrict__ ws, float* __restrict__ dw, int slices, int cg, int cx) {
  pdl_wait();       // predecessor grid complete + visible (no-op without the PDL launch attribute)
  pdl_trigger();    // the successor may be scheduled once every CTA got here
  // blockDim = (PX pairs, SG slice groups), PX * SG = 256 .. 1024 threads: slice group y sums slices y, y + SG, ...; the
  // groups are then added in a fixed order through (dynamic) shared memory (deterministic), and thread (x, 0) writes the
  // nine taps.  Many slices / few pairs (full-resolution layers: 148 slices of 4096 pairs) -> up to 32 groups; few slices
  // / many pairs (deep layers) -> 128 .. 256 pairs per block instead of one-warp blocks.
  extern __shared__ float part_raw[];       // [nsg][PX][9]
  const int px = blockDim.x;
  const long pairs = static_cast<long>(cg) * cx;
  const long total = 9 * pairs;
  const long i = static_cast<long>(blockIdx.x) * px + threadIdx.x;
  const int sg = threadIdx.y, nsg = blockDim.y;
  float acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = 0.f;
  if (i < pairs) {
    for (int s = sg; s < slices; s += nsg) {
      const float* src = ws + static_cast<long>(s) * total + i;
#pragma unroll
      for (int t = 0; t < 9; ++t) acc[t] += src[t * pairs];
    }
  }
  if (nsg > 1) {
    float* mine = part_raw + (static_cast<size_t>(sg) * px + threadIdx.x) * 9;
#pragma unroll
    for (int t = 0; t < 9; ++t) mine[t] = acc[t];
    __syncthreads();
    if (sg == 0) {
      for (int y = 1; y < nsg; ++y) {
        const float* other = part_raw + (static_cast<size_t>(y) * px + threadIdx.x) * 9;
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[t] += other[t];
      }
    }
  }
  if (sg == 0 && i < pairs) {
    float* dst = dw + i * 9;
#pragma unroll
    for (int t = 0; t < 9; ++t) dst[t] = acc[t];
  }
}

// First layer: dW[co][ci][tap] = sum_p g[p][co] * x[p + tap][ci] on CUDA cores (K = 9*Cin <= 27 columns of output).
// grid = (blocks, n): each block reduces a pixel range to ws[n][block][cout][cin*9]; a second kernel sums the rows.
template <typename T, int CIN>
__global__ void __launch_bounds__(256) wgrad_first_kernel(const T* __restrict__ g, const float* __restrict__ x, float* __restrict__ ws, int h0,
                                   int w0, int h, int w, int cout) {
  pdl_wait();       // predecessor grid complete + visible (no-op without the PDL launch attribute)
  pdl_trigger();    // the successor may be scheduled once every CTA got here
  extern __shared__ float sm[];
  const int n = blockIdx.y;
  const int cvs = cout >> 3;
  const int cv = threadIdx.x % cvs, slot = threadIdx.x / cvs, slots = blockDim.x / cvs;
  const float* xn = x + static_cast<size_t>(n) * CIN * h0 * w0;
  float acc[8][CIN * 9];
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int j = 0; j < CIN * 9; ++j) acc[k][j] = 0.f;
  const int npix = h * w;
  for (int pix = blockIdx.x * slots + slot; pix < npix; pix += gridDim.x * slots) {
    const int ph = pix / w, pw = pix - ph * w;
    float in[CIN * 9];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int yy = ph + r - 1, xx = pw + c - 1;
          in[ci * 9 + r * 3 + c] = (yy >= 0 && yy < h0 && xx >= 0 && xx < w0) ? __ldg(xn + (static_cast<size_t>(ci) * h0 + yy) * w0 + xx) : 0.f;
        }
    float gv[8];
    Vec8<T> v;
    v.load(g + (static_cast<size_t>(n) * npix + pix) * cout + cv * 8);
    v.to_float(gv);
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int j = 0; j < CIN * 9; ++j) acc[k][j] = fmaf(gv[k], in[j], acc[k][j]);
  }
  // deterministic block reduction over the pixel slots, one (co, j) at a time through shared memory
  constexpr int J = CIN * 9;
  float* out_row = ws + (static_cast<size_t>(n) * gridDim.x + blockIdx.x) * cout * J;
  for (int j = 0; j < J; ++j) {
#pragma unroll
    for (int k = 0; k < 8; ++k) sm[threadIdx.x * 8 + k] = acc[k][j];
    __syncthreads();
    for (int o = threadIdx.x; o < cout; o += blockDim.x) {
      const int ccv = o >> 3, k = o & 7;
      float a = 0.f;
      for (int sl = 0; sl < slots; ++sl) a += sm[(sl * cvs + ccv) * 8 + k];
      out_row[o * J + j] = a;
    }
    __syncthreads();
  }
}
__global__ void __launch_bounds__(256) wgrad_first_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int rows, int elems) {
  pdl_wait();       // predecessor grid complete + visible (no-op without the PDL launch attribute)
  pdl_trigger();    // the successor may be scheduled once every CTA got here
  // a block owns 32 consecutive weight elements (lane = element: coalesced rows); warp w adds rows w, w + 8, ... in fp64,
  // the eight warp sums are added in a fixed order (deterministic).  The 3-block version walked all 592 partial rows
  // serially per thread: 80 us for 1.4 MB.
  __shared__ double part[8][32];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  double a = 0.0;
  if (i < elems)
    for (int r = wrp; r < rows; r += 8) a += ws[static_cast<size_t>(r) * elems + i];
  part[wrp][lane] = a;
  __syncthreads();
  if (wrp == 0 && i < elems) {
    double t = part[0][lane];
#pragma unroll
    for (int k = 1; k < 8; ++k) t += part[k][lane];
    dw[i] = static_cast<float>(t);
  }
}

template <typename T>
__global__ void pack_convT_dgrad_kernel(const float* __restrict__ w, T* __restrict__ out, int cin, int cout) {
  pdl_wait();       // predecessor grid complete + visible (no-op without the PDL launch attribute)
  pdl_trigger();    // the successor may be scheduled once every CTA got here
  const long total = 4L * cout * cin;          // out[ci][tap*cout + co] = w[ci][co][tap]
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % (4 * cout));
    const int ci = static_cast<int>(i / (4 * cout));
    const int tap = k / cout, co = k - tap * cout;
    out[i] = to_operand<T>(w[(static_cast<long>(ci) * cout + co) * 4 + tap]);
  }
}

struct WgPlan {
  int tiles_w, tiles_h, total_tiles, slices, tiles_per_slice, mblks, xchunks, groups, stages, swap64;
  size_t smem;
};

// Cg = 64, 3x3: the operand-swapped kernel (B2U_WGRAD_SWAP64=0 selects the generic one for A/B runs)
static bool wg_use_swap64(const b2u_wgrad_desc* d) {
  if (d->taps != 9 || d->cg != 64 || d->layout != 0) return false;
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("B2U_WGRAD_SWAP64");
    env = (e && e[0] == '0') ? 0 : 1;
  }
  return env != 0;
}

static int wg_plan(const b2u_wgrad_desc* d, WgPlan* pl) {
  B2U_REQUIRE(d, "null descriptor");
  B2U_REQUIRE(d->dtype == B2U_BF16, "wgrad is implemented for bf16 only");
  B2U_REQUIRE(d->taps == 9 || d->taps == 1, "taps must be 9 or 1");
  B2U_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0, "empty tensor");
  B2U_REQUIRE(d->cg > 0 && d->cg % 64 == 0 && d->cx > 0 && d->cx % 64 == 0, "channel counts must be multiples of 64 (cg %d cx %d)", d->cg, d->cx);
  B2U_REQUIRE(d->x_cstride >= d->cx && d->x_cstride % 8 == 0, "bad x_cstride");
  B2U_REQUIRE(d->layout == 0 || (d->layout == 1 && d->taps == 1 && d->cg % 4 == 0), "bad layout");
  pl->tiles_w = (d->w + 7) / 8;
  pl->tiles_h = (d->h + 15) / 16;
  pl->total_tiles = d->n * pl->tiles_w * pl->tiles_h;
  pl->mblks = (d->cg + 127) / 128;
  pl->xchunks = d->cx / 64;
  pl->groups = d->taps == 9 ? 2 : 1;
  // the second tap group (tap 8 alone) is 1/8 of the work of the first: size the split for the heavy CTAs only
  const int pairs = pl->mblks * pl->xchunks;
  // split the pixel tiles only as far as needed to occupy every SM once: each extra slice costs a full fp32 copy of
  // the weight gradient in HBM (75 MB for the 1024 -> 1024 layer)
  int slices = b2u_num_sms() / pairs;                       // floor: slices * pairs CTAs fit one wave
  if (slices > pl->total_tiles) slices = pl->total_tiles;
  if (slices < 1) slices = 1;
  pl->tiles_per_slice = (pl->total_tiles + slices - 1) / slices;
  pl->slices = (pl->total_tiles + pl->tiles_per_slice - 1) / pl->tiles_per_slice;
  pl->stages = 3;
  int stage_bytes = kGBytes + (d->taps == 9 ? kXPatchStride : 128 * 128);
  pl->swap64 = wg_use_swap64(d) ? 1 : 0;
  if (pl->swap64) {
    pl->groups = 1;
    pl->stages = 4;
    stage_bytes = kSwapStageBytes;
  }
  pl->smem = static_cast<size_t>(pl->stages) * stage_bytes + 1024 + (2 * pl->stages + 1) * 8 + 16;
  return B2U_OK;
}

}  // namespace b2u

using namespace b2u;

extern "C" int b2u_wgrad_workspace_floats(const b2u_wgrad_desc* d, long long* floats) {
  WgPlan pl;
  int rc = wg_plan(d, &pl);
  if (rc) return rc;
  B2U_REQUIRE(floats, "null output");
  *floats = static_cast<long long>(pl.slices) * d->taps * d->cg * d->cx;
  return B2U_OK;
}

extern "C" int b2u_wgrad(const void* g, const void* x, float* workspace, float* dw, const b2u_wgrad_desc* d, void* stream) {
  WgPlan pl;
  int rc = wg_plan(d, &pl);
  if (rc) return rc;
  B2U_REQUIRE(g && x && workspace && dw, "null pointer");
  CUtensorMap tg, tx;
  {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(d->cg), static_cast<cuuint64_t>(d->w), static_cast<cuuint64_t>(d->h), static_cast<cuuint64_t>(d->n)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(d->cg) * 2, static_cast<cuuint64_t>(d->w) * d->cg * 2,
                             static_cast<cuuint64_t>(d->h) * d->w * d->cg * 2};
    cuuint32_t box[4] = {64, 8, 16, 1};
    rc = conv_encode_map(&tg, B2U_BF16, 4, g, dims, strides, box);
    if (rc) return rc;
  }
  {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(d->x_cstride), static_cast<cuuint64_t>(d->w), static_cast<cuuint64_t>(d->h), static_cast<cuuint64_t>(d->n)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(d->x_cstride) * 2, static_cast<cuuint64_t>(d->w) * d->x_cstride * 2,
                             static_cast<cuuint64_t>(d->h) * d->w * d->x_cstride * 2};
    cuuint32_t box9[4] = {64, 10, 18, 1};
    cuuint32_t box1[4] = {64, 8, 16, 1};
    rc = conv_encode_map(&tx, B2U_BF16, 4, x, dims, strides, d->taps == 9 ? box9 : box1);
    if (rc) return rc;
  }
  CUtensorMap tw = tg;
  // workspace [slices * taps * cg rows][cx] fp32: 32 x 32 blocks (128-byte rows).  Rows of a block must not run past the
  // tensor's cg rows of one (slice, tap): the 128-row M block is full only when cg is a multiple of 128.
  static int env_tma = -1;
  if (env_tma < 0) {
    const char* e = getenv("B2U_WGRAD_TMA_STORE");
    env_tma = (e && e[0] == '0') ? 0 : 1;
  }
  const bool tma_store = env_tma && !pl.swap64 && d->cg % 128 == 0 &&
                         static_cast<long long>(pl.slices) * d->taps * d->cg < (1ll << 31);
  if (tma_store) {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(d->cx), static_cast<cuuint64_t>(pl.slices) * d->taps * d->cg};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(d->cx) * 4};
    cuuint32_t box[2] = {32, 32};
    rc = conv_encode_map(&tw, B2U_F32, 2, workspace, dims, strides, box);
    if (rc) return rc;
  }
  WgradParams p;
  p.tma_store = tma_store ? 1 : 0;
  p.n = d->n; p.h = d->h; p.w = d->w;
  p.tiles_w = pl.tiles_w; p.tiles_h = pl.tiles_h; p.total_tiles = pl.total_tiles;
  p.tiles_per_slice = pl.tiles_per_slice; p.slices = pl.slices;
  p.cg = d->cg; p.cx = d->cx; p.taps = d->taps; p.xchunks = pl.xchunks; p.stages = pl.stages;
  p.ws = workspace;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(pl.slices, pl.mblks * pl.xchunks, pl.groups);
  if (pl.swap64) {
    B2U_SET_MAX_SMEM_ONCE((wgrad_swap64_kernel), 227 * 1024);
    B2U_PDL_LAUNCH((wgrad_swap64_kernel), grid, kWgThreads, pl.smem, st, tg, tx, p);
  } else if (d->taps == 9) {
    B2U_SET_MAX_SMEM_ONCE((wgrad_kernel<9>), 227 * 1024);
    B2U_PDL_LAUNCH((wgrad_kernel<9>), grid, kWgThreads, pl.smem, st, tg, tx, tw, p);
  } else {
    B2U_SET_MAX_SMEM_ONCE((wgrad_kernel<1>), 227 * 1024);
    B2U_PDL_LAUNCH((wgrad_kernel<1>), grid, kWgThreads, pl.smem, st, tg, tx, tw, p);
  }
  B2U_LAUNCH_CHECK();
  const long total = static_cast<long>(d->taps) * d->cg * d->cx;
  {
    static int skip_reduce = -1;                            // TIMING DIAGNOSTIC ONLY (stale gradients): tests/exp_train_skip.py
    if (skip_reduce < 0) skip_reduce = getenv("B2U_EXP_SKIP_WGRAD_REDUCE") ? 1 : 0;
    if (skip_reduce) return B2U_OK;
  }
  if (d->taps == 9 && d->layout == 0) {
    const long pairs = static_cast<long>(d->cg) * d->cx;
    // slice groups per block: enough threads in flight for the layers with many slices and few (cg, cx) pairs (64 x 64
    // at 148 slices: 128 blocks; with 8 groups every thread walked 19 slices x 9 taps serially, 26 us for 22 MB)
    int nsg = 1;
    while (nsg < 32 && nsg * 2 <= pl.slices && (nsg < 8 || pairs * nsg < 148L * 1024 * 2)) nsg *= 2;
    int px = 32;
    while (px * nsg < 256 && px < 256) px *= 2;               // at least 8 warps per block
    const size_t smem = nsg > 1 ? static_cast<size_t>(nsg) * px * 9 * sizeof(float) : 0;
    B2U_PDL_LAUNCH((wgrad_reduce9_kernel), static_cast<unsigned>((pairs + px - 1) / px), dim3(px, nsg), smem, st, workspace, dw, pl.slices, d->cg, d->cx);
  } else {
    int blocks = static_cast<int>((total + 255) / 256 < 2048 ? (total + 255) / 256 : 2048);
    B2U_PDL_LAUNCH((wgrad_reduce_kernel), blocks, 256, 0, st, workspace, dw, pl.slices, d->taps, d->cg, d->cx, d->layout, pl.slices);
  }
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

// pixel-range blocks per image of the first-layer weight gradient (= partial rows of its workspace): four 256-thread
// blocks per SM (64 blocks left most of the GPU idle: 165 us for a 44 MB read)
extern "C" int b2u_wgrad_first_rows(void) { return 4 * b2u_num_sms(); }

extern "C" int b2u_wgrad_first(const void* g, const float* x_nchw, float* workspace, float* dw, int n, int cin, int h0, int w0,
                               int h, int w, int cout, int dtype, void* stream) {
  B2U_REQUIRE(g && x_nchw && workspace && dw, "null pointer");
  B2U_REQUIRE(cin == 1 || cin == 3, "cin must be 1 or 3");
  B2U_REQUIRE(cout % 8 == 0 && cout / 8 <= 32, "cout must be a multiple of 8, at most 256");
  B2U_REQUIRE(dtype == B2U_BF16, "bf16 only");
  const int cvs = cout / 8;
  const int threads = (256 / cvs) * cvs;
  const int rows = b2u_wgrad_first_rows();                 // workspace: float[n][rows][cout][cin*9]
  dim3 grid(rows, n);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t smem = static_cast<size_t>(threads) * 8 * sizeof(float);
  if (cin == 1) B2U_PDL_LAUNCH((wgrad_first_kernel<__nv_bfloat16, 1>), grid, threads, smem, st, static_cast<const __nv_bfloat16*>(g), x_nchw, workspace, h0, w0, h, w, cout);
  else B2U_PDL_LAUNCH((wgrad_first_kernel<__nv_bfloat16, 3>), grid, threads, smem, st, static_cast<const __nv_bfloat16*>(g), x_nchw, workspace, h0, w0, h, w, cout);
  B2U_LAUNCH_CHECK();
  const int elems = cout * cin * 9;
  B2U_PDL_LAUNCH((wgrad_first_reduce_kernel), (elems + 31) / 32, 256, 0, st, workspace, dw, rows * n, elems);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_pack_convT2x2_dgrad_weight(const float* w, void* packed, int cin, int cout, int dtype, void* stream) {
  B2U_REQUIRE(w && packed && cin > 0 && cout > 0, "bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long total = 4L * cout * cin;
  const int blocks = static_cast<int>((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  if (dtype == B2U_F32) B2U_PDL_LAUNCH((pack_convT_dgrad_kernel<float>), blocks, 256, 0, st, w, static_cast<float*>(packed), cin, cout);
  else if (dtype == B2U_F16) B2U_PDL_LAUNCH((pack_convT_dgrad_kernel<__half>), blocks, 256, 0, st, w, static_cast<__half*>(packed), cin, cout);
  else B2U_PDL_LAUNCH((pack_convT_dgrad_kernel<__nv_bfloat16>), blocks, 256, 0, st, w, static_cast<__nv_bfloat16*>(packed), cin, cout);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}
