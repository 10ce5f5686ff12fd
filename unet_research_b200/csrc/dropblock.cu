// DropBlock mask construction (DropBlock2D.forward, reference utils_modules.py:46-58) as two
// integer kernels over a table of calls:
//
//  centers : reproduces `torch.rand(N,C,H-bs+1,W-bs+1, device='cuda') < gamma` bit-exactly for a given
//            (seed, philox offset): Philox4x32-10 keyed by the seed, counter = (offset/4 + trip, 0, idx, 0),
//            thread idx's trip-t draw supplies elements idx + Tn*(4t+ii), ii = 0..3, Tn = 256*grid
//            (ATen/native/cuda/DistributionTemplates.h:65-90).  One warp evaluates 32 consecutive idx,
//            compares the raw 32-bit words against host-computed thresholds (equivalent to the fp32
//            compare incl. the 1.0 -> 0.0 wrap) and writes four ballot words of a flat bitmap.
//  dilate  : block_mask = 1 - maxpool_{bs x bs, stride 1, pad bs/2}(zero_pad_{bs/2}(centre)) becomes
//            drop(h,w) = OR_{i,j<bs} centre[h-i][w-j]: a horizontal bit-smear, a vertical OR over a sliding
//            window of bs rows, a 32x32 bit transpose (channel-major -> NHWC pixel-major) with warp ballots
//            and a popcount for block_mask.sum().
#include "b2u_common.cuh"

#include <stdlib.h>

namespace b2u {

// 32x32 -> 64 multiply as ONE IMAD.WIDE.U32 (the C++ uint64 product makes ptxas add a zero high-word
// correction: 2 wasted adds per round, 20 per call)
__device__ __forceinline__ void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef B2U_PHILOX_SPLIT_MUL
  hi = __umulhi(a, b);
  lo = a * b;
#else
  asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
#endif
}

// rounds R0 .. 9 of Philox4x32-10; k0 / k1 are the ORIGINAL key words (round r uses key + r * Weyl constant)
template <int R0>
__device__ __forceinline__ void philox_rounds(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
  k0 += static_cast<uint32_t>(R0) * 0x9E3779B9u;
  k1 += static_cast<uint32_t>(R0) * 0xBB67AE85u;
#pragma unroll
  for (int r = R0; r < 10; ++r) {
    uint32_t h0, l0, h1, l1;
    mulhilo(0xD2511F53u, c0, h0, l0);
    mulhilo(0xCD9E8D57u, c2, h1, l1);
    const uint32_t n0 = h1 ^ c1 ^ k0;
    const uint32_t n2 = h0 ^ c3 ^ k1;
    c1 = l1;
    c3 = l0;
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
  philox_rounds<0>(c0, c1, c2, c3, k0, k1, out);
}

// grid = (SMs * 8, n_calls) blocks of 256 threads: block b of call k, thread t IS torch's thread idx = 256*b + t for
// every trip, so there is no index arithmetic beyond one add per trip; blocks beyond the call's torch grid exit.
// Blocks are SHORT on purpose: next to the high-priority forward pass the block scheduler hands freed SM slots to the
// forward's CTAs first, so the mask build only fills what the forward leaves idle (a persistent grid keeps its warps
// resident and was measured to destroy that overlap: 11.4 instead of 8.8 ms per Monte-Carlo step).
constexpr uint32_t kMaxFastTrips = 64;

__global__ void __launch_bounds__(256) dropblock_centers_kernel(const b2u_dropblock_call* __restrict__ table, uint64_t seed,
                                         const unsigned long long* __restrict__ offset_base,
                                         uint32_t* __restrict__ center_bits, uint32_t trips_per_block) {
  pdl_wait();       // predecessor grid complete + visible (no-op without the PDL launch attribute)
  pdl_trigger();    // the successor may be scheduled once every CTA got here
  const b2u_dropblock_call c = table[blockIdx.y];
  if (blockIdx.x >= c.grid) return;
  const uint32_t tn = c.grid * 256u;
  const uint32_t trips = (c.numel - 1u) / (tn * 4u) + 1u;     // numel <= 2^32-1, tn*4 <= 2^21 * ... fits
  // blockIdx.z splits a thread's trips into runs of trips_per_block (0 = all): a block lives ~1 us instead of up to ~10 us (36
  // trips at the largest site).  The forward's kernels (high-priority stream) can only start on an SM once resident
  // low-priority blocks have drained -- there is no preemption -- so the mask blocks' LIFETIME is a start-up delay for
  // every one of the ~63 forward launches of a step.
  const uint32_t trip_begin = trips_per_block ? blockIdx.z * trips_per_block : 0u;
  if (trip_begin >= trips) return;
  const uint32_t trip_end = trips_per_block ? min(trips, trip_begin + trips_per_block) : trips;
  const uint64_t off = c.philox_offset + (offset_base ? *offset_base : 0ull);
  const uint64_t ctr_base = off >> 2;                         // curand skipahead: offset counts 32-bit words
  const uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t idx = blockIdx.x * 256u + threadIdx.x;
  const uint32_t idx0 = idx & ~31u;
  const uint32_t lo_t = c.thresh_lo, hi_t = c.thresh_hi;
  // Store role, fixed before the loop so that the loop body stays branch-free: lane ii < 4 owns ballot word ii of
  // every trip, i.e. elements (4 trip + ii) * tn + idx0 .. + 31; it stores for trip < my_trips (0 for lanes >= 4).
  // (Selecting the word and bounds-checking per trip cost 40 % of the instructions of the first version.)
  uint32_t my_trips = 0u;
  const uint64_t first = static_cast<uint64_t>(lane) * tn + idx0;
  if (lane < 4u && first < c.numel) my_trips = static_cast<uint32_t>((c.numel - first - 1ull) / (4ull * tn)) + 1u;
  const uint32_t wstep = tn >> 3;                              // words per trip (4 tn bits)
  uint32_t* wp = center_bits + c.center_word_off + (first >> 5) + static_cast<size_t>(trip_begin) * wstep;
  const bool is1 = lane == 1u, is2 = lane == 2u, is3 = lane == 3u;
  // The counter is (offset/4 + trip, idx): its first word is the same for EVERY thread of the call and its second word is
  // fixed per thread, so the first two rounds split into a part that depends only on the trip (computed once per block
  // into shared memory: 3 words per trip) and a part that depends only on the thread (hoisted out of the trip loop).
  // Per trip that leaves 16 of the 20 wide multiplies -- the pipe the kernel is bound by.  Needs the low counter word
  // not to wrap inside the call (otherwise the plain loop below runs).
  __shared__ uint4 trip_u[kMaxFastTrips];
  const uint32_t ctr_lo = static_cast<uint32_t>(ctr_base), ctr_hi = static_cast<uint32_t>(ctr_base >> 32);
  const bool fast = trip_end - trip_begin <= kMaxFastTrips && ctr_lo + (trips - 1u) >= ctr_lo;   // uniform per block
  if (fast) {
    if (threadIdx.x < trip_end - trip_begin) {
      uint32_t h0, l0, h1, l1;
      mulhilo(0xD2511F53u, ctr_lo + trip_begin + threadIdx.x, h0, l0);    // round 1, counter word 0
      const uint32_t c2p = h0 ^ k1;                                       // (c3 = 0)
      mulhilo(0xCD9E8D57u, c2p, h1, l1);                                  // round 2, the trip-only product
      trip_u[threadIdx.x] = make_uint4(h1 ^ (k0 + 0x9E3779B9u), l1, l0 ^ (k1 + 0xBB67AE85u), 0u);
    }
    __syncthreads();
    uint32_t hi_i, lo_i, h0p, l0p;
    mulhilo(0xCD9E8D57u, idx, hi_i, lo_i);                               // round 1, counter word 2 = idx
    mulhilo(0xD2511F53u, hi_i ^ ctr_hi ^ k0, h0p, l0p);                  // round 2, the thread-only product
#pragma unroll 1
    for (uint32_t trip = trip_begin; trip < trip_end; ++trip) {
      const uint4 u = trip_u[trip - trip_begin];
      uint32_t r[4];
      philox_rounds<2>(u.x ^ lo_i, u.y, h0p ^ u.z, l0p, k0, k1, r);      // state after round 2, rounds 3..10 follow
      uint32_t words[4];
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) words[ii] = __ballot_sync(0xffffffffu, r[ii] < lo_t || r[ii] >= hi_t);
      uint32_t wsel = words[0];
      wsel = is1 ? words[1] : wsel;
      wsel = is2 ? words[2] : wsel;
      wsel = is3 ? words[3] : wsel;
      if (trip < my_trips) *wp = wsel;
      wp += wstep;
    }
    return;
  }
#pragma unroll 1
  for (uint32_t trip = trip_begin; trip < trip_end; ++trip) {
    const uint64_t ctr = ctr_base + trip;
    uint32_t r[4];
    philox4x32_10(static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), idx, 0u, k0, k1, r);
    uint32_t words[4];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) words[ii] = __ballot_sync(0xffffffffu, r[ii] < lo_t || r[ii] >= hi_t);
    uint32_t wsel = words[0];
    wsel = is1 ? words[1] : wsel;
    wsel = is2 ? words[2] : wsel;
    wsel = is3 ? words[3] : wsel;
    if (trip < my_trips) *wp = wsel;
    wp += wstep;
  }
}

// Dropblock2d_ichan (reference utils_modules.py:86-139) draws `torch.bernoulli(ones_like(x) * gamma)` over the FULL
// [N,C,H,W] tensor and then zeroes a border of bs/2 (:117-121).  ATen's bernoulli_tensor_cuda_kernel
// (DistributionTemplates.h:608-650) runs CUDA_tensor_apply2 with step 4 and 512-thread blocks, grid =
// ceil(numel / 2048): thread idx owns elements 4*idx .. 4*idx+3 and draws ONE curand_uniform4 after
// curand_init(seed, idx, offset), i.e. Philox counter (offset/4, 0, idx, 0); element k is 1 iff uniform_k <= p.
// The surviving (interior) centres are written into the same compact (H-bs+1) x (W-bs+1) bitmap the DropBlock2D
// path uses, so the dilate kernel is shared.  Centres are sparse (gamma ~ 0.3 %): set bits go through atomicOr
// onto a zeroed bitmap.  call.numel = n_img*c*h*w, call.thresh_lo = largest raw word whose uniform is <= p.
__global__ void __launch_bounds__(256) dropblock_centers_ichan_kernel(const b2u_dropblock_call* __restrict__ table, uint64_t seed,
                                               const unsigned long long* __restrict__ offset_base,
                                               uint32_t* __restrict__ center_bits) {
  const b2u_dropblock_call c = table[blockIdx.y];
  const uint32_t idx = blockIdx.x * 256u + threadIdx.x;
  const uint64_t e0 = static_cast<uint64_t>(idx) * 4ull;
  if (e0 >= c.numel) return;
  const uint64_t off = c.philox_offset + (offset_base ? *offset_base : 0ull);
  const uint64_t ctr = off >> 2;
  uint32_t r[4];
  philox4x32_10(static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), idx, 0u, static_cast<uint32_t>(seed),
                static_cast<uint32_t>(seed >> 32), r);
  const uint32_t t = c.thresh_lo;
  if (!(r[0] <= t || r[1] <= t || r[2] <= t || r[3] <= t)) return;
  const int ex = c.block_size >> 1;
  const int hc = c.h - c.block_size + 1, wc = c.w - c.block_size + 1;
  uint32_t* out = center_bits + c.center_word_off;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint64_t e = e0 + k;
    if (r[k] <= t && e < c.numel) {
      const uint32_t x = static_cast<uint32_t>(e % c.w);
      const uint64_t q = e / c.w;
      const uint32_t y = static_cast<uint32_t>(q % c.h);
      const uint32_t plane = static_cast<uint32_t>(q / c.h);
      if (static_cast<int>(y) >= ex && static_cast<int>(y) < c.h - ex && static_cast<int>(x) >= ex && static_cast<int>(x) < c.w - ex) {
        const uint32_t bit = (plane * hc + (y - ex)) * wc + (x - ex);
        atomicOr(out + (bit >> 5), 1u << (bit & 31u));
      }
    }
  }
}

__global__ void dropblock_centers_from_uniform_kernel(const float* __restrict__ u, uint32_t* __restrict__ bits,
                                                      long long numel, float gamma) {
  const long long nwords = (numel + 31) / 32;
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long warp_stride = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long wd = warp_global; wd < nwords; wd += warp_stride) {
    const long long i = wd * 32 + lane;
    const bool b = i < numel && u[i] < gamma;
    const uint32_t word = __ballot_sync(0xffffffffu, b);
    if (lane == 0) bits[wd] = word;
  }
}

// 32 bits of a centre row: bit k = centre[x = xa + k] (0 outside [0, wc)); row starts at bit `rowbit`
// of the flat bitmap.
// (all bit indices fit 32 bits: numel of one torch.rand call is < 2^32, checked by the host)
__device__ __forceinline__ uint32_t row_bits(const uint32_t* __restrict__ bits, uint32_t rowbit, int xa, int wc) {
  if (xa <= -32 || xa >= wc) return 0u;
  int shift_in = 0;
  if (xa < 0) {
    shift_in = -xa;
    xa = 0;
  }
  const uint32_t q = rowbit + static_cast<uint32_t>(xa);
  const uint32_t wi = q >> 5;
  const uint32_t sh = q & 31u;
  uint32_t v = __funnelshift_r(__ldg(bits + wi), __ldg(bits + wi + 1), sh);
  const int cnt = wc - xa;                                  // valid bits from xa
  if (cnt < 32) v &= (1u << cnt) - 1u;
  return shift_in ? (v << shift_in) : v;
}

// One warp: 32 channels (lane = channel) x 32 pixels of one image row band.
// Flat 1-D grid over all calls of the table: call k owns blocks [dilate_first_block_k, dilate_first_block_{k+1}),
// = planes(k) x ceil(items(k) / 4) blocks (b2u_dropblock_plan); a block finds its call by binary search.  (A 3-D grid
// sized for the largest call in every dimension launched 28x more blocks than there is work for the U-Net's sites.)
// BS_CT > 0: block size known at compile time (the reference default 7): the smear, the ring and the slot
// arithmetic unroll into straight-line code.
constexpr int kDilateBandRows = 37;
constexpr int kDilateWarps = 4;

static inline int dilate_items(const b2u_dropblock_call& c) {
  return ((c.h + kDilateBandRows - 1) / kDilateBandRows) * ((c.w + 31) / 32);
}
static inline int dilate_blocks(const b2u_dropblock_call& c) {
  return ((dilate_items(c) + kDilateWarps - 1) / kDilateWarps) * ((c.c / 32) * c.n_img);
}

template <int RING, int BS_CT>
__device__ __forceinline__ void dilate_block(const b2u_dropblock_call& c, int local, const uint32_t* __restrict__ center_bits,
                                             uint32_t* __restrict__ mask_bits, unsigned long long* __restrict__ keep_counts) {
  constexpr int band_rows = kDilateBandRows;
  const int cgs = c.c >> 5;
  const int wwords = (c.w + 31) >> 5;
  const int bands = (c.h + band_rows - 1) / band_rows;
  const int nitems = bands * wwords;
  const int blocks_x = (nitems + kDilateWarps - 1) / kDilateWarps;
  const int plane_grp = local / blocks_x;                               // (img, channel group)
  const int warp_in_block = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int item = (local - plane_grp * blocks_x) * kDilateWarps + warp_in_block;       // (band, wword)
  if (item >= nitems || plane_grp >= cgs * c.n_img) return;             // uniform per warp
  const int band = item / wwords, wj = item - band * wwords;
  const int img = plane_grp / cgs, cg = plane_grp - img * cgs;
  const int bs = BS_CT > 0 ? BS_CT : c.block_size;
  const int hc = c.h - bs + 1, wc = c.w - bs + 1;
  const int ch = cg * 32 + lane;
  const uint32_t* cb = center_bits + c.center_word_off;
  const uint32_t plane_bit = (static_cast<uint32_t>(img) * c.c + ch) * static_cast<uint32_t>(hc) * wc;
  const int w0 = wj * 32;
  const int h_begin = band * band_rows;
  const int h_end = min(h_begin + band_rows, c.h);
  const uint32_t valid_w = (c.w - w0 >= 32) ? 0xffffffffu : ((1u << (c.w - w0)) - 1u);
  uint32_t* mout = mask_bits + c.mask_word_off;

  // smeared centre rows in a register ring of bs <= RING entries
  uint32_t ring[RING];
#pragma unroll
  for (int i = 0; i < RING; ++i) ring[i] = 0u;
  unsigned long long keep = 0;
  // rows h-bs+1 .. h contribute to output row h; prime the window with rows h_begin-bs+1 .. h_begin-1.
  // The two 32-bit windows of row r+1 are fetched before row r is processed (software prefetch).
  auto fetch = [&](int r, uint32_t& lo, uint32_t& hi) {
    lo = hi = 0u;
    if (r >= 0 && r < hc) {
      const uint32_t rowbit = plane_bit + static_cast<uint32_t>(r) * wc;
      lo = row_bits(cb, rowbit, w0 - 32, wc);
      hi = row_bits(cb, rowbit, w0, wc);
    }
  };
  uint32_t nlo, nhi;
  fetch(h_begin - bs + 1, nlo, nhi);
  int slot = (((h_begin - bs + 1) % bs) + bs) % bs;          // ring slot = r mod bs, advanced incrementally
  uint32_t* orow = mout + ((static_cast<size_t>(img) * c.h + h_begin) * c.w + (w0 + lane)) * cgs + cg;
  const size_t orow_stride = static_cast<size_t>(c.w) * cgs;
  for (int r = h_begin - bs + 1; r < h_end; ++r) {
    const uint32_t lo = nlo, hi = nhi;
    if (r + 1 < h_end) fetch(r + 1, nlo, nhi);
    uint32_t sm = 0u;
    {
      uint64_t v = (static_cast<uint64_t>(hi) << 32) | lo;
      // OR of shifts 0..bs-1 by doubling
      int have = 1;
      while (have < bs) {
        const int add = min(have, bs - have);
        v |= v << add;
        have += add;
      }
      sm = static_cast<uint32_t>(v >> 32);
    }
#pragma unroll
    for (int i = 0; i < RING; ++i)
      if (i == slot) ring[i] = sm;
    if (++slot == bs) slot = 0;
    if (r < h_begin) continue;
    uint32_t drop = 0u;
#pragma unroll
    for (int i = 0; i < RING; ++i) drop |= ring[i];
    const uint32_t keepw = ~drop & valid_w;
    keep += __popc(keepw);
    // 32x32 bit transpose (lane = channel -> lane = pixel) with 5 butterfly exchanges instead of 32 ballots:
    // at distance j the off-diagonal j x j blocks of every 2j x 2j block swap between lanes i and i^j.
    uint32_t mine = keepw;
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
      const uint32_t m = j == 16 ? 0x0000FFFFu : j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
      const uint32_t y = __shfl_xor_sync(0xffffffffu, mine, j);
      mine = (lane & j) ? ((mine & ~m) | ((y & ~m) >> j)) : ((mine & m) | ((y & m) << j));
    }
    if (w0 + lane < c.w) *orow = mine;
    orow += orow_stride;
  }
  // integer count: atomics are order-independent, the result is deterministic
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
  if (lane == 0 && keep) atomicAdd(keep_counts + c.count_index, keep);
}

// One short block per entry of the flat block list (see the centre kernel for why not a persistent grid).
template <int RING, int BS_CT>
__global__ void __launch_bounds__(128) dropblock_dilate_kernel(const b2u_dropblock_call* __restrict__ table, int n_calls,
                                        int total_blocks, const uint32_t* __restrict__ center_bits,
                                        uint32_t* __restrict__ mask_bits, unsigned long long* __restrict__ keep_counts) {
  pdl_wait();       // predecessor grid complete + visible (no-op without the PDL launch attribute)
  pdl_trigger();    // the successor may be scheduled once every CTA got here
  int k = 0, hi_k = n_calls - 1;
  while (k < hi_k) {                                          // last call whose first block is <= blockIdx.x
    const int mid = (k + hi_k + 1) >> 1;
    if (__ldg(&table[mid].dilate_first_block) <= static_cast<int>(blockIdx.x)) k = mid;
    else hi_k = mid - 1;
  }
  int next_first = k + 1 < n_calls ? __ldg(&table[k + 1].dilate_first_block) : total_blocks;
  for (int b = blockIdx.x; b < total_blocks; b += gridDim.x) {
    while (b >= next_first) {
      ++k;
      next_first = k + 1 < n_calls ? __ldg(&table[k + 1].dilate_first_block) : total_blocks;
    }
    const b2u_dropblock_call c = table[k];
    dilate_block<RING, BS_CT>(c, b - c.dilate_first_block, center_bits, mask_bits, keep_counts);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// v2 of the dilation (round 2): no 32x32 bit transpose.  Only gamma ~ 0.3 % of the draws are centres, so
//   1. scatter : every set bit of the compact NCHW centre bitmap is OR-ed (atomicOr: order-independent) into a zeroed
//                NHWC WORD bitmap P[img][y + bs/2][x + bs/2][c / 32] (bit c % 32) -- the zero-padded centre map of
//                utils_modules.py:51-53 in the layout the keep mask is wanted in;
//   2. dilate  : keep = ~(bs x bs box OR of P) is then word-parallel over 32 channels: a thread owns one (pixel, channel
//                word) column of a 32-row band, ORs its bs horizontal neighbours per row and keeps the last bs row
//                results in a register ring (unrolled by bs, so the ring is static).
// ~35 instructions per output word instead of ~100; the flat call table is reused (`reserved` = first block of a call in
// the v2 grid).
constexpr int kD2Threads = 128;
constexpr int kD2Band = 32;

static inline int dilate2_blocks(const b2u_dropblock_call& c) {
  const int cgs = c.c / 32;
  return c.n_img * ((c.h + kD2Band - 1) / kD2Band) * ((c.w * cgs + kD2Threads - 1) / kD2Threads);
}

__global__ void __launch_bounds__(256) dropblock_scatter_kernel(const b2u_dropblock_call* __restrict__ table,
                                                                const uint32_t* __restrict__ center_bits, uint32_t* __restrict__ pbits) {
  pdl_wait();
  pdl_trigger();
  const b2u_dropblock_call c = table[blockIdx.y];
  const int bs = c.block_size, ex = bs >> 1;
  const uint32_t hc = c.h - bs + 1, wc = c.w - bs + 1;
  const uint32_t plane = hc * wc;
  const uint32_t nbits = static_cast<uint32_t>(c.n_img) * c.c * plane;     // < 2^32 (checked by the host)
  const uint32_t nchunks = (nbits + 127u) >> 7;                            // 128-bit chunks (the bitmap is padded)
  if (blockIdx.x * 256u >= nchunks) return;                                // the grid is sized for the largest call
  // A lane scans 128 consecutive bits per trip (one LDG.128): ~0.4 centres per chunk, so the decode below runs ~2.5
  // times per warp trip of 4096 bits instead of ~6 times with one word per lane.  The decode locates the chunk's first
  // bit with ONE fp64 reciprocal division (exact after a one-step correction) and walks from there; row / channel
  // splits use 32-bit multiply-high reciprocals (dividends < 2^20 and < 2^16: floor(2^32 / d) + 1 is exact there).
  // (the three reciprocals are computed by ONE thread per block: as per-thread prologue they were 80 % of the kernel's
  // instructions -- 570 k warps x ~300 instructions of software division)
  __shared__ double sh_inv;
  __shared__ uint32_t sh_mw, sh_mc;
  if (threadIdx.x == 0) {
    sh_inv = 1.0 / static_cast<double>(plane);
    sh_mw = wc > 1 ? 0xFFFFFFFFu / wc + 1u : 0u;               // floor(2^32 / d) + 1 for d >= 2
    sh_mc = c.c > 1 ? 0xFFFFFFFFu / static_cast<uint32_t>(c.c) + 1u : 0u;
  }
  __syncthreads();
  const double inv_plane = sh_inv;
  const uint32_t mw = sh_mw, mc = sh_mc;
  const uint4* cb = reinterpret_cast<const uint4*>(center_bits + c.center_word_off);
  uint32_t* pb = pbits + c.mask_word_off;
  const uint32_t cgs = c.c >> 5;
  const uint32_t prow = static_cast<uint32_t>(c.w) * cgs, pimg = static_cast<uint32_t>(c.h) * prow;   // < 2^31 words per call (host)
  for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < nchunks; i += gridDim.x * 256u) {
    const uint4 q = __ldg(cb + i);
    uint32_t wds[4] = {q.x, q.y, q.z, q.w};
    if ((q.x | q.y | q.z | q.w) == 0u) continue;
    // position of the chunk's first bit: plane (= img * C + ch), row, column -- one fp64 reciprocal division (exact after
    // a one-step correction) and two multiply-high divisions per NON-EMPTY chunk; the set bits then walk from there
    // with compare-and-wrap only (a chunk spans 128 positions: a handful of row wraps, at most a few plane wraps)
    const uint32_t e0 = i << 7;
    uint32_t pl0 = __double2uint_rz(static_cast<double>(e0) * inv_plane);
    uint32_t rem0 = e0 - pl0 * plane;
    if (static_cast<int32_t>(rem0) < 0) { --pl0; rem0 += plane; }
    else if (rem0 >= plane) { ++pl0; rem0 -= plane; }
    const uint32_t y0 = mw ? __umulhi(rem0, mw) : rem0, x0 = rem0 - y0 * wc;
    const uint32_t img0 = mc ? __umulhi(pl0, mc) : pl0, ch0 = pl0 - img0 * c.c;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t wd = wds[k];
      while (wd) {
        const uint32_t b = __ffs(wd) - 1;
        wd &= wd - 1u;
        const uint32_t d = 32u * k + b;
        if (e0 + d >= nbits) break;                                        // padding bits of the last chunk
        uint32_t x = x0 + d, y = y0, ch = ch0, img = img0;
        while (x >= wc) { x -= wc; ++y; }
        while (y >= hc) { y -= hc; if (++ch == static_cast<uint32_t>(c.c)) { ch = 0; ++img; } }
        atomicOr(pb + (img * pimg + (y + ex) * prow + (x + ex) * cgs + (ch >> 5)), 1u << (ch & 31u));
      }
    }
  }
}

// One (image, 32-row band, 128-word strip) per block; CGS = channel words per pixel when it is one of the U-Net's
// (compile-time neighbour offsets: the seven loads of a row need no address arithmetic), 0 = generic.
template <int BS, int CGS>
__device__ __forceinline__ void dilate_nhwc_block(const b2u_dropblock_call& c, int local, const uint32_t* __restrict__ pbits,
                                                  uint32_t* __restrict__ mask_bits, unsigned long long* __restrict__ keep_counts) {
  constexpr int EX = BS / 2;
  const int cgs = CGS > 0 ? CGS : (c.c >> 5);
  const int roww = c.w * cgs;                                 // words per image row
  const int strips = (roww + kD2Threads - 1) / kD2Threads;
  const int bands = (c.h + kD2Band - 1) / kD2Band;
  const int strip = local % strips;
  local /= strips;
  const int band = local % bands;
  const int img = local / bands;
  const int j = strip * kD2Threads + threadIdx.x;             // word column inside the row: (pixel, channel word)
  const bool col_ok = j < roww;
  const int px = col_ok ? j / cgs : 0;
  const int h_begin = band * kD2Band;
  const int h_end = min(h_begin + kD2Band, c.h);
  const size_t base = c.mask_word_off + static_cast<size_t>(img) * c.h * roww + j;
  const uint32_t* pcol = pbits + base;                        // this column, row 0
  uint32_t* mcol = mask_bits + base;
  // interior columns (all BS horizontal neighbours exist) take the unpredicated path; a warp holds 32 / cgs pixels, so
  // only the warps at the two ends of a row diverge
  const bool interior = col_ok && px >= EX && px + EX < c.w;
  bool nb_ok[BS];
#pragma unroll
  for (int t = 0; t < BS; ++t) nb_ok[t] = col_ok && px + t - EX >= 0 && px + t - EX < c.w;
  // running pointers instead of r * roww products: the input pointer walks rows h_begin - EX .. h_end + EX - 1 (rows
  // outside the image contribute zeros), the output pointer rows h_begin .. h_end - 1
  const uint32_t* in_row = pcol + (static_cast<long>(h_begin) - EX) * roww;
  uint32_t* out_row = mcol + static_cast<size_t>(h_begin) * roww;
  int rin = h_begin - EX;
  auto hsmear_next = [&]() -> uint32_t {
    uint32_t v = 0u;
    if (rin >= 0 && rin < c.h) {                              // uniform per block
      if (interior) {
        uint32_t q[BS];
#pragma unroll
        for (int t = 0; t < BS; ++t) q[t] = __ldg(in_row + (t - EX) * cgs);
#pragma unroll
        for (int t = 0; t < BS; ++t) v |= q[t];
      } else {
#pragma unroll
        for (int t = 0; t < BS; ++t)
          if (nb_ok[t]) v |= __ldg(in_row + (t - EX) * cgs);
      }
    }
    in_row += roww;
    ++rin;
    return v;
  };
  uint32_t ring[BS];
#pragma unroll
  for (int t = 0; t < BS - 1; ++t) ring[t] = hsmear_next();                   // rows h_begin - EX .. h_begin + EX - 1
  unsigned int keep = 0;
  // output row r needs rows r - EX .. r + EX; the loop is unrolled by BS so that the ring slot is a compile-time index
  for (int r0 = h_begin; r0 < h_end; r0 += BS) {
#pragma unroll
    for (int u = 0; u < BS; ++u) {
      if (r0 + u < h_end) {
        ring[(u + BS - 1) % BS] = hsmear_next();
        uint32_t drop = 0u;
#pragma unroll
        for (int t = 0; t < BS; ++t) drop |= ring[t];
        if (col_ok) {
          const uint32_t kw = ~drop;
          keep += __popc(kw);
          *out_row = kw;
        }
        out_row += roww;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
  __shared__ unsigned int wsum[kD2Threads / 32];
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = keep;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long tot = 0;
    for (int wv = 0; wv < kD2Threads / 32; ++wv) tot += wsum[wv];
    if (tot) atomicAdd(keep_counts + c.count_index, tot);       // integer atomics: order-independent, deterministic
  }
}

template <int BS>
__global__ void __launch_bounds__(kD2Threads) dropblock_dilate_nhwc_kernel(const b2u_dropblock_call* __restrict__ table, int n_calls,
                                                                         int total_blocks, const uint32_t* __restrict__ pbits,
                                                                         uint32_t* __restrict__ mask_bits,
                                                                         unsigned long long* __restrict__ keep_counts) {
  pdl_wait();
  pdl_trigger();
  int k = 0, hi_k = n_calls - 1;
  while (k < hi_k) {                                          // last call whose first v2 block is <= blockIdx.x
    const int mid = (k + hi_k + 1) >> 1;
    if (__ldg(&table[mid].reserved) <= static_cast<int>(blockIdx.x)) k = mid;
    else hi_k = mid - 1;
  }
  const b2u_dropblock_call c = table[k];
  const int local = static_cast<int>(blockIdx.x) - c.reserved;
  switch (c.c >> 5) {                                         // uniform per block
    case 2: dilate_nhwc_block<BS, 2>(c, local, pbits, mask_bits, keep_counts); break;
    case 4: dilate_nhwc_block<BS, 4>(c, local, pbits, mask_bits, keep_counts); break;
    case 8: dilate_nhwc_block<BS, 8>(c, local, pbits, mask_bits, keep_counts); break;
    case 16: dilate_nhwc_block<BS, 16>(c, local, pbits, mask_bits, keep_counts); break;
    case 32: dilate_nhwc_block<BS, 32>(c, local, pbits, mask_bits, keep_counts); break;
    default: dilate_nhwc_block<BS, 0>(c, local, pbits, mask_bits, keep_counts); break;
  }
}

}  // namespace b2u

using namespace b2u;

static int centers_launch(const b2u_dropblock_call* table, int n_calls, uint64_t seed, const unsigned long long* offset_base,
                          uint32_t* center_bits, int max_trips, void* stream) {
  B2U_REQUIRE(table && center_bits && n_calls > 0, "bad arguments");
  int sms = 0, mt = 0;
  int rc = b2u_device_info(&sms, &mt);
  if (rc) return rc;
  // torch's grid.x never exceeds SMs * (maxThreadsPerSM / 256); calls with a smaller grid exit early.  With max_trips > 0
  // (the largest trip count of the table, from the host copy) grid.z splits the trips into runs of 4: short blocks.
  static int env_tpb = -1;
  if (env_tpb < 0) {
    const char* e = getenv("B2U_CENTERS_TPB");
    env_tpb = e ? atoi(e) : 12;
    if (env_tpb < 1 || env_tpb > 64) env_tpb = 12;
  }
  // (4 trips per block DOUBLED the kernel's instructions: the per-thread prologue -- call table, thread-only Philox
  // products, store role -- is ~150 instructions against 60 per trip)
  const unsigned tpb = max_trips > 0 ? static_cast<unsigned>(env_tpb) : 0u;
  dim3 grid(sms * (mt / 256), n_calls, tpb ? (static_cast<unsigned>(max_trips) + tpb - 1u) / tpb : 1u);
  B2U_PDL_LAUNCH((dropblock_centers_kernel), grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), table, seed, offset_base, center_bits, tpb);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_dropblock_centers(const b2u_dropblock_call* table, int n_calls, uint64_t seed,
                                     const unsigned long long* offset_base, uint32_t* center_bits, void* stream) {
  return centers_launch(table, n_calls, seed, offset_base, center_bits, 0, stream);
}

extern "C" int b2u_dropblock_centers_ex(const b2u_dropblock_call* table, int n_calls, const b2u_dropblock_call* host_table,
                                        uint64_t seed, const unsigned long long* offset_base, uint32_t* center_bits, void* stream) {
  B2U_REQUIRE(host_table, "bad arguments");
  int max_trips = 1;
  for (int i = 0; i < n_calls; ++i) {
    const b2u_dropblock_call& c = host_table[i];
    B2U_REQUIRE(c.grid > 0 && c.numel > 0, "call %d: empty grid", i);
    const unsigned long long per_trip = static_cast<unsigned long long>(c.grid) * 1024ull;
    const int t = static_cast<int>((c.numel - 1ull) / per_trip) + 1;
    if (t > max_trips) max_trips = t;
  }
  return centers_launch(table, n_calls, seed, offset_base, center_bits, max_trips, stream);
}

extern "C" int b2u_dropblock_centers_ichan(const b2u_dropblock_call* table, int n_calls, const b2u_dropblock_call* host_table,
                                           uint64_t seed, const unsigned long long* offset_base, uint32_t* center_bits,
                                           long long center_words_total, void* stream) {
  B2U_REQUIRE(table && host_table && center_bits && n_calls > 0 && center_words_total > 0, "bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint32_t max_numel = 0;
  for (int i = 0; i < n_calls; ++i) {
    B2U_REQUIRE(host_table[i].block_size >= 1 && (host_table[i].block_size & 1), "block_size must be odd");
    if (host_table[i].numel > max_numel) max_numel = host_table[i].numel;
  }
  B2U_CHECK_CUDA(cudaMemsetAsync(center_bits, 0, static_cast<size_t>(center_words_total) * 4, st));
  const unsigned threads_needed = (max_numel + 3u) / 4u;
  dim3 grid((threads_needed + 255u) / 256u, n_calls);
  dropblock_centers_ichan_kernel<<<grid, 256, 0, st>>>(table, seed, offset_base, center_bits);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_dropblock_plan(b2u_dropblock_call* host_table, int n_calls, long long* total_blocks) {
  B2U_REQUIRE(host_table && n_calls > 0, "bad arguments");
  long long total = 0;
  for (int i = 0; i < n_calls; ++i) {
    b2u_dropblock_call& c = host_table[i];
    B2U_REQUIRE(c.block_size >= 1 && c.block_size <= 31 && (c.block_size & 1), "block_size must be odd and <= 31 (got %d)", c.block_size);
    B2U_REQUIRE(c.c % 32 == 0 && c.c > 0, "channels must be a multiple of 32 (got %d)", c.c);
    B2U_REQUIRE(c.h >= c.block_size && c.w >= c.block_size, "feature map %dx%d smaller than block_size %d", c.h, c.w, c.block_size);
    B2U_REQUIRE(total < 0x7fffffffll, "too many dilate blocks");
    c.dilate_first_block = static_cast<int32_t>(total);
    total += dilate_blocks(c);
  }
  long long total2 = 0;                                       // v2 (NHWC) dilation grid: `reserved` = first block of the call
  for (int i = 0; i < n_calls; ++i) {
    host_table[i].reserved = static_cast<int32_t>(total2);
    total2 += dilate2_blocks(host_table[i]);
    B2U_REQUIRE(total2 < 0x7fffffffll, "too many dilate blocks");
  }
  B2U_REQUIRE(total < 0x7fffffffll, "too many dilate blocks");
  if (total_blocks) *total_blocks = total;
  return B2U_OK;
}

extern "C" int b2u_dropblock_dilate(const b2u_dropblock_call* table, int n_calls, const b2u_dropblock_call* host_table,
                                    const uint32_t* center_bits, uint32_t* mask_bits, unsigned long long* keep_counts,
                                    void* stream) {
  B2U_REQUIRE(table && host_table && center_bits && mask_bits && keep_counts && n_calls > 0, "bad arguments");
  long long total = 0;
  int max_bs = 0;
  bool all7 = true;
  for (int i = 0; i < n_calls; ++i) {
    const b2u_dropblock_call& c = host_table[i];
    B2U_REQUIRE(c.block_size >= 1 && c.block_size <= 31 && (c.block_size & 1), "block_size must be odd and <= 31 (got %d)", c.block_size);
    B2U_REQUIRE(c.c % 32 == 0 && c.c > 0, "channels must be a multiple of 32 (got %d)", c.c);
    B2U_REQUIRE(c.dilate_first_block == total, "call table was not planned (b2u_dropblock_plan) or changed shape afterwards");
    total += dilate_blocks(c);
    if (c.block_size > max_bs) max_bs = c.block_size;
    all7 = all7 && c.block_size == 7;
  }
  dim3 grid(static_cast<unsigned>(total));
  const int tb = static_cast<int>(total);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (all7)
    B2U_PDL_LAUNCH((dropblock_dilate_kernel<7, 7>), grid, 128, 0, st, table, n_calls, tb, center_bits, mask_bits, keep_counts);
  else if (max_bs <= 7)
    B2U_PDL_LAUNCH((dropblock_dilate_kernel<7, 0>), grid, 128, 0, st, table, n_calls, tb, center_bits, mask_bits, keep_counts);
  else
    B2U_PDL_LAUNCH((dropblock_dilate_kernel<31, 0>), grid, 128, 0, st, table, n_calls, tb, center_bits, mask_bits, keep_counts);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_dropblock_centers_from_uniform(const float* u, uint32_t* center_bits, long long numel, float gamma,
                                                  void* stream) {
  B2U_REQUIRE(u && center_bits && numel > 0, "bad arguments");
  long long warps = (numel + 31) / 32;
  long long blocks = (warps + 7) / 8;
  if (blocks > b2u_num_sms() * 8) blocks = b2u_num_sms() * 8;
  dropblock_centers_from_uniform_kernel<<<static_cast<int>(blocks), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(u, center_bits, numel, gamma);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_dropblock_dilate_v2(const b2u_dropblock_call* table, int n_calls, const b2u_dropblock_call* host_table,
                                       const uint32_t* center_bits, uint32_t* scatter_bits, long long mask_words_total,
                                       uint32_t* mask_bits, unsigned long long* keep_counts, void* stream) {
  B2U_REQUIRE(table && host_table && center_bits && scatter_bits && mask_bits && keep_counts && n_calls > 0 && mask_words_total > 0,
              "bad arguments");
  long long total = 0;
  uint32_t max_words = 0;
  for (int i = 0; i < n_calls; ++i) {
    const b2u_dropblock_call& c = host_table[i];
    B2U_REQUIRE(c.block_size == 7, "b2u_dropblock_dilate_v2 is built for block_size 7 (got %d): use b2u_dropblock_dilate", c.block_size);
    B2U_REQUIRE(c.c % 32 == 0 && c.c > 0, "channels must be a multiple of 32 (got %d)", c.c);
    B2U_REQUIRE(c.reserved == total, "call table was not planned (b2u_dropblock_plan) or changed shape afterwards");
    total += dilate2_blocks(c);
    const double nbits = static_cast<double>(c.n_img) * c.c * (c.h - c.block_size + 1) * (c.w - c.block_size + 1);
    B2U_REQUIRE(nbits < 4294967295.0, "one DropBlock call is limited to 2^32 - 1 centre positions");
    const uint32_t words = static_cast<uint32_t>((static_cast<unsigned long long>(nbits) + 31ull) >> 5);
    if (words > max_words) max_words = words;
    B2U_REQUIRE(static_cast<double>(c.n_img) * c.h * c.w * (c.c / 32) < 2147483648.0, "one DropBlock call is limited to 2^31 mask words");
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B2U_CHECK_CUDA(cudaMemsetAsync(scatter_bits, 0, static_cast<size_t>(mask_words_total) * 4, st));
  const unsigned max_chunks = (max_words + 3u) / 4u;
  unsigned bx = (max_chunks + 256u * 16u - 1u) / (256u * 16u);          // ~16 chunks of 128 bits per thread (grid-stride)
  if (bx < 1u) bx = 1u;
  dim3 sgrid(bx, n_calls);
  dropblock_scatter_kernel<<<sgrid, 256, 0, st>>>(table, center_bits, scatter_bits);
  B2U_LAUNCH_CHECK();
  B2U_PDL_LAUNCH((dropblock_dilate_nhwc_kernel<7>), dim3(static_cast<unsigned>(total)), kD2Threads, 0, st, table, n_calls,
                 static_cast<int>(total), scatter_bits, mask_bits, keep_counts);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}
