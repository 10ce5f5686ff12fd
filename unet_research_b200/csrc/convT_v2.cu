// convT2x2 (stride 2, no bias) v2: persistent tcgen05 GEMM  [M = N*H*W input pixels] x [K = Cin] x [N = 4*Cout].
//
// v1 (conv_gemm.cu, mode 1) launches one CTA per (128-pixel tile, tap, 64..256 output channels): at the shallow
// decoder levels a CTA issues 8 MMAs and spends most of its life in set-up (TMEM allocation, barrier init, pipeline
// fill) -- 13 320 CTAs and 6 % tensor-pipe activity for the 128->64 layer (profiles/r01_step_metrics.md).  v2:
//
//   * the packed weight [tap][Cout][Cin] is read as ONE K-major matrix of 4*Cout rows, so an N tile of 256 columns
//     spans all four taps when Cout = 64 (two when Cout = 128): the activation tile is fetched once per N tile
//     instead of once per tap;
//   * persistent CTAs (grid = SM count) walk the (pixel tile, N tile) items, N-tile-major so concurrently running
//     CTAs share weight tiles in L2; a 4-stage TMA ring runs ahead across item boundaries;
//   * TMEM holds two 256-column accumulators: the epilogue of item j (pixel-shuffle store into the 2H x 2W output
//     + GroupNorm partial sums) overlaps the MMAs of item j+1.  The shallow layers are HBM-write bound, so the
//     epilogue, not the tensor pipe, sets the pace.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..9 = epilogue: two
// warps per TMEM lane quarter, each draining four of the eight 32-column chunks of an accumulator.
//
// TMA-STORE EPILOGUE (kTmaStore, 16-bit formats).  A thread owns one input pixel, i.e. one 128-byte run of the
// pixel-shuffled output per (tap, 64 channels), and neighbouring threads own runs 256 bytes apart: a warp-wide
// STG.128 touched 32 different lines with 16 bytes each (the two shallow levels ran at 0.39-0.43 of their HBM-write
// floor, profiles/r01e_step_metrics.md).  Now every epilogue warp stages its 32 pixels x 64 channels (4 KB, 128-byte
// swizzle: conflict-free STS.128) in its own double-buffered slot and one lane issues a 5-D TMA store whose box
// {64 ch, 8 w, 1 dy, 4 h, 1 n} over the output viewed as [n][h][dy][w][dx * cout + c] is the pixel shuffle; the
// tensor bounds clip ragged tiles.  Only __syncwarp() is needed: no warp ever touches another warp's slot.
#include "b2u_common.cuh"
#include "conv_host.cuh"

#include <stdlib.h>

namespace b2u {

struct ConvTParams {
  int n, h, w;                // input pixel grid
  int tiles_w, tiles_h;       // 8 x 16 pixel tiles per image
  int total_tiles;            // n * tiles_w * tiles_h
  int num_items;              // total_tiles * n_tiles
  int kc_chunks;              // Cin / (128 B of channels)
  int cout;
  int stages;
  int sgs_log2;               // statistics sub-group size (log2), -1 = none
  void* y;                    // [n, 2h, 2w, cout]
  int stage_off;              // byte offset of the TMA-store staging slots inside the aligned shared memory (kTmaStore)
  float* partials;            // [n][tiles_per_image * 4][cout / sgs][2]
};

constexpr int kTBlockN = 256;
constexpr int kTABytes = 128 * 128;
constexpr int kTBBytes = kTBlockN * 128;
constexpr int kTThreads = 320;            // producer + MMA + 8 epilogue warps

template <int NV>
__device__ __forceinline__ void convT_epilogue_stats(const float (&x)[32], bool valid, int lane, float* scratch) {
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

template <int kFmt, bool kTmaStore>
__global__ void __launch_bounds__(kTThreads, 1)
convT_v2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmY,
                const ConvTParams p) {
  constexpr bool kTf32 = kFmt == 1;
  static_assert(!(kTmaStore && kTf32), "the TMA-store epilogue stages 16-bit rows");
  using OutT = typename FmtTraits<kFmt>::T;
  constexpr int kKElems = kTf32 ? 32 : 64;
  constexpr int kUmmaK = kTf32 ? 8 : 16;
  constexpr uint32_t kIdesc = umma_idesc(128, kTBlockN, FmtTraits<kFmt>::kIdescFmt);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.stages;
  uint8_t* smA = smem;                                         // [S][16 KB]
  uint8_t* smB = smem + S * kTABytes;                          // [S][32 KB]
  uint64_t* full = reinterpret_cast<uint64_t*>(smB + S * kTBBytes);
  uint64_t* empty = full + S;
  uint64_t* t_full = empty + S;                                // [2]
  uint64_t* t_empty = t_full + 2;                              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  float* stat_scratch = reinterpret_cast<float*>(tmem_slot + 2);   // [2 epilogue groups][4 lane quarters][256]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_per_image = p.tiles_w * p.tiles_h;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (kTmaStore) tma_prefetch_desc(&tmY);
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 4); }
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
  // PDL: barrier init, TMEM allocation and descriptor prefetch above overlapped the predecessor's tail; from here on the
  // kernel touches tensors the predecessor wrote (or still reads)
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        if (item + static_cast<int>(gridDim.x) >= p.num_items) pdl_trigger();   // last item of this CTA: let the successor launch
        const int nt = item / p.total_tiles;
        const int tile = item - nt * p.total_tiles;
        const int img = tile / tiles_per_image;
        const int r = tile - img * tiles_per_image;
        const int ty = r / p.tiles_w, tx = r - ty * p.tiles_w;
        for (int kc = 0; kc < p.kc_chunks; ++kc) {
          mbar_wait(&empty[s], ph);
          mbar_arrive_expect_tx(&full[s], kTABytes + kTBBytes);
          tma_load_4d(smA + s * kTABytes, &tmA, &full[s], kc * kKElems, tx * 8, ty * 16, img);
          tma_load_2d(smB + s * kTBBytes, &tmB, &full[s], kc * kKElems, nt * kTBlockN);
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // the whole warp runs the loop (converged control flow); one elected lane issues (see umma_ss_conv)
    {
      const uint32_t leader = elect_one() ? 1u : 0u;
      const uint32_t tmem_u = warp_uniform(tmem_base);
      int s = 0, buf = 0;
      uint32_t ph = 0, pt = 1;
      const uint64_t adesc0 = umma_desc_k_sw128(smem_u32(smA));
      const uint64_t bdesc0 = umma_desc_k_sw128(smem_u32(smB));
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        mbar_wait(&t_empty[buf], pt);
        tc_fence_after();
        const uint32_t acc = tmem_u + buf * kTBlockN;
        for (int kc = 0; kc < p.kc_chunks; ++kc) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint64_t adesc = adesc0 + static_cast<uint64_t>((s * kTABytes) >> 4);
          const uint64_t bdesc = bdesc0 + static_cast<uint64_t>((s * kTBBytes) >> 4);
          static_assert(kKElems / kUmmaK == 4, "one swizzle row = four K steps");
          umma_ss_conv4<kTf32>(acc, adesc, bdesc, kIdesc, kc != 0 ? 1u : 0u, leader);
          umma_commit_conv(&empty[s], leader);
          if (++s == S) { s = 0; ph ^= 1; }
        }
        umma_commit_conv(&t_full[buf], leader);
        if (++buf == 2) { buf = 0; pt ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // Two independent groups of four warps (one warp per TMEM lane quarter): group g owns accumulator buffer g, i.e.
    // every other item of this CTA, and all 256 of its columns.  The groups never synchronise with each other, so
    // while one sits in its statistics barrier or waits for TMEM the other keeps the store pipe busy -- with all
    // eight warps on the same item the epilogue was one latency chain of ~10k cycles per item next to 512 cycles
    // of MMA (shallow levels: 2.4 TB/s of the HBM-write-bound 6.4).
    const int q = warp & 3;                                      // TMEM lane quarter this warp may access
    const int grp = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int dh = row >> 3, dw = row & 7;
    float* grp_scratch = stat_scratch + grp * 1024;              // [4 lane quarters][256]
    float* my_scratch = grp_scratch + q * 256;
    const int et = threadIdx.x - 64 - grp * 128;                 // 0..127 within the group
    const int out_w = 2 * p.w;
    const int buf = grp;
    // this warp's two 4 KB staging slots (1024-byte aligned: the 128-byte swizzle is a function of the address)
    const uint32_t my_stage = smem_u32(smem + p.stage_off) + static_cast<uint32_t>((warp - 2) * 8192 + lane * 128);
    uint32_t pf = 0;
    int li = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++li) {
      if ((li & 1) != grp) continue;
      const int nt = item / p.total_tiles;
      const int tile = item - nt * p.total_tiles;
      const int img = tile / tiles_per_image;
      const int r = tile - img * tiles_per_image;
      const int ty = r / p.tiles_w, tx = r - ty * p.tiles_w;
      const int hh = ty * 16 + dh, ww = tx * 8 + dw;
      const bool valid = hh < p.h && ww < p.w;
      mbar_wait(&t_full[buf], pf);
      tc_fence_after();
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kTBlockN;
      auto process = [&](uint32_t (&rr)[32], int chunk) {
        float x[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(rr[i]);
        if (p.sgs_log2 >= 0) {
          switch (p.sgs_log2) {
            case 1: convT_epilogue_stats<32>(x, valid, lane, my_scratch + chunk * 32); break;
            case 2: convT_epilogue_stats<16>(x, valid, lane, my_scratch + chunk * 16); break;
            case 3: convT_epilogue_stats<8>(x, valid, lane, my_scratch + chunk * 8); break;
            case 4: convT_epilogue_stats<4>(x, valid, lane, my_scratch + chunk * 4); break;
            default: convT_epilogue_stats<2>(x, valid, lane, my_scratch + chunk * 2); break;
          }
        }
        if constexpr (kTmaStore) {
          // chunks (2k, 2k+1) are the two 64-byte halves of one 128-byte row (64 channels of one tap): slot k & 1
          const int pair = chunk >> 1, half = chunk & 1;
          const uint32_t slot = my_stage + static_cast<uint32_t>((pair & 1) * 4096);
          if (half == 0) {
            // the store issued two pairs ago read this slot: at most the previous pair's store may still be in flight
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
            sts128(slot + static_cast<uint32_t>((((half << 2) | i) ^ (lane & 7)) << 4), pack8<OutT>(x + 8 * i));
          if (half == 1) {
            fence_proxy_async();                                   // generic-proxy writes -> visible to the TMA engine
            __syncwarp();
            if (lane == 0) {
              const int gcol = nt * kTBlockN + pair * 64;          // column of the [pixels] x [4*Cout] product
              const int tap = gcol / p.cout;
              const int co0 = gcol - tap * p.cout;
              tma_store_5d(&tmY, slot, (tap & 1) * p.cout + co0, tx * 8, tap >> 1, ty * 16 + q * 4, img);
              bulk_commit();
            }
          }
        } else if (valid) {
          const int gcol = nt * kTBlockN + chunk * 32;             // column of the [pixels] x [4*Cout] product
          const int tap = gcol / p.cout;
          const int co0 = gcol - tap * p.cout;
          const size_t opix = (static_cast<size_t>(img) * (2 * p.h) + (2 * hh + (tap >> 1))) * out_w + (2 * ww + (tap & 1));
          OutT* yrow = reinterpret_cast<OutT*>(p.y) + opix * p.cout + co0;
          store_chunk32<OutT>(yrow, x);
        }
      };
      // the TMEM load of chunk c+1 is in flight while chunk c is reduced and stored
      uint32_t ra[32], rb[32];
      tmem_ld_32x32(tbase, ra);
#pragma unroll 1
      for (int chunk = 0; chunk < kTBlockN / 32; chunk += 2) {
        tmem_ld_wait();
        tmem_ld_32x32(tbase + (chunk + 1) * 32, rb);
        process(ra, chunk);
        tmem_ld_wait();
        if (chunk + 2 < kTBlockN / 32) tmem_ld_32x32(tbase + (chunk + 2) * 32, ra);
        process(rb, chunk + 1);
      }
      // the accumulator has been read: release it before the (shared-memory only) statistics reduction
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[buf]);
      pf ^= 1;
      if (p.sgs_log2 >= 0) {
        if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");   // all quarter-tile partials of this group are in smem
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        const int nslots = (kTBlockN * 2) >> p.sgs_log2;
        const int slots_per_row = (p.cout * 2) >> p.sgs_log2;
        for (int e = et; e < nslots; e += 128) {
          const float s = grp_scratch[e] + grp_scratch[256 + e] + grp_scratch[512 + e] + grp_scratch[768 + e];
          const int gcol = nt * kTBlockN + ((e >> 1) << p.sgs_log2);
          const int tap = gcol / p.cout;
          const int cot = gcol - tap * p.cout;
          p.partials[((static_cast<size_t>(img) * tiles_per_image + r) * 4 + tap) * slots_per_row + ((cot >> p.sgs_log2) << 1) + (e & 1)] = s;
        }
        if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");   // scratch may be overwritten by this group's next item
        else asm volatile("bar.sync 2, 128;" ::: "memory");
      }
    }
  }
  if (kTmaStore && warp >= 2 && lane == 0) bulk_wait<0>();       // staging slots are read (and the stores complete) before the CTA exits
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ----------------------------------------------------------------------------- host side
static int convT_v2_check(const b2u_conv_desc* d) {
  int rc = conv_validate_desc(d);
  if (rc) return rc;
  B2U_REQUIRE((4 * d->cout) % kTBlockN == 0, "convT v2 needs 4*cout to be a multiple of %d", kTBlockN);
  return B2U_OK;
}

int convT_v2_stat_layout(const b2u_conv_desc* d, int* rows_per_image, int* subgroup_size) {
  int rc = convT_v2_check(d);
  if (rc) return rc;
  if (rows_per_image) *rows_per_image = ((d->w + 7) / 8) * ((d->h + 15) / 16) * 4;
  if (subgroup_size) *subgroup_size = conv_stat_subgroup(d->cout, d->num_groups);
  return B2U_OK;
}

template <int TF, bool TS>
static int convT_v2_launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ty, const ConvTParams& gp, int grid, size_t smem,
                           cudaStream_t st) {
  B2U_SET_MAX_SMEM_ONCE((convT_v2_kernel<TF, TS>), 227 * 1024);
  B2U_PDL_LAUNCH((convT_v2_kernel<TF, TS>), grid, kTThreads, smem, st, ta, tb, ty, gp);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

// TMA-store epilogue: 16-bit outputs (a 64-channel run is one 128-byte swizzle row).  B2U_CONVT_TMA_STORE=0 / reserved[1] == 1
// select the per-thread store path (A/B runs); the fp32-storage (tf32) kernels always use it.
static bool convT_use_tma_store(const b2u_conv_desc* d) {
  if (d->dtype == B2U_F32 || d->cout % 64 != 0) return false;
  if (d->reserved[1] == 1) return false;
  if (d->reserved[1] == 2) return true;
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("B2U_CONVT_TMA_STORE");
    env = (e && e[0] == '0') ? 0 : 1;
  }
  return env != 0;
}

int convT_v2_run(const void* x, const void* wpacked, void* y, float* partials, const b2u_conv_desc* d, void* stream) {
  int rc = convT_v2_check(d);
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
    cuuint32_t box[4] = {static_cast<cuuint32_t>(ke), 8, 16, 1};
    rc = conv_encode_map(&ta, d->dtype, 4, x, dims, strides, box);
    if (rc) return rc;
  }
  {
    // packed weight [tap][cout][cin] viewed as a K-major matrix of 4*cout rows
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(d->cin), static_cast<cuuint64_t>(4) * d->cout};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(d->cin) * es};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(ke), static_cast<cuuint32_t>(kTBlockN)};
    rc = conv_encode_map(&tb, d->dtype, 2, wpacked, dims, strides, box);
    if (rc) return rc;
  }
  const bool tma_store = convT_use_tma_store(d);
  CUtensorMap ty = ta;
  if (tma_store) {
    // output [n][2h][2w][cout] viewed as [n][h][dy][w][dx * cout + c]: the box {64, 8, 1, 4, 1} at (dx * cout + co0, w0, dy, h0, n)
    // is the pixel-shuffled destination of 4 x 8 input pixels for one tap; w and h are clipped at the tensor bounds
    cuuint64_t dims[5] = {static_cast<cuuint64_t>(2) * d->cout, static_cast<cuuint64_t>(d->w), 2, static_cast<cuuint64_t>(d->h),
                          static_cast<cuuint64_t>(d->n)};
    const cuuint64_t row = static_cast<cuuint64_t>(2) * d->w * d->cout * es;          // one output row
    cuuint64_t strides[4] = {static_cast<cuuint64_t>(2) * d->cout * es, row, 2 * row, 2 * row * d->h};
    cuuint32_t box[5] = {64, 8, 1, 4, 1};
    rc = conv_encode_map(&ty, d->dtype, 5, y, dims, strides, box);
    if (rc) return rc;
  }
  ConvTParams gp;
  gp.n = d->n; gp.h = d->h; gp.w = d->w;
  gp.tiles_w = (d->w + 7) / 8; gp.tiles_h = (d->h + 15) / 16;
  gp.total_tiles = d->n * gp.tiles_w * gp.tiles_h;
  gp.num_items = gp.total_tiles * (4 * d->cout / kTBlockN);
  gp.kc_chunks = d->cin / ke;
  gp.cout = d->cout;
  // the staging slots (8 warps x 2 x 4 KB) take the place of the fourth pipeline stage
  gp.stages = tma_store ? 3 : 4;
  const int sgs = conv_stat_subgroup(d->cout, d->num_groups);
  gp.sgs_log2 = sgs > 0 ? conv_ilog2(sgs) : -1;
  gp.y = y;
  gp.partials = partials;
  size_t smem = static_cast<size_t>(gp.stages) * (kTABytes + kTBBytes) + 1024 + (2 * gp.stages + 4) * 8 + 16 + 2 * 4 * 256 * 4;
  gp.stage_off = 0;
  if (tma_store) {
    gp.stage_off = static_cast<int>((smem - 1024 + 1023) & ~static_cast<size_t>(1023));   // offset from the ALIGNED base
    smem = static_cast<size_t>(gp.stage_off) + 8 * 8192 + 1024;
  }
  int grid = b2u_num_sms();
  if (grid > gp.num_items) grid = gp.num_items;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (d->dtype == B2U_F32) return convT_v2_launch<1, false>(ta, tb, ty, gp, grid, smem, st);
  if (d->dtype == B2U_F16) return tma_store ? convT_v2_launch<2, true>(ta, tb, ty, gp, grid, smem, st) : convT_v2_launch<2, false>(ta, tb, ty, gp, grid, smem, st);
  return tma_store ? convT_v2_launch<0, true>(ta, tb, ty, gp, grid, smem, st) : convT_v2_launch<0, false>(ta, tb, ty, gp, grid, smem, st);
}

}  // namespace b2u
