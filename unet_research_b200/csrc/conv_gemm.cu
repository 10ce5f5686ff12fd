// Tensor-core convolutions of the U-Net as TMA-fed tcgen05 implicit GEMMs (sm_100a).
//
//   conv3x3 (stride 1, zero "same" padding, no bias):  M = N*H*W pixels, N = Cout, K = 9*Cin
//   convT2x2 (stride 2, no bias):                      M = N*H*W input pixels, N = 4*Cout, K = Cin
//
// A operand: a 4-D TMA box {128 B of channels, bw, bh, 1} of the NHWC activation, shifted by the
// filter tap; TMA's out-of-bounds zero fill IS the convolution's zero padding.  The box lands in
// shared memory as 128 rows (pixels) x 128 B with the 128-byte swizzle, which is exactly the K-major
// UMMA operand layout, so the tile feeds tcgen05.mma with no register staging.
// B operand: packed weights [tap][Cout][Cin] (K-major), box {128 B, BLOCK_N, 1}.
// D: fp32 accumulators in TMEM (128 lanes x BLOCK_N columns).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (tcgen05.ld -> GroupNorm partial sums -> bf16/fp32 -> global).
// One output tile per CTA; several CTAs co-reside per SM (shared memory permitting) so one CTA's
// epilogue overlaps another's main loop.
#include "b2u_common.cuh"
#include "conv_host.cuh"

#include <mutex>

namespace b2u {

struct GemmParams {
  int n, h, w;                 // input pixel grid
  int tiles_w, tiles_h;        // tiles per image
  int bw_log2, bh;             // TMA box (bw * bh == 128)
  int taps;                    // K taps iterated by one CTA: 9 (conv3x3) or 1 (convT)
  int kc_per_tap;              // Cin / K-elements-per-block
  int cout;                    // output channels (per tap for convT)
  int mode;                    // 0 = conv3x3, 1 = convT2x2
  int n_tiles_per_tap;         // cout / BLOCK_N
  int out_h, out_w;
  int stages;
  int sgs_log2;                // log2(statistics sub-group size), -1 = no statistics
  int rows_per_image;          // rows of the partial buffer per image
  void* y;
  float* partials;
};

constexpr int kBlockM = 128;
constexpr int kABytes = kBlockM * 128;           // 16 KB per stage
constexpr int kNumThreads = 192;

template <int NV>
__device__ __forceinline__ void epilogue_stats(const float (&x)[32], bool valid, int sgs_log2, int lane,
                                               float* scratch /* this warp, this chunk: NV floats */) {
  constexpr int NSG = NV / 2;                  // sub-groups inside the 32-column chunk
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
  constexpr int LPV = 32 / NV;                 // lanes holding the same value (NV <= 32)
  if (lane % LPV == 0) scratch[lane / LPV] = v[0];
}

template <int BLOCK_N, int kFmt>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmParams p) {
  constexpr bool kTf32 = kFmt == 1;
  using OutT = typename FmtTraits<kFmt>::T;
  constexpr int kBBytes = BLOCK_N * 128;
  constexpr int kKElems = kTf32 ? 32 : 64;       // K elements per 128-byte row
  constexpr int kUmmaK = kTf32 ? 8 : 16;         // 32 bytes of K per tcgen05.mma
  constexpr uint32_t kIdesc = umma_idesc(kBlockM, BLOCK_N, FmtTraits<kFmt>::kIdescFmt);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = p.stages;
  uint8_t* smA = smem;
  uint8_t* smB = smem + stages * kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smB + stages * kBBytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full_bar = empty_bar + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* stat_scratch = reinterpret_cast<float*>(tmem_slot + 2);   // [4 warps][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates
  int mt = blockIdx.x;
  const int tiles_per_image = p.tiles_w * p.tiles_h;
  const int img = mt / tiles_per_image;
  const int mt_in_img = mt - img * tiles_per_image;
  const int ty = mt_in_img / p.tiles_w;
  const int tx = mt_in_img - ty * p.tiles_w;
  const int bw = 1 << p.bw_log2;
  const int h0 = ty * p.bh, w0 = tx * bw;
  const int nt = blockIdx.y;
  const int tap_fixed = nt / p.n_tiles_per_tap;              // convT: tap of this N tile (conv: 0)
  const int n0 = (nt - tap_fixed * p.n_tiles_per_tap) * BLOCK_N;
  const int num_kb = p.taps * p.kc_per_tap;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BLOCK_N);
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
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;
      for (int tap = 0; tap < p.taps; ++tap) {
        int dy = 0, dx = 0, btap = p.mode == 1 ? tap_fixed : 0;
        if (p.mode == 0) {
          dy = tap / 3 - 1;
          dx = tap % 3 - 1;
          btap = tap;
        }
        for (int kc = 0; kc < p.kc_per_tap; ++kc) {
          mbar_wait(&empty_bar[s], ph);
          mbar_arrive_expect_tx(&full_bar[s], kABytes + kBBytes);
          tma_load_4d(smA + s * kABytes, &tmA, &full_bar[s], kc * kKElems, w0 + dx, h0 + dy, img);
          tma_load_3d(smB + s * kBBytes, &tmB, &full_bar[s], kc * kKElems, n0, btap);
          if (++s == stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // the whole warp runs the loop (converged control flow); one elected lane issues (see umma_ss_conv)
    {
      const uint32_t leader = elect_one() ? 1u : 0u;
      const uint32_t tmem_u = warp_uniform(tmem_base);
      int s = 0;
      uint32_t ph = 0;
      const uint64_t adesc0 = umma_desc_k_sw128(smem_u32(smA));
      const uint64_t bdesc0 = umma_desc_k_sw128(smem_u32(smB));
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint64_t adesc = adesc0 + static_cast<uint64_t>((s * kABytes) >> 4);
        const uint64_t bdesc = bdesc0 + static_cast<uint64_t>((s * kBBytes) >> 4);
        // +32 bytes of K inside the 128-byte swizzle atom = +2 in the 16-byte address field (four steps per stage)
        static_assert(kKElems / kUmmaK == 4, "one swizzle row = four K steps");
        umma_ss_conv4<kTf32>(tmem_u, adesc, bdesc, kIdesc, kb != 0 ? 1u : 0u, leader);
        umma_commit_conv(&empty_bar[s], leader);           // frees the smem stage once these MMAs retire
        if (++s == stages) { s = 0; ph ^= 1; }
      }
      umma_commit_conv(tmem_full_bar, leader);             // accumulator complete
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                   // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;            // accumulator row = pixel inside the tile
    const int dh = row >> p.bw_log2, dw = row & (bw - 1);
    const int hh = h0 + dh, ww = w0 + dw;
    const bool valid = (hh < p.h) && (ww < p.w);
    size_t out_pix;
    if (p.mode != 1) {
      out_pix = (static_cast<size_t>(img) * p.out_h + hh) * p.out_w + ww;
    } else {
      out_pix = (static_cast<size_t>(img) * p.out_h + (2 * hh + (tap_fixed >> 1))) * p.out_w + (2 * ww + (tap_fixed & 1));
    }
    OutT* yrow = reinterpret_cast<OutT*>(p.y) + out_pix * p.cout + n0;
    float* my_scratch = stat_scratch + q * 128;

    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + chunk * 32, r);
      tmem_ld_wait();
      float x[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(r[i]);
      if (p.sgs_log2 >= 0) {
        switch (p.sgs_log2) {
          case 1: epilogue_stats<32>(x, valid, 1, lane, my_scratch + chunk * 32); break;
          case 2: epilogue_stats<16>(x, valid, 2, lane, my_scratch + chunk * 16); break;
          case 3: epilogue_stats<8>(x, valid, 3, lane, my_scratch + chunk * 8); break;
          case 4: epilogue_stats<4>(x, valid, 4, lane, my_scratch + chunk * 4); break;
          default: epilogue_stats<2>(x, valid, 5, lane, my_scratch + chunk * 2); break;
        }
      }
      if (valid) {
        store_chunk32<OutT>(yrow + chunk * 32, x);
      }
    }
    if (p.sgs_log2 >= 0) {
      // deterministic cross-warp sum of the four quarter-tile partials, then one plain store per slot
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int nslots = (BLOCK_N * 2) >> p.sgs_log2;
      const int t = threadIdx.x - 64;
      if (t < nslots) {
        float s = stat_scratch[t] + stat_scratch[128 + t] + stat_scratch[256 + t] + stat_scratch[384 + t];
        const int row_idx = (p.mode != 1) ? mt_in_img : mt_in_img * 4 + tap_fixed;
        const int slots_per_row = (p.cout * 2) >> p.sgs_log2;
        p.partials[(static_cast<size_t>(img) * p.rows_per_image + row_idx) * slots_per_row +
                   ((n0 * 2) >> p.sgs_log2) + t] = s;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BLOCK_N);
  }
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
  });
  return fn;
}

int conv_encode_map(CUtensorMap* map, int dtype, int rank, const void* base, const cuuint64_t* dims,
                    const cuuint64_t* strides_bytes, const cuuint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    b2u_set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return B2U_ERR_CUDA;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, dtype == B2U_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (dtype == B2U_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16), rank,
                  const_cast<void*>(base), dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    b2u_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", static_cast<int>(r), rank);
    return B2U_ERR_CUDA;
  }
  return B2U_OK;
}

int conv_ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

// Pick the 128-pixel box (bw x bh, bw a power of two) that wastes the fewest accumulator rows.
static void choose_box(int h, int w, int* bw_out, int* bh_out) {
  long best = -1;
  const int cands[6] = {16, 8, 32, 64, 128, 4};
  for (int i = 0; i < 6; ++i) {
    int bw = cands[i], bh = 128 / bw;
    long tiles = static_cast<long>((w + bw - 1) / bw) * ((h + bh - 1) / bh);
    if (best < 0 || tiles < best) {
      best = tiles;
      *bw_out = bw;
      *bh_out = bh;
    }
  }
}

int conv_stat_subgroup(int cout, int num_groups) {
  if (num_groups <= 0) return 0;
  int gsize = cout / num_groups;
  return gsize < 32 ? gsize : 32;
}

struct Plan {
  int bw, bh, tiles_w, tiles_h, block_n, stages, sgs;
  size_t smem;
};

int conv_validate_desc(const b2u_conv_desc* d) {
  B2U_REQUIRE(d != nullptr, "null descriptor");
  B2U_REQUIRE(d->dtype == B2U_BF16 || d->dtype == B2U_F32 || d->dtype == B2U_F16, "dtype must be B2U_BF16, B2U_F16 or B2U_F32");
  const int ke = d->dtype == B2U_F32 ? 32 : 64;
  B2U_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0, "empty tensor (n=%d h=%d w=%d)", d->n, d->h, d->w);
  B2U_REQUIRE(d->cin > 0 && d->cin % ke == 0, "cin=%d must be a positive multiple of %d", d->cin, ke);
  B2U_REQUIRE(d->cout > 0 && d->cout % 64 == 0, "cout=%d must be a positive multiple of 64", d->cout);
  B2U_REQUIRE(d->x_cstride >= d->cin && d->x_cstride % 8 == 0, "x_cstride=%d must be >= cin and a multiple of 8",
              d->x_cstride);
  if (d->num_groups > 0) {
    B2U_REQUIRE(d->cout % d->num_groups == 0, "cout=%d not divisible by num_groups=%d", d->cout, d->num_groups);
    int gs = d->cout / d->num_groups;
    B2U_REQUIRE(gs >= 2 && (gs & (gs - 1)) == 0, "group size %d must be a power of two >= 2", gs);
  }
  return B2U_OK;
}

static int make_plan(const b2u_conv_desc* d, bool conv_t, Plan* pl) {
  int vrc = conv_validate_desc(d);
  if (vrc) return vrc;
  choose_box(d->h, d->w, &pl->bw, &pl->bh);
  pl->tiles_w = (d->w + pl->bw - 1) / pl->bw;
  pl->tiles_h = (d->h + pl->bh - 1) / pl->bh;
  int bn = d->cout % 256 == 0 ? 256 : (d->cout % 128 == 0 ? 128 : 64);
  if (d->reserved[0] == 64 || d->reserved[0] == 128 || d->reserved[0] == 256) {
    B2U_REQUIRE(d->cout % d->reserved[0] == 0, "BLOCK_N override %d does not divide cout", d->reserved[0]);
    bn = d->reserved[0];
  }
  pl->block_n = bn;
  const int stage_bytes = kABytes + bn * 128;
  int stages = bn == 256 ? 4 : (bn == 128 ? 3 : 4);     // 192 KB / 96 KB / 96 KB: the small tiles co-reside 2 per SM
  if (d->reserved[1] >= 2 && d->reserved[1] <= 8) stages = d->reserved[1];
  while (stages > 2 && static_cast<size_t>(stages) * stage_bytes + 4096 > 227 * 1024) --stages;
  pl->stages = stages;
  pl->smem = static_cast<size_t>(stages) * stage_bytes + 1024 /*align*/ + (2 * stages + 1) * 8 + 16 + 4 * 128 * 4;
  pl->sgs = conv_stat_subgroup(d->cout, d->num_groups);
  (void)conv_t;
  return B2U_OK;
}

template <int BN, int TF>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& gp, dim3 grid, size_t smem,
                       cudaStream_t st) {
  B2U_SET_MAX_SMEM_ONCE((conv_gemm_kernel<BN, TF>), 227 * 1024);
  B2U_PDL_LAUNCH((conv_gemm_kernel<BN, TF>), grid, kNumThreads, smem, st, ta, tb, gp);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

// mode 0 = conv3x3 (v1), 1 = convT2x2, 2 = plain 1x1 GEMM
static int run_gemm(const void* x, const void* wpacked, void* y, float* partials, const b2u_conv_desc* d,
                    int mode, void* stream) {
  const bool conv_t = mode == 1;
  Plan pl;
  int rc = make_plan(d, conv_t, &pl);
  if (rc) return rc;
  B2U_REQUIRE(x && wpacked && y, "null tensor pointer");
  B2U_REQUIRE(d->num_groups == 0 || partials != nullptr, "partials required when num_groups > 0");
  const int es = d->dtype == B2U_F32 ? 4 : 2;
  const int ke = 128 / es;
  const int taps_w = mode == 1 ? 4 : (mode == 0 ? 9 : 1);

  CUtensorMap ta, tb;
  {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(d->x_cstride), static_cast<cuuint64_t>(d->w),
                          static_cast<cuuint64_t>(d->h), static_cast<cuuint64_t>(d->n)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(d->x_cstride) * es,
                             static_cast<cuuint64_t>(d->w) * d->x_cstride * es,
                             static_cast<cuuint64_t>(d->h) * d->w * d->x_cstride * es};
    cuuint32_t box[4] = {static_cast<cuuint32_t>(ke), static_cast<cuuint32_t>(pl.bw), static_cast<cuuint32_t>(pl.bh), 1};
    rc = conv_encode_map(&ta, d->dtype, 4, x, dims, strides, box);
    if (rc) return rc;
  }
  {
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(d->cin), static_cast<cuuint64_t>(d->cout),
                          static_cast<cuuint64_t>(taps_w)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(d->cin) * es, static_cast<cuuint64_t>(d->cout) * d->cin * es};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(ke), static_cast<cuuint32_t>(pl.block_n), 1};
    rc = conv_encode_map(&tb, d->dtype, 3, wpacked, dims, strides, box);
    if (rc) return rc;
  }

  GemmParams gp;
  gp.n = d->n; gp.h = d->h; gp.w = d->w;
  gp.tiles_w = pl.tiles_w; gp.tiles_h = pl.tiles_h;
  gp.bw_log2 = conv_ilog2(pl.bw); gp.bh = pl.bh;
  gp.taps = mode == 0 ? 9 : 1;
  gp.kc_per_tap = d->cin / ke;
  gp.cout = d->cout;
  gp.mode = mode;
  gp.n_tiles_per_tap = d->cout / pl.block_n;
  gp.out_h = conv_t ? 2 * d->h : d->h;
  gp.out_w = conv_t ? 2 * d->w : d->w;
  gp.stages = pl.stages;
  gp.sgs_log2 = pl.sgs > 0 ? conv_ilog2(pl.sgs) : -1;
  gp.rows_per_image = pl.tiles_w * pl.tiles_h * (conv_t ? 4 : 1);
  gp.y = y;
  gp.partials = partials;

  dim3 grid(static_cast<unsigned>(d->n) * pl.tiles_w * pl.tiles_h, gp.n_tiles_per_tap * (conv_t ? 4 : 1));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define B2U_V1_CASE(BN)                                                                   \
  if (d->dtype == B2U_F32) return launch_gemm<BN, 1>(ta, tb, gp, grid, pl.smem, st);         \
  if (d->dtype == B2U_F16) return launch_gemm<BN, 2>(ta, tb, gp, grid, pl.smem, st);         \
  return launch_gemm<BN, 0>(ta, tb, gp, grid, pl.smem, st);
  switch (pl.block_n) {
    case 64: B2U_V1_CASE(64)
    case 128: B2U_V1_CASE(128)
    default: B2U_V1_CASE(256)
  }
#undef B2U_V1_CASE
}

// ----------------------------------------------------------------------------- weight packing
template <typename T>
__global__ void pack_conv3x3_kernel(const float* __restrict__ w, T* __restrict__ out, int cout, int cin, int tflip) {
  // out[tap][row][col]; forward: row = co, col = ci, tap = r*3+s.  transpose_flip: row = ci, col = co, tap' = 8-tap.
  long total = 9L * cout * cin;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    int rows = tflip ? cin : cout, cols = tflip ? cout : cin;
    int col = static_cast<int>(i % cols);
    int row = static_cast<int>((i / cols) % rows);
    int tap = static_cast<int>(i / (static_cast<long>(cols) * rows));
    int co = tflip ? col : row, ci = tflip ? row : col;
    int src_tap = tflip ? 8 - tap : tap;
    out[i] = to_operand<T>(w[(static_cast<long>(co) * cin + ci) * 9 + src_tap]);
  }
}
// Both packed layouts of one 3x3 weight in a single pass (training: the weights change every optimiser step).
// A block stages a 32 (co) x 32 (ci) x 9 tile through shared memory: the fp32 source is read once, fully
// coalesced (288 contiguous floats per output channel), and both destinations are written as 64-byte runs:
//   fwd  [tap][co][ci]      dgrad [8 - tap][ci][co]   (transposed, taps rotated by 180 degrees).
template <typename T>
__global__ void __launch_bounds__(256) pack_conv3x3_pair_kernel(const float* __restrict__ w, T* __restrict__ fwd, T* __restrict__ dgrad,
                                                                int cout, int cin) {
  pdl_wait();       // predecessor grid complete + visible (no-op without the PDL launch attribute)
  pdl_trigger();    // the successor may be scheduled once every CTA got here
  __shared__ T tile[9][32][33];
  const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
  for (int f = threadIdx.x; f < 32 * 288; f += 256) {
    const int co_l = f / 288, rem = f - co_l * 288;
    const int ci_l = rem / 9, tap = rem - ci_l * 9;
    tile[tap][co_l][ci_l] = to_operand<T>(w[(static_cast<long>(co0 + co_l) * cin + ci0) * 9 + rem]);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  for (int r = wrp; r < 9 * 32; r += 8) {
    const int tap = r >> 5, row = r & 31;
    fwd[(static_cast<long>(tap) * cout + co0 + row) * cin + ci0 + lane] = tile[tap][row][lane];
    if (dgrad) dgrad[(static_cast<long>(8 - tap) * cin + ci0 + row) * cout + co0 + lane] = tile[tap][lane][row];
  }
}

template <typename T>
__global__ void pack_convT_kernel(const float* __restrict__ w, T* __restrict__ out, int cin, int cout) {
  pdl_wait();       // predecessor grid complete + visible (no-op without the PDL launch attribute)
  pdl_trigger();    // the successor may be scheduled once every CTA got here
  long total = 4L * cout * cin;     // out[tap][co][ci] = w[ci][co][tap]
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    int ci = static_cast<int>(i % cin);
    int co = static_cast<int>((i / cin) % cout);
    int tap = static_cast<int>(i / (static_cast<long>(cin) * cout));
    out[i] = to_operand<T>(w[(static_cast<long>(ci) * cout + co) * 4 + tap]);
  }
}

// Batched repack (b2u_pack_batched): block -> (tensor, 32 x 32 tile) through the table's first_block prefix; the tile
// bodies are the per-tensor kernels' (same rounding, same outputs).
template <typename T>
__global__ void __launch_bounds__(256) pack_batched_kernel(const b2u_pack_entry* __restrict__ table, int n) {
  pdl_wait();       // predecessor grid complete + visible (no-op without the PDL launch attribute)
  pdl_trigger();    // the successor may be scheduled once every CTA got here
  __shared__ T tile[9][32][33];
  __shared__ int s_entry;
  if (threadIdx.x == 0) {
    int e = 0;
    while (e + 1 < n && static_cast<int>(blockIdx.x) >= table[e + 1].first_block) ++e;
    s_entry = e;
  }
  __syncthreads();
  const b2u_pack_entry en = table[s_entry];
  const int b = blockIdx.x - en.first_block;
  const int cin = en.cin, cout = en.cout;
  const int tiles_ci = cin >> 5;
  const int ci0 = (b % tiles_ci) * 32, co0 = (b / tiles_ci) * 32;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const float* __restrict__ w = en.w;
  if (en.kind == 0) {
    T* __restrict__ fwd = static_cast<T*>(en.out0);
    T* __restrict__ dgrad = static_cast<T*>(en.out1);
    for (int f = threadIdx.x; f < 32 * 288; f += 256) {
      const int co_l = f / 288, rem = f - co_l * 288;
      const int ci_l = rem / 9, tap = rem - ci_l * 9;
      tile[tap][co_l][ci_l] = to_operand<T>(w[(static_cast<long>(co0 + co_l) * cin + ci0) * 9 + rem]);
    }
    __syncthreads();
    for (int r = wrp; r < 9 * 32; r += 8) {
      const int tap = r >> 5, row = r & 31;
      fwd[(static_cast<long>(tap) * cout + co0 + row) * cin + ci0 + lane] = tile[tap][row][lane];
      if (dgrad) dgrad[(static_cast<long>(8 - tap) * cin + ci0 + row) * cout + co0 + lane] = tile[tap][lane][row];
    }
  } else {
    // w[ci][co][tap]: a ci row of the tile is 128 contiguous floats
    T* __restrict__ fwd = static_cast<T*>(en.out0);            // [tap][co][ci]
    T* __restrict__ dgrad = static_cast<T*>(en.out1);          // [ci][tap * cout + co]
    for (int f = threadIdx.x; f < 32 * 128; f += 256) {
      const int ci_l = f >> 7, rem = f & 127;
      tile[rem & 3][ci_l][rem >> 2] = to_operand<T>(w[(static_cast<long>(ci0 + ci_l) * cout + co0) * 4 + rem]);
    }
    __syncthreads();
    for (int r = wrp; r < 4 * 32; r += 8) {
      const int tap = r >> 5, row = r & 31;
      fwd[(static_cast<long>(tap) * cout + co0 + row) * cin + ci0 + lane] = tile[tap][lane][row];
      if (dgrad) dgrad[static_cast<long>(ci0 + row) * (4 * cout) + tap * cout + co0 + lane] = tile[tap][row][lane];
    }
  }
}

}  // namespace b2u

using namespace b2u;

// reserved[2] == 1 selects the v1 kernel (one 16x8 tile per CTA, per-tap activation reload); default is v2.
static bool use_v1(const b2u_conv_desc* d) { return d != nullptr && d->reserved[2] == 1; }

extern "C" int b2u_conv3x3_stat_layout(const b2u_conv_desc* d, int* rows_per_image, int* subgroup_size) {
  if (!use_v1(d)) return conv3x3_v2_stat_layout(d, rows_per_image, subgroup_size);
  Plan pl;
  int rc = make_plan(d, false, &pl);
  if (rc) return rc;
  if (rows_per_image) *rows_per_image = pl.tiles_w * pl.tiles_h;
  if (subgroup_size) *subgroup_size = pl.sgs;
  return B2U_OK;
}
extern "C" int b2u_convT2x2_stat_layout(const b2u_conv_desc* d, int* rows_per_image, int* subgroup_size) {
  if (!use_v1(d)) return convT_v2_stat_layout(d, rows_per_image, subgroup_size);
  Plan pl;
  int rc = make_plan(d, true, &pl);
  if (rc) return rc;
  if (rows_per_image) *rows_per_image = pl.tiles_w * pl.tiles_h * 4;
  if (subgroup_size) *subgroup_size = pl.sgs;
  return B2U_OK;
}
extern "C" int b2u_conv3x3_fwd(const void* x, const void* wpacked, void* y, float* partials, const b2u_conv_desc* d,
                               void* stream) {
  if (!use_v1(d)) return conv3x3_v2_run(x, wpacked, y, partials, d, stream);
  return run_gemm(x, wpacked, y, partials, d, 0, stream);
}
extern "C" int b2u_conv3x3_pro_fwd(const void* x_raw, const float* coef, const void* mask_bits, const void* wpacked, void* y,
                                   float* partials, const b2u_conv_desc* d, int relu, int x_shared, void* stream) {
  B2U_REQUIRE(d != nullptr, "null descriptor");
  V2Prologue pro;
  pro.coef = coef;
  pro.mask = mask_bits;
  pro.relu = relu;
  pro.x_shared = x_shared;
  return conv3x3_v2_run(x_raw, wpacked, y, partials, d, stream, &pro);
}
extern "C" int b2u_convT2x2_fwd(const void* x, const void* wpacked, void* y, float* partials, const b2u_conv_desc* d,
                                void* stream) {
  if (!use_v1(d)) return convT_v2_run(x, wpacked, y, partials, d, stream);
  return run_gemm(x, wpacked, y, partials, d, 1, stream);
}
extern "C" int b2u_gemm1x1_fwd(const void* x, const void* wpacked, void* y, const b2u_conv_desc* d, void* stream) {
  B2U_REQUIRE(d && d->num_groups == 0, "b2u_gemm1x1_fwd computes no statistics: set num_groups = 0");
  return run_gemm(x, wpacked, y, nullptr, d, 2, stream);
}
extern "C" int b2u_pack_conv3x3_weight(const float* w, void* packed, int cout, int cin, int dtype, int transpose_flip,
                                       void* stream) {
  B2U_REQUIRE(w && packed && cout > 0 && cin > 0, "bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  long total = 9L * cout * cin;
  int blocks = static_cast<int>((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  if (dtype == B2U_F32) pack_conv3x3_kernel<float><<<blocks, 256, 0, st>>>(w, static_cast<float*>(packed), cout, cin, transpose_flip);
  else if (dtype == B2U_F16) pack_conv3x3_kernel<__half><<<blocks, 256, 0, st>>>(w, static_cast<__half*>(packed), cout, cin, transpose_flip);
  else pack_conv3x3_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(w, static_cast<__nv_bfloat16*>(packed), cout, cin, transpose_flip);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}
extern "C" int b2u_pack_conv3x3_weight_pair(const float* w, void* packed_fwd, void* packed_dgrad, int cout, int cin, int dtype,
                                            void* stream) {
  B2U_REQUIRE(w && packed_fwd && cout > 0 && cin > 0, "bad arguments");
  B2U_REQUIRE(cout % 32 == 0 && cin % 32 == 0, "pair packing needs cout and cin to be multiples of 32 (got %d, %d)", cout, cin);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(cin / 32, cout / 32);
  if (dtype == B2U_F32)
    B2U_PDL_LAUNCH((pack_conv3x3_pair_kernel<float>), grid, 256, 0, st, w, static_cast<float*>(packed_fwd), static_cast<float*>(packed_dgrad), cout, cin);
  else if (dtype == B2U_F16)
    B2U_PDL_LAUNCH((pack_conv3x3_pair_kernel<__half>), grid, 256, 0, st, w, static_cast<__half*>(packed_fwd), static_cast<__half*>(packed_dgrad), cout, cin);
  else
    B2U_PDL_LAUNCH((pack_conv3x3_pair_kernel<__nv_bfloat16>), grid, 256, 0, st, w, static_cast<__nv_bfloat16*>(packed_fwd), static_cast<__nv_bfloat16*>(packed_dgrad), cout, cin);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}
extern "C" int b2u_pack_batched_plan(b2u_pack_entry* entries, int n, int* total_blocks) {
  B2U_REQUIRE(entries && total_blocks && n > 0, "bad arguments");
  int blocks = 0;
  for (int i = 0; i < n; ++i) {
    b2u_pack_entry& e = entries[i];
    B2U_REQUIRE(e.kind == 0 || e.kind == 1, "entry %d: kind must be 0 (Conv2d 3x3) or 1 (ConvTranspose2d 2x2)", i);
    B2U_REQUIRE(e.w && e.out0, "entry %d: null pointer", i);
    B2U_REQUIRE(e.cout > 0 && e.cin > 0 && e.cout % 32 == 0 && e.cin % 32 == 0, "entry %d: cout and cin must be multiples of 32 (got %d, %d)", i,
                e.cout, e.cin);
    e.first_block = blocks;
    blocks += (e.cin / 32) * (e.cout / 32);
  }
  *total_blocks = blocks;
  return B2U_OK;
}
extern "C" int b2u_pack_batched(const b2u_pack_entry* entries_dev, int n, int total_blocks, int dtype, void* stream) {
  B2U_REQUIRE(entries_dev && n > 0 && total_blocks > 0, "bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == B2U_F32) B2U_PDL_LAUNCH((pack_batched_kernel<float>), total_blocks, 256, 0, st, entries_dev, n);
  else if (dtype == B2U_F16) B2U_PDL_LAUNCH((pack_batched_kernel<__half>), total_blocks, 256, 0, st, entries_dev, n);
  else B2U_PDL_LAUNCH((pack_batched_kernel<__nv_bfloat16>), total_blocks, 256, 0, st, entries_dev, n);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}
extern "C" int b2u_pack_convT2x2_weight(const float* w, void* packed, int cin, int cout, int dtype, void* stream) {
  B2U_REQUIRE(w && packed && cout > 0 && cin > 0, "bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  long total = 4L * cout * cin;
  int blocks = static_cast<int>((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  if (dtype == B2U_F32) B2U_PDL_LAUNCH((pack_convT_kernel<float>), blocks, 256, 0, st, w, static_cast<float*>(packed), cin, cout);
  else if (dtype == B2U_F16) B2U_PDL_LAUNCH((pack_convT_kernel<__half>), blocks, 256, 0, st, w, static_cast<__half*>(packed), cin, cout);
  else B2U_PDL_LAUNCH((pack_convT_kernel<__nv_bfloat16>), blocks, 256, 0, st, w, static_cast<__nv_bfloat16*>(packed), cin, cout);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}
