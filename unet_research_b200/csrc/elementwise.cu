// HBM-bound kernels of the U-Net path: first-layer direct conv, GroupNorm finalise, fused
// normalise/DropBlock-mask/ReLU apply (+ 2x2 max-pool, + concat-buffer store), output head with
// Monte-Carlo accumulation, rotation.  All NHWC, 8 channels (16 B bf16 / 32 B fp32) per thread,
// fp32 math, deterministic reductions (no floating-point atomics).
#include "b2u_common.cuh"

#include <math.h>
#include <stdlib.h>

namespace b2u {

template <typename T> __device__ __forceinline__ void round_for_storage(float (&f)[8]) {}
template <> __device__ __forceinline__ void round_for_storage<float>(float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = round_tf32(f[i]);     // the consumer is a kind::tf32 MMA
}

// ---------------------------------------------------------------------------------------------
// Deterministic block reduction of per-thread 8-channel (sum, sumsq) accumulators into one row of
// the GroupNorm partial buffer.  Thread t owns channel vector cv = t % cvs and pixel slot t / cvs.
// smem: float[blockDim.x * 16] + float[c * 2].
__device__ void block_stats_to_partials(const float (&s)[8], const float (&q)[8], int c, int sgs,
                                        float* __restrict__ smem, float* __restrict__ out_row) {
  const int cvs = c >> 3;
  const int t = threadIdx.x;
  const int slots = blockDim.x / cvs;
  float* mine = smem + t * 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mine[i] = s[i];
    mine[8 + i] = q[i];
  }
  __syncthreads();
  float* chan = smem + blockDim.x * 16;            // [c][2]
  for (int o = t; o < c * 2; o += blockDim.x) {
    const int ch = o >> 1, st = o & 1;
    const int cv = ch >> 3, ci = ch & 7;
    float acc = 0.f;
    for (int sl = 0; sl < slots; ++sl) acc += smem[(sl * cvs + cv) * 16 + st * 8 + ci];
    chan[o] = acc;
  }
  __syncthreads();
  const int nsg = c / sgs;
  for (int o = t; o < nsg * 2; o += blockDim.x) {
    const int sg = o >> 1, st = o & 1;
    float acc = 0.f;
    for (int i = 0; i < sgs; ++i) acc += chan[(sg * sgs + i) * 2 + st];
    out_row[o] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// First layer: direct 3x3 conv, Cin in {1,3}, fp32 NCHW input read with autopad semantics.
constexpr int kFirstStrip = 4;
template <typename T, int CIN>
__global__ void __launch_bounds__(256) conv_first_kernel(const float* __restrict__ x, const float* __restrict__ wgt, T* __restrict__ y,
                                  float* __restrict__ partials, int h0, int w0, int h, int w, int cout, int sgs) {
  pdl_wait();       // predecessor grid complete + visible (no-op without the PDL launch attribute)
  pdl_trigger();    // the successor may be scheduled once every CTA got here
  extern __shared__ float sm[];
  float* wsm = sm;                                  // [CIN*9][cout]: tap-major, so a thread's 8 channels are two LDS.128
  float* red = sm + cout * CIN * 9;
  for (int i = threadIdx.x; i < cout * CIN * 9; i += blockDim.x) {
    const int co = i / (CIN * 9), j = i - co * (CIN * 9);
    wsm[j * cout + co] = wgt[i];
  }
  __syncthreads();
  const int n = blockIdx.y;
  const int cvs = cout >> 3;
  const int cv = threadIdx.x % cvs;
  const int slot = threadIdx.x / cvs;
  const int slots = blockDim.x / cvs;
  const float* xn = x + static_cast<size_t>(n) * CIN * h0 * w0;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  // one thread = 8 output channels x a strip of kFirstStrip pixels along W: the 3 x (strip + 2) input window is
  // loaded once, every weight vector (2 x LDS.128) feeds kFirstStrip x 8 FMAs
  constexpr int S = kFirstStrip;
  const int strips_w = w / S;
  const int nstrips = h * strips_w;
  for (int strip = blockIdx.x * slots + slot; strip < nstrips; strip += gridDim.x * slots) {
    const int ph = strip / strips_w, pw = (strip - ph * strips_w) * S;
    float in[CIN][3][S + 2];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < S + 2; ++c) {
          const int yy = ph + r - 1, xx = pw + c - 1;
          in[ci][r][c] = (yy >= 0 && yy < h0 && xx >= 0 && xx < w0) ? __ldg(xn + (static_cast<size_t>(ci) * h0 + yy) * w0 + xx) : 0.f;
        }
    float o[S][8];
#pragma unroll
    for (int px = 0; px < S; ++px)
#pragma unroll
      for (int k = 0; k < 8; ++k) o[px][k] = 0.f;
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int j = ci * 9 + r * 3 + c;
          const float4 wa = *reinterpret_cast<const float4*>(wsm + j * cout + cv * 8);
          const float4 wb = *reinterpret_cast<const float4*>(wsm + j * cout + cv * 8 + 4);
#pragma unroll
          for (int px = 0; px < S; ++px) {
            const float v = in[ci][r][px + c];
            o[px][0] = fmaf(v, wa.x, o[px][0]); o[px][1] = fmaf(v, wa.y, o[px][1]);
            o[px][2] = fmaf(v, wa.z, o[px][2]); o[px][3] = fmaf(v, wa.w, o[px][3]);
            o[px][4] = fmaf(v, wb.x, o[px][4]); o[px][5] = fmaf(v, wb.y, o[px][5]);
            o[px][6] = fmaf(v, wb.z, o[px][6]); o[px][7] = fmaf(v, wb.w, o[px][7]);
          }
        }
    const size_t pix0 = static_cast<size_t>(n) * h * w + static_cast<size_t>(ph) * w + pw;
#pragma unroll
    for (int px = 0; px < S; ++px) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        s[k] += o[px][k];
        q[k] += o[px][k] * o[px][k];
      }
      Vec8<T> v;
      v.from_float(o[px]);
      v.store(y + (pix0 + px) * cout + cv * 8);
    }
  }
  if (partials) {
    block_stats_to_partials(s, q, cout, sgs, red,
                            partials + (static_cast<size_t>(n) * gridDim.x + blockIdx.x) * (cout / sgs) * 2);
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void gn_finalize_kernel(const float* __restrict__ partials, int rows, int sgs, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float2* __restrict__ coef, int c, int num_groups,
                                   double count, float eps, const unsigned long long* __restrict__ keep,
                                   int images_per_call, double numel_per_call, float2* __restrict__ mean_rstd,
                                   int shared_partials) {
  pdl_wait();       // predecessor grid complete + visible (no-op without the PDL launch attribute)
  pdl_trigger();    // the successor may be scheduled once every CTA got here
  const int g = blockIdx.x, n = blockIdx.y;
  const int gsize = c / num_groups;
  const int nsg_total = c / sgs;
  const int sg_per_group = gsize / sgs;              // >= 1 (sgs = min(gsize, 32))
  const float* base = partials + (shared_partials ? 0 : static_cast<size_t>(n) * rows * nsg_total * 2);
  double s = 0.0, q = 0.0;
  // four independent (sum, sum of squares) loads in flight per thread, added in the original order (the loop was a
  // chain of L2-latency-bound trips: 21 of them for the 2664 tile rows of a 592x576 layer)
  const int total = rows * sg_per_group;
  const int bd = blockDim.x;
  auto part = [&](int i) {
    const int r = sg_per_group == 1 ? i : i / sg_per_group, k = sg_per_group == 1 ? 0 : i - r * sg_per_group;
    return __ldg(reinterpret_cast<const float2*>(base + (static_cast<size_t>(r) * nsg_total + g * sg_per_group + k) * 2));
  };
  int i = threadIdx.x;
  for (; i + 3 * bd < total; i += 4 * bd) {
    const float2 v0 = part(i), v1 = part(i + bd), v2 = part(i + 2 * bd), v3 = part(i + 3 * bd);
    s += static_cast<double>(v0.x); q += static_cast<double>(v0.y);
    s += static_cast<double>(v1.x); q += static_cast<double>(v1.y);
    s += static_cast<double>(v2.x); q += static_cast<double>(v2.y);
    s += static_cast<double>(v3.x); q += static_cast<double>(v3.y);
  }
  for (; i < total; i += bd) {
    const float2 v = part(i);
    s += static_cast<double>(v.x);
    q += static_cast<double>(v.y);
  }
  // fixed-shape tree over the block (128 or 512 threads; the thread -> row assignment and the tree depend on the block
  // size only, which the launcher derives from the row count alone: per-image results never depend on the batch)
  __shared__ double sh[2][512];
  sh[0][threadIdx.x] = s;
  sh[1][threadIdx.x] = q;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  const double mean = sh[0][0] / count;
  double var = sh[1][0] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  float scale = 1.f;
  if (keep) scale = static_cast<float>(numel_per_call / static_cast<double>(keep[n / images_per_call]));
  if (mean_rstd && threadIdx.x == 0) mean_rstd[static_cast<size_t>(n) * num_groups + g] = make_float2(static_cast<float>(mean), rstd);
  for (int i = threadIdx.x; i < gsize; i += blockDim.x) {
    const int ch = g * gsize + i;
    const float a = gamma[ch] * rstd;
    const float b = beta[ch] - static_cast<float>(mean) * a;
    coef[static_cast<size_t>(n) * c + ch] = make_float2(a * scale, b * scale);
  }
}

// ---------------------------------------------------------------------------------------------
struct ApplyParams {
  int n, h, w, c;
  int relu;
  int out_cstride, out_coffset;
  int mask2_cstride, mask2_coffset;
  int images_per_call2;
  double numel_per_call2;
  int x_shared;                 // 1: every image reads the input tensor of image 0 (Monte-Carlo: same image, same first conv)
};

struct Coef8 {
  float a[8], b[8];
  __device__ __forceinline__ void load(const float2* __restrict__ cf) {   // 8 consecutive channels = 64 B
    const float4* p = reinterpret_cast<const float4*>(cf);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = __ldg(p + i);
      a[2 * i] = t.x; b[2 * i] = t.y; a[2 * i + 1] = t.z; b[2 * i + 1] = t.w;
    }
  }
};

__device__ __forceinline__ void apply8(float (&f)[8], const Coef8& cf, uint32_t m1, bool relu) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float v = fmaf(f[i], cf.a[i], cf.b[i]);
    v = ((m1 >> i) & 1u) ? v : 0.f;
    f[i] = relu ? fmaxf(v, 0.f) : v;
  }
}

// grid = (blocks per image, n); thread t owns channel vector t % cvs (coefficients live in registers) and
// streams kApplyUnroll pixels per trip with all loads issued before the first use.  Specialised by mask presence
// (M1 = this unit's DropBlock mask, M2 = the concat-site mask); per-image base pointers with the channel slice folded in
// and 32-bit in-image offsets (the launcher checks h*w*max(c, out_cstride, mask2_cstride) < 2^31) keep the index
// arithmetic to a handful of instructions per vector -- the kernel shares the SM with the mask build, so every issue
// slot it does not need is one the Philox warps get.
constexpr int kApplyUnroll = 4;
template <typename T, bool M1, bool M2>
__global__ void __launch_bounds__(256, 4) gn_apply_kernel(const T* __restrict__ x, const float2* __restrict__ coef, const uint8_t* __restrict__ mask1,
                                const uint8_t* __restrict__ mask2, const unsigned long long* __restrict__ keep2,
                                T* __restrict__ out, ApplyParams p) {
  pdl_wait();       // predecessor grid complete + visible (no-op without the PDL launch attribute)
  pdl_trigger();    // the successor may be scheduled once every CTA got here
  const int n = blockIdx.y;
  const int cvs = p.c >> 3;
  const int cv = threadIdx.x % cvs;
  const int slot = threadIdx.x / cvs;
  const int slots = blockDim.x / cvs;
  const int hw = p.h * p.w;
  Coef8 cf;
  cf.load(coef + static_cast<size_t>(n) * p.c + cv * 8);
  float s2 = 1.f;
  if (M2) s2 = static_cast<float>(p.numel_per_call2 / static_cast<double>(keep2[n / p.images_per_call2]));
  const bool relu = p.relu != 0;
  const size_t img0 = static_cast<size_t>(n) * hw;
  const int m2s = p.mask2_cstride >> 3, m2o = (p.mask2_coffset >> 3) + cv;
  const T* const xb = x + (p.x_shared ? 0 : img0 * p.c) + cv * 8;
  const uint8_t* const m1b = M1 ? mask1 + img0 * cvs + cv : nullptr;
  const uint8_t* const m2b = M2 ? mask2 + img0 * m2s + m2o : nullptr;
  T* const ob = out + img0 * p.out_cstride + p.out_coffset + cv * 8;
  const int stride = gridDim.x * slots * kApplyUnroll;
  for (int base = blockIdx.x * slots * kApplyUnroll + slot; base < hw; base += stride) {
    Vec8<T> vec[kApplyUnroll];
    uint32_t m1[kApplyUnroll], m2[kApplyUnroll];
#pragma unroll
    for (int u = 0; u < kApplyUnroll; ++u) {
      const int pl = base + u * slots;
      if (pl < hw) {
        vec[u].load(xb + static_cast<uint32_t>(pl) * static_cast<uint32_t>(p.c));
        m1[u] = M1 ? m1b[static_cast<uint32_t>(pl) * static_cast<uint32_t>(cvs)] : 0xFFu;
        m2[u] = M2 ? m2b[static_cast<uint32_t>(pl) * static_cast<uint32_t>(m2s)] : 0xFFu;
      }
    }
#pragma unroll
    for (int u = 0; u < kApplyUnroll; ++u) {
      const int pl = base + u * slots;
      if (pl < hw) {
        float f[8];
        vec[u].to_float(f);
        apply8(f, cf, m1[u], relu);
        if (M2) {
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = ((m2[u] >> i) & 1u) ? f[i] * s2 : 0.f;
        }
        round_for_storage<T>(f);
        Vec8<T> o;
        o.from_float(f);
        o.store(ob + static_cast<uint32_t>(pl) * static_cast<uint32_t>(p.out_cstride));
      }
    }
  }
}

// Encoder tail: apply + skip store (with concat mask) + 2x2 max-pool + pooled GroupNorm partials.
constexpr int kPoolUnroll = 2;
// ATen's max-pool update rule `val > maxval || isnan(val)` without the window index = a NaN-propagating maximum
__device__ __forceinline__ float max_nan(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
// ARGMAX: also record the window index of the maximum (training: the backward scatters through it); the inference
// instantiation drops the index bookkeeping (a third of the kernel's instructions).
// MODE 1: both DropBlock masks and the skip store present (Monte-Carlo / training), 0: neither mask (eval forward),
// 2: decided at run time (any other combination) -- the specialised modes are free of per-tap null checks.
template <typename T, bool ARGMAX, int MODE>
__global__ void __launch_bounds__(256, 2) gn_apply_pool_kernel(const T* __restrict__ x, const float2* __restrict__ coef, const uint8_t* __restrict__ mask1,
                                     const uint8_t* __restrict__ mask2, const unsigned long long* __restrict__ keep2,
                                     T* __restrict__ skip_out, T* __restrict__ pooled, float* __restrict__ pool_partials,
                                     uint8_t* __restrict__ argmax, int pool_sgs, ApplyParams p) {
  pdl_wait();       // predecessor grid complete + visible (no-op without the PDL launch attribute)
  pdl_trigger();    // the successor may be scheduled once every CTA got here
  extern __shared__ float sm[];
  const int n = blockIdx.y;
  const int cvs = p.c >> 3;
  const int cv = threadIdx.x % cvs;
  const int slot = threadIdx.x / cvs;
  const int slots = blockDim.x / cvs;
  const int ph = p.h >> 1, pw = p.w >> 1;
  const int npool = ph * pw;
  Coef8 cf;
  cf.load(coef + static_cast<size_t>(n) * p.c + cv * 8);
  float s2 = 1.f;
  if (mask2) s2 = static_cast<float>(p.numel_per_call2 / static_cast<double>(keep2[n / p.images_per_call2]));
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  const bool relu = p.relu != 0;
  const int m2s = p.mask2_cstride >> 3, m2o = (p.mask2_coffset >> 3) + cv;
  constexpr int U = kPoolUnroll;                        // pooled pixels per trip: 8 x 16 B loads in flight per thread
  // Addressing: 64-bit base pointers per image (channel slice folded in), 32-bit element offsets inside the image
  // (the launcher checks h*w*max(c, out_cstride) < 2^31), window taps at constant offsets, and the pooled (row, col)
  // advanced incrementally instead of divided out per trip -- the first version of this loop spent two thirds of its
  // ~225 instructions per 16-byte vector on 64-bit index arithmetic.
  const size_t img_pix = static_cast<size_t>(n) * p.h * p.w;
  const T* const xb = x + img_pix * p.c + cv * 8;
  const bool has1 = MODE == 2 ? mask1 != nullptr : MODE == 1;
  const bool has2 = MODE == 2 ? mask2 != nullptr : MODE == 1;
  const bool has_skip = MODE == 2 ? skip_out != nullptr : true;
  const uint8_t* const m1b = mask1 + img_pix * cvs + cv;
  const uint8_t* const m2b = mask2 + img_pix * m2s + m2o;
  T* const sb = skip_out + img_pix * p.out_cstride + p.out_coffset + cv * 8;
  T* const pb = pooled + static_cast<size_t>(n) * npool * p.c + cv * 8;
  uint8_t* const ab = ARGMAX ? argmax + static_cast<size_t>(n) * npool * p.c + cv * 8 : nullptr;
  const uint32_t tap_x[4] = {0u, static_cast<uint32_t>(p.c), static_cast<uint32_t>(p.w) * p.c, static_cast<uint32_t>(p.w) * p.c + p.c};
  const uint32_t tap_m1[4] = {0u, static_cast<uint32_t>(cvs), static_cast<uint32_t>(p.w) * cvs, static_cast<uint32_t>(p.w) * cvs + cvs};
  const uint32_t tap_m2[4] = {0u, static_cast<uint32_t>(m2s), static_cast<uint32_t>(p.w) * m2s, static_cast<uint32_t>(p.w) * m2s + m2s};
  const uint32_t tap_s[4] = {0u, static_cast<uint32_t>(p.out_cstride), static_cast<uint32_t>(p.w) * p.out_cstride,
                             static_cast<uint32_t>(p.w) * p.out_cstride + p.out_cstride};
  const int stride = gridDim.x * slots * U;
  const int dq = stride / pw, dr = stride - dq * pw;    // (row, col) step of one trip
  int py[U], px[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int pp0 = blockIdx.x * slots * U + slot + u * slots;
    py[u] = pp0 / pw;
    px[u] = pp0 - py[u] * pw;
  }
  for (int base = blockIdx.x * slots * U + slot; base < npool; base += stride) {
    Vec8<T> vec[U][4];
    uint32_t m1[U][4], m2[U][4];
    uint32_t q00[U];                                    // in-image pixel index of the window's top-left tap
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pp = base + u * slots;
      q00[u] = static_cast<uint32_t>(2 * py[u]) * p.w + 2 * px[u];
      if (pp < npool) {
        const uint32_t ox = q00[u] * p.c, om1 = q00[u] * cvs, om2 = q00[u] * m2s;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          vec[u][k].load(xb + (ox + tap_x[k]));
          m1[u][k] = has1 ? m1b[om1 + tap_m1[k]] : 0xFFu;
          m2[u][k] = has2 ? m2b[om2 + tap_m2[k]] : 0xFFu;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pp = base + u * slots;
      if (pp < npool) {
        float best[8];
        uint32_t arg_lo = 0u, arg_hi = 0u;                 // window index (0..3) of the maximum, one byte per channel
        const uint32_t os = q00[u] * p.out_cstride;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float f[8];
          vec[u][k].to_float(f);
          apply8(f, cf, m1[u][k], relu);
          if (ARGMAX) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              // ATen max_pool2d: first maximum in row-major window order wins (val > maxval || isnan(val))
              if (k == 0 || f[i] > best[i] || f[i] != f[i]) {
                best[i] = f[i];
                if (k > 0) {
                  uint32_t& a = i < 4 ? arg_lo : arg_hi;
                  const int sh = 8 * (i & 3);
                  a = (a & ~(0xFFu << sh)) | (static_cast<uint32_t>(k) << sh);
                }
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) best[i] = k == 0 ? f[i] : max_nan(best[i], f[i]);
          }
          if (has_skip) {
            float g[8];
            if (has2) {
#pragma unroll
              for (int i = 0; i < 8; ++i) g[i] = ((m2[u][k] >> i) & 1u) ? f[i] * s2 : 0.f;
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) g[i] = f[i];
            }
            round_for_storage<T>(g);
            Vec8<T> o;
            o.from_float(g);
            o.store(sb + (os + tap_s[k]));
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s[i] += best[i];
          q[i] += best[i] * best[i];
        }
        const uint32_t po = static_cast<uint32_t>(pp) * p.c;
        Vec8<T> o;
        o.from_float(best);
        o.store(pb + po);
        if (ARGMAX) *reinterpret_cast<uint2*>(ab + po) = make_uint2(arg_lo, arg_hi);
      }
      py[u] += dq;
      px[u] += dr;
      if (px[u] >= pw) {
        px[u] -= pw;
        ++py[u];
      }
    }
  }
  if (pool_partials) {
    block_stats_to_partials(s, q, p.c, pool_sgs, sm,
                            pool_partials + (static_cast<size_t>(n) * gridDim.x + blockIdx.x) * (p.c / pool_sgs) * 2);
  }
}

// ---------------------------------------------------------------------------------------------
// Output head.  LPP = c/8 lanes cooperate on one pixel (c <= 256); each lane owns 8 channels.
constexpr int kHeadPix = 2;
template <typename T>
__global__ void __launch_bounds__(256, 3) head_kernel(const T* __restrict__ x, const float2* __restrict__ coef, const uint8_t* __restrict__ mask1,
                            const float* __restrict__ w_head, float* __restrict__ out, float* __restrict__ logits,
                            const float* __restrict__ fov, double* __restrict__ acc, float* __restrict__ samples,
                            const long long* __restrict__ iter_base, b2u_head_desc d) {
  pdl_wait();       // predecessor grid complete + visible (no-op without the PDL launch attribute)
  pdl_trigger();    // the successor may be scheduled once every CTA got here
  const int lpp = d.c >> 3;
  const int lane_in = threadIdx.x % lpp;
  const long group = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) / lpp;
  const long ngroups = (static_cast<long>(gridDim.x) * blockDim.x) / lpp;
  const long npix0 = static_cast<long>(d.h0) * d.w0;
  const long img_stride = static_cast<long>(d.h) * d.w;
  float wh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) wh[i] = __ldg(w_head + lane_in * 8 + i);
  const long long base_iter = iter_base ? *iter_base : 0;
  // every lane of a warp runs the same trip count (shuffles below need the full warp).  A lane group walks
  // kHeadPix pixels per trip; the loads of image n+1 are issued before image n is reduced (register double
  // buffer), so 2 * kHeadPix 16-byte loads per thread stay in flight across the shuffle / exp / fp64 section.
  constexpr int P = kHeadPix;
  const long trips = (npix0 + ngroups * P - 1) / (ngroups * P);
  for (long tr = 0; tr < trips; ++tr) {
    long op[P];
    bool active[P];
    long pixbase[P];
    double s1[P], s2[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
      op[j] = (tr * P + j) * ngroups + group;
      active[j] = op[j] < npix0;
      const long opc = active[j] ? op[j] : 0;
      const int oh = static_cast<int>(opc / d.w0), ow = static_cast<int>(opc - static_cast<long>(oh) * d.w0);
      pixbase[j] = static_cast<long>(oh) * d.w + ow;
      s1[j] = s2[j] = 0.0;
    }
    Vec8<T> cur[P], nxt[P];
    uint32_t mcur[P], mnxt[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
      cur[j].load(x + pixbase[j] * d.c + lane_in * 8);
      mcur[j] = mask1 ? mask1[pixbase[j] * lpp + lane_in] : 0xFFu;
    }
    for (int n = 0; n < d.n; ++n) {
      if (n + 1 < d.n) {
#pragma unroll
        for (int j = 0; j < P; ++j) {
          const long pix = static_cast<long>(n + 1) * img_stride + pixbase[j];
          nxt[j].load(x + pix * d.c + lane_in * 8);
          mnxt[j] = mask1 ? mask1[pix * lpp + lane_in] : 0xFFu;
        }
      }
      Coef8 cf;
      cf.load(coef + static_cast<size_t>(n) * d.c + lane_in * 8);
#pragma unroll
      for (int j = 0; j < P; ++j) {
        float f[8];
        cur[j].to_float(f);
        apply8(f, cf, mcur[j], true);
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) dot = fmaf(f[i], wh[i], dot);
        for (int o = lpp >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        if (active[j] && lane_in == 0) {
          float yv = 1.f / (1.f + expf(-dot));
          yv = fminf(fmaxf(yv, 0.f), 1.f);
          if (yv != yv) yv = 0.f;
          if (logits) logits[static_cast<long>(n) * npix0 + op[j]] = dot;
          if (out) out[static_cast<long>(n) * npix0 + op[j]] = yv;
          if (acc) {
            const float fv = fov ? fov[(d.fov_per_image ? static_cast<long>(n) * npix0 : 0) + op[j]] : 1.f;
            const float v = yv * fv;
            s1[j] += static_cast<double>(v);
            s2[j] += static_cast<double>(v) * static_cast<double>(v);
            const long long it = base_iter + n;
            if (samples && it < d.return_num) samples[it * npix0 + op[j]] = v;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < P; ++j) {
        cur[j] = nxt[j];
        mcur[j] = mnxt[j];
      }
    }
#pragma unroll
    for (int j = 0; j < P; ++j) {
      if (acc && active[j] && lane_in == 0) {
        acc[op[j]] += s1[j];
        acc[npix0 + op[j]] += s2[j];
      }
    }
  }
}

// Fast path of the head for c = 64 and 16-bit activations (the canonical U-Net): a group of 8 lanes owns 8 CONSECUTIVE
// output pixels per trip.  Lane l loads channel slice l of all 8 pixels (one pixel = the group's 128 contiguous bytes),
// the 8 partial dot products per lane are reduced with a transposing butterfly (7 shuffles for 8 pixels instead of 24;
// same summation tree as the generic kernel, so the logits are bit-identical), and lane l then runs the sigmoid /
// clamp / fov / fp64-accumulate tail for pixel l -- every lane active, stores of 32 consecutive floats per warp --
// instead of one lane in eight.  The generic kernel above spends ~100 issue slots per 16-byte vector and is issue
// bound at 29 % of HBM bandwidth; this one needs ~45.
template <typename T>
__global__ void __launch_bounds__(256, 2) head8_kernel(const T* __restrict__ x, const float2* __restrict__ coef, const uint8_t* __restrict__ mask1,
                             const float* __restrict__ w_head, float* __restrict__ out, float* __restrict__ logits,
                             const float* __restrict__ fov, double* __restrict__ acc, float* __restrict__ samples,
                             const long long* __restrict__ iter_base, b2u_head_desc d, int trips) {
  pdl_wait();       // predecessor grid complete + visible (no-op without the PDL launch attribute)
  pdl_trigger();    // the successor may be scheduled once every CTA got here
  const int l = threadIdx.x & 7;
  const int gbase = threadIdx.x & 24;                         // first lane of this group within the warp
  const long group = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 3;
  const long ngroups = gridDim.x * static_cast<long>(blockDim.x >> 3);
  const long npix0 = static_cast<long>(d.h0) * d.w0;
  const long img_stride = static_cast<long>(d.h) * d.w;
  float wh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) wh[i] = __ldg(w_head + l * 8 + i);
  const long long base_iter = iter_base ? *iter_base : 0;
  const bool h4 = (l & 4) != 0, h2 = (l & 2) != 0, h1 = (l & 1) != 0;
  for (int tr = 0; tr < trips; ++tr) {
    const long op = (tr * ngroups + group) * 8 + l;           // this lane's own output pixel
    const bool active = op < npix0;
    const long opc = active ? op : npix0 - 1;
    const int oh = static_cast<int>(opc / d.w0), ow = static_cast<int>(opc - static_cast<long>(oh) * d.w0);
    const int mypb = oh * d.w + ow;                           // pixel index inside the padded image
    int pb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) pb[j] = __shfl_sync(0xffffffffu, mypb, gbase + j);
    double s1 = 0.0, s2 = 0.0, a1 = 0.0, a2 = 0.0;
    float fv0 = 1.f;
    if (acc && active) {                                      // fetched a whole trip before their use
      a1 = acc[op];
      a2 = acc[npix0 + op];
      if (fov && !d.fov_per_image) fv0 = fov[op];
    }
    for (int n = 0; n < d.n; ++n) {
      Vec8<T> vec[8];
      uint32_t m1[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const long pix = static_cast<long>(n) * img_stride + pb[j];
        vec[j].load(x + pix * 64 + l * 8);
        m1[j] = mask1 ? mask1[pix * 8 + l] : 0xFFu;
      }
      Coef8 cf;
      cf.load(coef + static_cast<size_t>(n) * 64 + l * 8);
      float ds[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float f[8];
        vec[j].to_float(f);
        apply8(f, cf, m1[j], true);
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) dot = fmaf(f[i], wh[i], dot);
        ds[j] = dot;
      }
      // transposing butterfly: after the xor-4 step a lane keeps the 4 pixels of its half, then 2, then its own
      float e4[4], e2[2];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float send = h4 ? ds[k] : ds[k + 4], keep = h4 ? ds[k + 4] : ds[k];
        e4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float send = h2 ? e4[k] : e4[k + 2], keep = h2 ? e4[k + 2] : e4[k];
        e2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      }
      const float send1 = h1 ? e2[0] : e2[1], keep1 = h1 ? e2[1] : e2[0];
      const float dot = keep1 + __shfl_xor_sync(0xffffffffu, send1, 1);
      if (active) {
        float yv = 1.f / (1.f + expf(-dot));
        yv = fminf(fmaxf(yv, 0.f), 1.f);
        if (yv != yv) yv = 0.f;
        if (logits) logits[static_cast<long>(n) * npix0 + op] = dot;
        if (out) out[static_cast<long>(n) * npix0 + op] = yv;
        if (acc) {
          const float fv = (fov && d.fov_per_image) ? fov[static_cast<long>(n) * npix0 + op] : fv0;
          const float v = yv * fv;
          s1 += static_cast<double>(v);
          s2 += static_cast<double>(v) * static_cast<double>(v);
          const long long it = base_iter + n;
          if (samples && it < d.return_num) samples[it * npix0 + op] = v;
        }
      }
    }
    if (acc && active) {
      acc[op] = a1 + s1;
      acc[npix0 + op] = a2 + s2;
    }
  }
}

// head8 with the 16-byte activation vectors (and the 8 mask bytes of a pixel) moved by cp.async through a per-thread
// shared-memory ring of S items (item = one image of one trip: 8 vectors per thread): the loads of the next S - 1 items
// are in flight WHILE an item is reduced, at no register cost (head8_kernel holds its 8 vectors in registers and has
// nothing in flight during the butterfly / sigmoid / fp64 tail -- 128 registers leave no room for a second set).  A
// thread reads back only the vectors it copied itself (no barrier); the mask bytes are copied by the pixel's own lane
// (8 bytes) and read by the eight lanes of its group after a __syncwarp.  Same arithmetic in the same order as head8_kernel.
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t saddr, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, int S>
__global__ void __launch_bounds__(256, 2) head8_async_kernel(const T* __restrict__ x, const float2* __restrict__ coef, const uint8_t* __restrict__ mask1,
                             const float* __restrict__ w_head, float* __restrict__ out, float* __restrict__ logits,
                             const float* __restrict__ fov, double* __restrict__ acc, float* __restrict__ samples,
                             const long long* __restrict__ iter_base, b2u_head_desc d, int trips) {
  extern __shared__ __align__(16) uint8_t head_ring[];        // [S][8][nt] uint4, then [S][nt] 8-byte pixel masks
  pdl_wait();       // predecessor grid complete + visible (no-op without the PDL launch attribute)
  pdl_trigger();    // the successor may be scheduled once every CTA got here
  const int nt = blockDim.x;
  const int l = threadIdx.x & 7;
  const int gbase = threadIdx.x & 24;                         // first lane of this group within the warp
  const long group = (blockIdx.x * static_cast<long>(nt) + threadIdx.x) >> 3;
  const long ngroups = gridDim.x * static_cast<long>(nt >> 3);
  const long npix0 = static_cast<long>(d.h0) * d.w0;
  const long img_stride = static_cast<long>(d.h) * d.w;
  const uint32_t ring = smem_u32(head_ring) + threadIdx.x * 16u;
  const uint32_t mring = smem_u32(head_ring) + static_cast<uint32_t>(S * 8 * nt) * 16u;
  const uint32_t stage_bytes = static_cast<uint32_t>(8 * nt) * 16u;
  const uint32_t mstage_bytes = static_cast<uint32_t>(nt) * 8u;
  float wh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) wh[i] = __ldg(w_head + l * 8 + i);
  const long long base_iter = iter_base ? *iter_base : 0;
  const bool h4 = (l & 4) != 0, h2 = (l & 2) != 0, h1 = (l & 1) != 0;
  const int items = trips * d.n;

  // producer cursor: item (p_tr, p_n) goes to stage p_s; p_pb = this lane's own pixel of trip p_tr (clamped)
  int p_tr = 0, p_n = 0, p_s = 0, p_pb = 0;
  auto own_pixel = [&](int tr, long& op, bool& active) {
    op = (tr * ngroups + group) * 8 + l;
    active = op < npix0;
    const long opc = active ? op : npix0 - 1;
    const int oh = static_cast<int>(opc / d.w0), ow = static_cast<int>(opc - static_cast<long>(oh) * d.w0);
    return oh * d.w + ow;
  };
  {
    long op_; bool a_;
    p_pb = own_pixel(0, op_, a_);
  }
  auto issue = [&]() {
    if (p_tr < trips) {
      const long img_off = static_cast<long>(p_n) * img_stride;
      const uint32_t sdst = ring + static_cast<uint32_t>(p_s) * stage_bytes;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int pbj = __shfl_sync(0xffffffffu, p_pb, gbase + j);
        cp_async16(sdst + static_cast<uint32_t>(j * nt) * 16u, x + (img_off + pbj) * 64 + l * 8);
      }
      if (mask1) cp_async8(mring + static_cast<uint32_t>(p_s) * mstage_bytes + threadIdx.x * 8u, mask1 + (img_off + p_pb) * 8);
      if (++p_n == d.n) {
        p_n = 0;
        ++p_tr;
        long op_; bool a_;
        p_pb = own_pixel(p_tr < trips ? p_tr : 0, op_, a_);
      }
      if (++p_s == S) p_s = 0;
    }
    cp_async_commit();                                          // (an empty group past the end keeps the wait count uniform)
  };
#pragma unroll
  for (int s = 0; s < S - 1; ++s) issue();

  int c_n = 0, c_tr = 0, c_s = 0;
  long op = 0;
  bool active = false;
  double s1 = 0.0, s2 = 0.0, a1 = 0.0, a2 = 0.0;
  float fv0 = 1.f;
  for (int k = 0; k < items; ++k) {
    cp_async_wait<S - 2>();                                     // item k has landed (this thread's copies)
    __syncwarp();                                               // ... and the other lanes' mask bytes; stage (k - 1) % S is free
    issue();                                                    // item k + S - 1 into the stage consumed last iteration
    if (c_n == 0) {
      // a trip's own-pixel state; the fov value and the accumulator pair are fetched HERE, a whole trip before their use
      // (they were two exposed DRAM round trips per trip: every warp of the SM reaches them at about the same time)
      own_pixel(c_tr, op, active);
      s1 = 0.0; s2 = 0.0;
      if (acc && active) {
        a1 = acc[op];
        a2 = acc[npix0 + op];
        if (fov && !d.fov_per_image) fv0 = fov[op];
      }
    }
    const int n = c_n;
    Coef8 cf;
    cf.load(coef + static_cast<size_t>(n) * 64 + l * 8);
    const uint32_t ssrc = ring + static_cast<uint32_t>(c_s) * stage_bytes;
    const uint32_t msrc = mring + static_cast<uint32_t>(c_s) * mstage_bytes + static_cast<uint32_t>(threadIdx.x & ~7) * 8u + l;
    float ds[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      Vec8<T> v;
      v.raw = lds128(ssrc + static_cast<uint32_t>(j * nt) * 16u);
      uint32_t m1 = 0xFFu;
      if (mask1) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(m1) : "r"(msrc + j * 8u));
      float f[8];
      v.to_float(f);
      apply8(f, cf, m1, true);
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) dot = fmaf(f[i], wh[i], dot);
      ds[j] = dot;
    }
    float e4[4], e2[2];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float send = h4 ? ds[kk] : ds[kk + 4], keep = h4 ? ds[kk + 4] : ds[kk];
      e4[kk] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const float send = h2 ? e4[kk] : e4[kk + 2], keep = h2 ? e4[kk + 2] : e4[kk];
      e2[kk] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    const float send1 = h1 ? e2[0] : e2[1], keep1 = h1 ? e2[1] : e2[0];
    const float dot = keep1 + __shfl_xor_sync(0xffffffffu, send1, 1);
    if (active) {
      float yv = 1.f / (1.f + expf(-dot));
      yv = fminf(fmaxf(yv, 0.f), 1.f);
      if (yv != yv) yv = 0.f;
      if (logits) logits[static_cast<long>(n) * npix0 + op] = dot;
      if (out) out[static_cast<long>(n) * npix0 + op] = yv;
      if (acc) {
        const float fv = (fov && d.fov_per_image) ? fov[static_cast<long>(n) * npix0 + op] : fv0;
        const float v = yv * fv;
        s1 += static_cast<double>(v);
        s2 += static_cast<double>(v) * static_cast<double>(v);
        const long long it = base_iter + n;
        if (samples && it < d.return_num) samples[it * npix0 + op] = v;
      }
    }
    if (++c_n == d.n) {
      c_n = 0;
      ++c_tr;
      if (acc && active) {
        acc[op] = a1 + s1;
        acc[npix0 + op] = a2 + s2;
      }
    }
    if (++c_s == S) c_s = 0;
  }
  cp_async_wait<0>();
}

__global__ void mc_finalize_kernel(const double* __restrict__ acc, float* __restrict__ mean, float* __restrict__ stdv,
                                   long long npix, long long t) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < npix;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const double s1 = acc[i], s2 = acc[npix + i];
    const double m = s1 / static_cast<double>(t);
    double var = (s2 - s1 * s1 / static_cast<double>(t)) / static_cast<double>(t - 1);
    if (var < 0.0) var = 0.0;
    mean[i] = static_cast<float>(m);
    stdv[i] = static_cast<float>(sqrt(var));
  }
}

__global__ void mc_accumulate_kernel(const float* __restrict__ x, const float* __restrict__ fov, double* __restrict__ acc,
                                     float* __restrict__ samples, const long long* __restrict__ iter_base, int n,
                                     long long npix, int return_num) {
  const long long base_iter = iter_base ? *iter_base : 0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < npix;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float fv = fov ? fov[i] : 1.f;
    double s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < n; ++k) {
      const float v = x[static_cast<long long>(k) * npix + i] * fv;
      s1 += static_cast<double>(v);
      s2 += static_cast<double>(v) * static_cast<double>(v);
      const long long it = base_iter + k;
      if (samples && it < return_num) samples[it * npix + i] = v;
    }
    acc[i] += s1;
    acc[npix + i] += s2;
  }
}

__global__ void advance_counter_kernel(long long* c, long long delta) {
  pdl_wait();       // predecessor grid complete + visible (no-op without the PDL launch attribute)
  pdl_trigger();    // the successor may be scheduled once every CTA got here
  *c += delta;
}

// Confusion counts of a probability map against a binary ground truth inside the field-of-view mask
// (utils_metrics.py:157-173: masked round() -> sklearn f1 / accuracy): counts[0..3] = TP, FP, FN, TN.
// pred = rint(seg) == 1 (numpy rounds half to even: 0.5 -> 0), truth = (long)gt == 1, pixel used iff (long)mask != 0.
__global__ void __launch_bounds__(256) confusion_kernel(const float* __restrict__ seg, const float* __restrict__ gt, const float* __restrict__ mask,
                                                        long long n, unsigned long long* __restrict__ counts) {
  unsigned int c[4] = {0u, 0u, 0u, 0u};
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
    if (static_cast<long long>(mask[i]) != 0) {
      const bool p = rintf(seg[i]) == 1.f;
      const bool t = static_cast<long long>(gt[i]) == 1;
      c[p ? (t ? 0 : 1) : (t ? 2 : 3)]++;
    }
  }
  __shared__ unsigned int sh[4][8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    unsigned int v = c[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    unsigned long long tot = 0;
    for (int w = 0; w < 8; ++w) tot += sh[threadIdx.x][w];
    if (tot) atomicAdd(counts + threadIdx.x, tot);      // integer atomics: order-independent, deterministic
  }
}

// ---------------------------------------------------------------------------------------------
// torchvision TF.rotate(BILINEAR, fill=0) restated in one gather kernel (see include/b2u.h).
struct RotParams {
  float t00, t01, t02, t10, t11, t12;   // rescaled theta^T columns: gx = x*t00 + y*t01 + t02, gy = x*t10 + y*t11 + t12
};
__device__ __forceinline__ float linspace_val(float start, float end, int steps, int i) {
  // ATen linspace (RangeFactories.cu): symmetric evaluation around the midpoint
  if (steps == 1) return start;
  const float step = (end - start) / static_cast<float>(steps - 1);
  const int half = steps / 2;
  return i < half ? start + step * i : end - step * (steps - i - 1);
}
constexpr int kMaxAnglesPerLaunch = 64;
struct RotBatch {
  RotParams r[kMaxAnglesPerLaunch];
};
__global__ void rotate_kernel(const float* __restrict__ x, float* __restrict__ out, int n, int c, int h, int w,
                              const RotBatch rp, long x_batch_stride) {
  const long total = static_cast<long>(n) * h * w;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int ow = static_cast<int>(i % w);
    const int oh = static_cast<int>((i / w) % h);
    const int img = static_cast<int>(i / (static_cast<long>(w) * h));
    const RotParams r = rp.r[img];
    const float bx = linspace_val(-w * 0.5f + 0.5f, w * 0.5f + 0.5f - 1.f, w, ow);
    const float by = linspace_val(-h * 0.5f + 0.5f, h * 0.5f + 0.5f - 1.f, h, oh);
    const float gx = bx * r.t00 + by * r.t01 + r.t02;
    const float gy = bx * r.t10 + by * r.t11 + r.t12;
    // grid_sample, align_corners=False: unnormalise, bilinear, zero padding
    const float ix = ((gx + 1.f) * w - 1.f) * 0.5f;
    const float iy = ((gy + 1.f) * h - 1.f) * 0.5f;
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
    const float ax = ix - fx, ay = iy - fy;
    const float wnw = (1.f - ax) * (1.f - ay), wne = ax * (1.f - ay), wsw = (1.f - ax) * ay, wse = ax * ay;
    const bool inx0 = x0 >= 0 && x0 < w, inx1 = x0 + 1 >= 0 && x0 + 1 < w;
    const bool iny0 = y0 >= 0 && y0 < h, iny1 = y0 + 1 >= 0 && y0 + 1 < h;
    float m = 0.f;
    if (inx0 && iny0) m += wnw;
    if (inx1 && iny0) m += wne;
    if (inx0 && iny1) m += wsw;
    if (inx1 && iny1) m += wse;
    for (int ch = 0; ch < c; ++ch) {
      const float* src = x + img * x_batch_stride + static_cast<long>(ch) * h * w;
      float v = 0.f;
      if (inx0 && iny0) v += src[static_cast<long>(y0) * w + x0] * wnw;
      if (inx1 && iny0) v += src[static_cast<long>(y0) * w + x0 + 1] * wne;
      if (inx0 && iny1) v += src[static_cast<long>(y0 + 1) * w + x0] * wsw;
      if (inx1 && iny1) v += src[static_cast<long>(y0 + 1) * w + x0 + 1] * wse;
      out[(static_cast<long>(img) * c + ch) * h * w + static_cast<long>(oh) * w + ow] = v * m;
    }
  }
}

// Bilinear sample of one rotated pixel (shared by the three rotation kernels): returns the blend weight m of the
// ones channel and the four taps (offset, weight); out-of-image taps get weight 0 and offset 0.
struct RotTaps {
  int o[4];
  float wgt[4];
  float m;
};
__device__ __forceinline__ RotTaps rot_taps(const RotParams& r, int oh, int ow, int h, int w) {
  const float bx = linspace_val(-w * 0.5f + 0.5f, w * 0.5f + 0.5f - 1.f, w, ow);
  const float by = linspace_val(-h * 0.5f + 0.5f, h * 0.5f + 0.5f - 1.f, h, oh);
  const float gx = bx * r.t00 + by * r.t01 + r.t02;
  const float gy = bx * r.t10 + by * r.t11 + r.t12;
  const float ix = ((gx + 1.f) * w - 1.f) * 0.5f;
  const float iy = ((gy + 1.f) * h - 1.f) * 0.5f;
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
  const float ax = ix - fx, ay = iy - fy;
  const bool inx0 = x0 >= 0 && x0 < w, inx1 = x0 + 1 >= 0 && x0 + 1 < w;
  const bool iny0 = y0 >= 0 && y0 < h, iny1 = y0 + 1 >= 0 && y0 + 1 < h;
  RotTaps t;
  t.wgt[0] = (inx0 && iny0) ? (1.f - ax) * (1.f - ay) : 0.f;
  t.wgt[1] = (inx1 && iny0) ? ax * (1.f - ay) : 0.f;
  t.wgt[2] = (inx0 && iny1) ? (1.f - ax) * ay : 0.f;
  t.wgt[3] = (inx1 && iny1) ? ax * ay : 0.f;
  t.o[0] = (inx0 && iny0) ? y0 * w + x0 : 0;
  t.o[1] = (inx1 && iny0) ? y0 * w + x0 + 1 : 0;
  t.o[2] = (inx0 && iny1) ? (y0 + 1) * w + x0 : 0;
  t.o[3] = (inx1 && iny1) ? (y0 + 1) * w + x0 + 1 : 0;
  // same accumulation order as rotate_kernel / grid_sample: nw, ne, sw, se (skipped taps add nothing)
  float m = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) m += t.wgt[k];
  t.m = m;
  return t;
}

// Rotation ensemble, CUDA-graph form (Rotational_Uncertainty.py:51-60): the per-angle coefficients come from a DEVICE
// table indexed by the global angle index `*iter_base + k`, so one captured graph serves every step of the loop.
// rotate_in:  x [c][h][w] (one image) -> out [n][c][h][w], image k rotated by table[min(*iter_base + k, table_len - 1)].
__global__ void __launch_bounds__(256) rotate_in_table_kernel(const float* __restrict__ x, float* __restrict__ out, int n, int c, int h, int w,
                                                            const RotParams* __restrict__ table, int table_len,
                                                            const long long* __restrict__ iter_base) {
  const long long base = *iter_base;
  const int hw = h * w;
  const long total = static_cast<long>(n) * hw;
  for (long i = blockIdx.x * 256L + threadIdx.x; i < total; i += gridDim.x * 256L) {
    const int img = static_cast<int>(i / hw);
    const int pix = static_cast<int>(i - static_cast<long>(img) * hw);
    long long ti = base + img;
    if (ti > table_len - 1) ti = table_len - 1;
    const RotTaps t = rot_taps(table[ti], pix / w, pix % w, h, w);
    for (int ch = 0; ch < c; ++ch) {
      const float* src = x + static_cast<long>(ch) * hw;
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (t.wgt[k] != 0.f) v += src[t.o[k]] * t.wgt[k];
      out[(static_cast<long>(img) * c + ch) * hw + pix] = v * t.m;
    }
  }
}
// rotate_back + FOV mask + per-pixel fp64 (sum, sum of squares) + the first `return_num` samples in ONE pass
// (Rotational_Uncertainty.py:58-63): seg [n][h][w]; image k is angle index *iter_base + k and is skipped when that index
// is >= *iter_limit (the tail of the last batch).
__global__ void __launch_bounds__(256) rotate_back_accumulate_kernel(const float* __restrict__ seg, const float* __restrict__ fov,
                                                                   double* __restrict__ acc, float* __restrict__ samples, int n, int h,
                                                                   int w, int return_num, const RotParams* __restrict__ table,
                                                                   int table_len, const long long* __restrict__ iter_base,
                                                                   const long long* __restrict__ iter_limit) {
  const long long base = *iter_base, limit = *iter_limit;
  const int hw = h * w;
  for (int pix = blockIdx.x * 256 + threadIdx.x; pix < hw; pix += gridDim.x * 256) {
    const float fv = fov ? fov[pix] : 1.f;
    double s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < n; ++k) {
      const long long it = base + k;
      if (it >= limit || it >= table_len) break;
      const RotTaps t = rot_taps(table[it], pix / w, pix % w, h, w);
      const float* src = seg + static_cast<long>(k) * hw;
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (t.wgt[j] != 0.f) v += src[t.o[j]] * t.wgt[j];
      v = v * t.m * fv;
      s1 += static_cast<double>(v);
      s2 += static_cast<double>(v) * static_cast<double>(v);
      if (samples && it < return_num) samples[it * hw + pix] = v;
    }
    acc[pix] += s1;
    acc[hw + pix] += s2;
  }
}

static int grid_for(long work_items, int threads) {
  long blocks = (work_items + threads - 1) / threads;
  long cap = static_cast<long>(b2u_num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

static int stat_sgs(int c, int num_groups) {
  int gs = c / num_groups;
  return gs < 32 ? gs : 32;
}

static int blocks_per_image_for(long items, int per_block) {
  long b = (items + per_block - 1) / per_block;
  long cap = static_cast<long>(b2u_num_sms()) * 4;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

static int pick_threads(int cvs) {
  // blockDim must be a multiple of the channel-vector count so each thread keeps one cv
  int t = 256;
  if (cvs > 256) return 0;
  t = (256 / cvs) * cvs;
  return t;
}

}  // namespace b2u

using namespace b2u;

extern "C" int b2u_conv_first_stat_layout(int h, int w, int cout, int num_groups, int* rows_per_image,
                                          int* subgroup_size) {
  B2U_REQUIRE(h > 0 && w > 0 && cout > 0 && cout % 8 == 0 && cout <= 256, "bad shape h=%d w=%d cout=%d", h, w, cout);
  B2U_REQUIRE(num_groups > 0 && cout % num_groups == 0, "cout %d not divisible by groups %d", cout, num_groups);
  const int threads = pick_threads(cout / 8);
  B2U_REQUIRE(w % kFirstStrip == 0, "padded width %d must be a multiple of %d", w, kFirstStrip);
  if (rows_per_image) *rows_per_image = blocks_per_image_for(static_cast<long>(h) * (w / kFirstStrip), threads / (cout / 8));
  if (subgroup_size) *subgroup_size = stat_sgs(cout, num_groups);
  return B2U_OK;
}

extern "C" int b2u_conv_first_fwd(const float* x_nchw, const float* w, void* y, float* partials, int n, int cin,
                                  int h0, int w0, int h, int wd, int cout, int num_groups, int dtype, void* stream) {
  B2U_REQUIRE(x_nchw && w && y, "null pointer");
  B2U_REQUIRE(cin == 1 || cin == 3, "first-layer kernel supports cin 1 or 3, got %d", cin);
  B2U_REQUIRE(n > 0 && h0 > 0 && w0 > 0 && h >= h0 && wd >= w0, "bad sizes");
  int rows = 0, sgs = 0;
  int rc = b2u_conv_first_stat_layout(h, wd, cout, num_groups > 0 ? num_groups : 1, &rows, &sgs);
  if (rc) return rc;
  const int threads = pick_threads(cout / 8);
  const size_t smem = (static_cast<size_t>(cout) * cin * 9 + threads * 16 + cout * 2) * sizeof(float);
  dim3 grid(rows, n);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* parts = num_groups > 0 ? partials : nullptr;
  B2U_REQUIRE(num_groups == 0 || partials, "partials required");
#define B2U_LAUNCH_FIRST(T, CIN)                                                                                  \
  do {                                                                                                            \
    B2U_SET_MAX_SMEM_ONCE((conv_first_kernel<T, CIN>), 96 * 1024);                                                \
    B2U_PDL_LAUNCH((conv_first_kernel<T, CIN>), grid, threads, smem, st, x_nchw, w, static_cast<T*>(y), parts, h0, w0, h, wd, cout, sgs); \
  } while (0)
  if (dtype == B2U_F32) {
    if (cin == 1) B2U_LAUNCH_FIRST(float, 1); else B2U_LAUNCH_FIRST(float, 3);
  } else if (dtype == B2U_F16) {
    if (cin == 1) B2U_LAUNCH_FIRST(__half, 1); else B2U_LAUNCH_FIRST(__half, 3);
  } else {
    if (cin == 1) B2U_LAUNCH_FIRST(__nv_bfloat16, 1); else B2U_LAUNCH_FIRST(__nv_bfloat16, 3);
  }
#undef B2U_LAUNCH_FIRST
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_gn_finalize(const float* partials, int rows_per_image, int subgroup_size, const float* gamma,
                               const float* beta, float* coef, int n, int c, int num_groups, double count, float eps,
                               const unsigned long long* keep_counts, int images_per_call, double numel_per_call,
                               float* mean_rstd, void* stream) {
  return b2u_gn_finalize_ex(partials, rows_per_image, subgroup_size, gamma, beta, coef, n, c, num_groups, count, eps,
                            keep_counts, images_per_call, numel_per_call, mean_rstd, 0, stream);
}

extern "C" int b2u_gn_finalize_ex(const float* partials, int rows_per_image, int subgroup_size, const float* gamma,
                                  const float* beta, float* coef, int n, int c, int num_groups, double count, float eps,
                                  const unsigned long long* keep_counts, int images_per_call, double numel_per_call,
                                  float* mean_rstd, int shared_partials, void* stream) {
  B2U_REQUIRE(partials && gamma && beta && coef, "null pointer");
  B2U_REQUIRE(n > 0 && c > 0 && num_groups > 0 && c % num_groups == 0, "bad n/c/groups");
  B2U_REQUIRE(subgroup_size > 0 && (c / num_groups) % subgroup_size == 0, "subgroup size %d does not divide group size %d",
              subgroup_size, c / num_groups);
  B2U_REQUIRE(!keep_counts || images_per_call > 0, "images_per_call must be positive");
  dim3 grid(num_groups, n);
  // the reduction is a chain of L2-latency-bound trips (4 loads in flight per thread): 512 threads for the long partial
  // lists of the shallow levels (2664 tile rows at 592x576: 2 trips instead of 6), 128 for the short ones
  const int fin_threads = static_cast<long>(rows_per_image) * ((c / num_groups) / subgroup_size) >= 1024 ? 512 : 128;
  B2U_PDL_LAUNCH((gn_finalize_kernel), grid, fin_threads, 0, reinterpret_cast<cudaStream_t>(stream), partials, rows_per_image, subgroup_size, gamma, beta, reinterpret_cast<float2*>(coef), c, num_groups, count, eps, keep_counts, images_per_call > 0 ? images_per_call : 1, numel_per_call, reinterpret_cast<float2*>(mean_rstd), shared_partials);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

// Occupancy limiter for the memory-bound forward kernels: B2U_APPLY_SMEM_KB (dynamic shared memory requested per
// block, unused by the kernel) caps the resident blocks per SM so that registers stay free for the DropBlock mask-build
// blocks of the NEXT step, which the Monte-Carlo loop runs concurrently on a low-priority stream.
static size_t apply_pad_smem() {
  static long v = -1;
  if (v < 0) {
    const char* e = getenv("B2U_APPLY_SMEM_KB");
    v = e ? atol(e) * 1024 : 0;
    if (v < 0 || v > 200 * 1024) v = 0;
  }
  return static_cast<size_t>(v);
}

static int fill_apply(const b2u_apply_desc* d, ApplyParams* p) {
  B2U_REQUIRE(d, "null descriptor");
  B2U_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0 && d->c > 0 && d->c % 8 == 0, "bad tensor shape");
  B2U_REQUIRE(d->out_cstride >= d->c && d->out_cstride % 8 == 0 && d->out_coffset % 8 == 0, "bad output channel layout");
  p->n = d->n; p->h = d->h; p->w = d->w; p->c = d->c; p->relu = d->relu;
  p->out_cstride = d->out_cstride; p->out_coffset = d->out_coffset;
  p->mask2_cstride = d->mask2_cstride; p->mask2_coffset = d->mask2_coffset;
  p->images_per_call2 = d->images_per_call2 > 0 ? d->images_per_call2 : 1;
  p->numel_per_call2 = d->numel_per_call2;
  p->x_shared = d->reserved[0] == 1 ? 1 : 0;
  return B2U_OK;
}

extern "C" int b2u_gn_apply(const void* x, const float* coef, const uint32_t* mask1, const uint32_t* mask2,
                            const unsigned long long* keep_counts2, void* out, const b2u_apply_desc* d, void* stream) {
  ApplyParams p;
  int rc = fill_apply(d, &p);
  if (rc) return rc;
  B2U_REQUIRE(x && coef && out, "null pointer");
  B2U_REQUIRE(!mask2 || (keep_counts2 && d->mask2_cstride % 8 == 0 && d->mask2_coffset % 8 == 0), "mask2 needs keep counts and 8-aligned channel layout");
  B2U_REQUIRE(d->c / 8 <= 256, "at most 2048 channels");
  {
    long cmax = d->out_cstride > d->c ? d->out_cstride : d->c;
    if (mask2 && d->mask2_cstride > cmax) cmax = d->mask2_cstride;
    B2U_REQUIRE(static_cast<long>(d->h) * d->w * cmax < (1l << 31), "one image must stay below 2^31 elements (32-bit in-image offsets)");
  }
  const int threads = pick_threads(d->c / 8);
  const int slots = threads / (d->c / 8);
  const long hw = static_cast<long>(d->h) * d->w;
  long bpi = (hw + static_cast<long>(slots) * kApplyUnroll - 1) / (static_cast<long>(slots) * kApplyUnroll);
  const long cap = (static_cast<long>(b2u_num_sms()) * 16 + d->n - 1) / d->n;
  if (bpi > cap) bpi = cap;
  if (bpi < 1) bpi = 1;
  dim3 grid(static_cast<unsigned>(bpi), d->n);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t pad = apply_pad_smem();
#define B2U_APPLY_TM(T, A, B)                                                                                       \
  do {                                                                                                              \
    if (pad > 48 * 1024) B2U_SET_MAX_SMEM_ONCE((gn_apply_kernel<T, A, B>), 200 * 1024);                             \
    B2U_PDL_LAUNCH((gn_apply_kernel<T, A, B>), grid, threads, pad, st, static_cast<const T*>(x),                    \
                   reinterpret_cast<const float2*>(coef), reinterpret_cast<const uint8_t*>(mask1),                  \
                   reinterpret_cast<const uint8_t*>(mask2), keep_counts2, static_cast<T*>(out), p);                 \
  } while (0)
#define B2U_APPLY_T(T)                                                                                              \
  do {                                                                                                              \
    if (mask1 && mask2) B2U_APPLY_TM(T, true, true);                                                                \
    else if (mask1) B2U_APPLY_TM(T, true, false);                                                                   \
    else if (mask2) B2U_APPLY_TM(T, false, true);                                                                   \
    else B2U_APPLY_TM(T, false, false);                                                                             \
  } while (0)
  if (d->dtype == B2U_F32) B2U_APPLY_T(float);
  else if (d->dtype == B2U_F16) B2U_APPLY_T(__half);
  else B2U_APPLY_T(__nv_bfloat16);
#undef B2U_APPLY_TM
#undef B2U_APPLY_T
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_pool_stat_layout(int h, int w, int c, int num_groups, int* rows_per_image, int* subgroup_size) {
  B2U_REQUIRE(h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0, "pool needs even h,w (got %d,%d)", h, w);
  B2U_REQUIRE(c > 0 && c % 8 == 0 && c / 8 <= 256, "bad channel count %d", c);
  B2U_REQUIRE(num_groups > 0 && c % num_groups == 0, "c %d not divisible by groups %d", c, num_groups);
  const int threads = pick_threads(c / 8);
  if (rows_per_image) *rows_per_image = blocks_per_image_for(static_cast<long>(h / 2) * (w / 2), threads / (c / 8) * kPoolUnroll);
  if (subgroup_size) *subgroup_size = stat_sgs(c, num_groups);
  return B2U_OK;
}

extern "C" int b2u_gn_apply_pool(const void* x, const float* coef, const uint32_t* mask1, const uint32_t* mask2,
                                 const unsigned long long* keep_counts2, void* skip_out, void* pooled,
                                 float* pool_partials, uint8_t* argmax, int pool_num_groups, const b2u_apply_desc* d,
                                 void* stream) {
  ApplyParams p;
  int rc = fill_apply(d, &p);
  if (rc) return rc;
  B2U_REQUIRE(x && coef && pooled, "null pointer");
  B2U_REQUIRE(!mask2 || keep_counts2, "mask2 needs keep counts");
  B2U_REQUIRE(p.x_shared == 0, "b2u_gn_apply_pool does not support a shared input");
  {
    const long cmax = d->out_cstride > d->c ? d->out_cstride : d->c;
    B2U_REQUIRE(static_cast<long>(d->h) * d->w * (cmax > d->mask2_cstride ? cmax : d->mask2_cstride) < (1l << 31),
                "one image must stay below 2^31 elements (32-bit in-image offsets)");
  }
  int rows = 0, sgs = 1;
  rc = b2u_pool_stat_layout(d->h, d->w, d->c, pool_num_groups > 0 ? pool_num_groups : 1, &rows, &sgs);
  if (rc) return rc;
  B2U_REQUIRE(pool_num_groups == 0 || pool_partials, "pool partials required");
  const int threads = pick_threads(d->c / 8);
  const size_t smem = (static_cast<size_t>(threads) * 16 + d->c * 2) * sizeof(float);
  dim3 grid(rows, d->n);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* parts = pool_num_groups > 0 ? pool_partials : nullptr;
  const int mode = (mask1 && mask2 && skip_out) ? 1 : (!mask1 && !mask2 && skip_out) ? 0 : 2;
#define B2U_POOL_TAM(T, A, M)                                                                                       \
  B2U_PDL_LAUNCH((gn_apply_pool_kernel<T, A, M>), grid, threads, smem, st, static_cast<const T*>(x),                    \
                 reinterpret_cast<const float2*>(coef), reinterpret_cast<const uint8_t*>(mask1),                        \
                 reinterpret_cast<const uint8_t*>(mask2), keep_counts2, static_cast<T*>(skip_out), static_cast<T*>(pooled), \
                 parts, argmax, sgs, p)
#define B2U_POOL_TA(T, A)                                                                                           \
  do {                                                                                                              \
    if (mode == 1) B2U_POOL_TAM(T, A, 1);                                                                           \
    else if (mode == 0) B2U_POOL_TAM(T, A, 0);                                                                      \
    else B2U_POOL_TAM(T, A, 2);                                                                                     \
  } while (0)
#define B2U_POOL_T(T)                                                                                               \
  do {                                                                                                              \
    if (argmax) B2U_POOL_TA(T, true);                                                                               \
    else B2U_POOL_TA(T, false);                                                                                     \
  } while (0)
  if (d->dtype == B2U_F32) B2U_POOL_T(float);
  else if (d->dtype == B2U_F16) B2U_POOL_T(__half);
  else B2U_POOL_T(__nv_bfloat16);
#undef B2U_POOL_T
#undef B2U_POOL_TA
#undef B2U_POOL_TAM
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

// Launch shape of the c = 64, 16-bit head (host only; CPU-tested through the ABI).  plan[0] = variant (1: cp.async ring
// kernel, 0: register kernel), [1] = threads per block, [2] = blocks, [3] = trips per block, [4] = ring stages,
// [5] = dynamic shared memory per block.  Default: the ring kernel with 2 stages and 224 or 192 threads per block,
// whichever fills the last trip better (tests/exp_head.py, profiles/r02i_exp_head.log: at 10 x 584 x 565 the register
// kernel at 256 threads takes 104 us, the ring kernel 95 us with 2 stages x 224 threads, 101 us with 3 stages, 110 us
// with 4 -- the larger the ring the smaller L1 -- and 110-126 us at 256 threads, whose 9472 resident groups leave the
// fifth trip 35 % full).  Diagnostic overrides, read per call: B2U_HEAD_ASYNC=0 (register kernel, 256 threads),
// B2U_HEAD_THREADS, B2U_HEAD_STAGES.
extern "C" int b2u_head_plan(int h0, int w0, int num_sms, int* plan) {
  B2U_REQUIRE(plan && h0 > 0 && w0 > 0 && num_sms > 0, "bad arguments");
  // 8 pixels per 8-lane group and trip; grid = at most the resident slots (2 blocks per SM), every block walks ceil(trips)
  // of them -- the last trip is partly empty.  (An integral trip count made 258 blocks of 5 trips at 584x565: 110 SMs
  // with two blocks, 38 with one -- the busiest SM did 10 trips where 9 suffice.)
  const long gtrips = (static_cast<long>(h0) * w0 + 7) / 8;
  int nt = 0, use_async = 1, stages = 2;
  if (const char* e = getenv("B2U_HEAD_ASYNC")) use_async = atoi(e) != 0;
  if (const char* e = getenv("B2U_HEAD_THREADS")) nt = atoi(e);
  if (const char* e = getenv("B2U_HEAD_STAGES")) stages = atoi(e);
  const long sms2 = static_cast<long>(num_sms) * 2;                       // blocks resident at once
  if (nt == 0 && !use_async) nt = 256;
  if (nt == 0) {
    double best = -1.0;
    for (int cand = 224; cand >= 192; cand -= 32) {
      const long resident = sms2 * (cand / 8);
      const long tr = (gtrips + resident - 1) / resident;
      const double fill = gtrips >= resident ? static_cast<double>(gtrips) / static_cast<double>(tr * resident) : 1.0;
      if (fill > best + 1e-9) { best = fill; nt = cand; }
    }
  }
  B2U_REQUIRE(nt >= 64 && nt <= 256 && nt % 32 == 0, "B2U_HEAD_THREADS must be 64..256, a multiple of 32");
  B2U_REQUIRE(stages >= 2 && stages <= 4, "B2U_HEAD_STAGES must be 2, 3 or 4");
  const long gpb = nt / 8;                                                // groups per block
  const long slots = sms2 * gpb;                                          // groups resident at once
  const long grid8 = gtrips >= slots ? sms2 : (gtrips + gpb - 1) / gpb;
  const long trips = (gtrips + gpb * grid8 - 1) / (gpb * grid8);
  const int ring_bytes = use_async ? nt * stages * (8 * 16 + 8) : 0;
  B2U_REQUIRE(2 * (ring_bytes + 1024) <= 227 * 1024, "head ring of %d bytes does not fit twice per SM", ring_bytes);
  plan[0] = use_async; plan[1] = nt; plan[2] = static_cast<int>(grid8); plan[3] = static_cast<int>(trips);
  plan[4] = stages; plan[5] = ring_bytes;
  return B2U_OK;
}

extern "C" int b2u_head_fwd(const void* x, const float* coef, const uint32_t* mask1, const float* w_head, float* out,
                            float* logits, const float* fov, double* acc, float* samples, const long long* iter_base,
                            const b2u_head_desc* d, void* stream) {
  B2U_REQUIRE(d && x && coef && w_head, "null pointer");
  B2U_REQUIRE(d->c % 8 == 0 && d->c >= 8 && d->c <= 256 && ((d->c / 8) & (d->c / 8 - 1)) == 0,
              "head supports c in {8,16,32,64,128,256}, got %d", d->c);
  B2U_REQUIRE(d->h0 > 0 && d->w0 > 0 && d->h0 <= d->h && d->w0 <= d->w && d->n > 0, "bad sizes");
  B2U_REQUIRE(out || acc, "nothing to write: out and acc are both NULL");
  const long groups = static_cast<long>(d->h0) * d->w0;
  const int grid = grid_for((groups + kHeadPix - 1) / kHeadPix * (d->c / 8), 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (d->c == 64 && d->dtype != B2U_F32 && static_cast<long>(d->h) * d->w < (1l << 31)) {
    int hp[6];
    const int prc = b2u_head_plan(d->h0, d->w0, b2u_num_sms(), hp);
    if (prc) return prc;
    const int use_async = hp[0], nt = hp[1], grid8 = hp[2], trips = hp[3], stages = hp[4], ring_bytes = hp[5];
#define B2U_HEAD8_T(T)                                                                                              \
  do {                                                                                                              \
    if (use_async && stages == 3) {                                                                                 \
      B2U_SET_MAX_SMEM_ONCE((head8_async_kernel<T, 3>), 112 * 1024);                                                \
      B2U_PDL_LAUNCH((head8_async_kernel<T, 3>), grid8, nt, ring_bytes, st, static_cast<const T*>(x),               \
                     reinterpret_cast<const float2*>(coef), reinterpret_cast<const uint8_t*>(mask1), w_head, out,   \
                     logits, fov, acc, samples, iter_base, *d, trips);                                              \
    } else if (use_async && stages == 2) {                                                                          \
      B2U_SET_MAX_SMEM_ONCE((head8_async_kernel<T, 2>), 112 * 1024);                                                \
      B2U_PDL_LAUNCH((head8_async_kernel<T, 2>), grid8, nt, ring_bytes, st, static_cast<const T*>(x),               \
                     reinterpret_cast<const float2*>(coef), reinterpret_cast<const uint8_t*>(mask1), w_head, out,   \
                     logits, fov, acc, samples, iter_base, *d, trips);                                              \
    } else if (use_async) {                                                                                         \
      B2U_SET_MAX_SMEM_ONCE((head8_async_kernel<T, 4>), 112 * 1024);                                                \
      B2U_PDL_LAUNCH((head8_async_kernel<T, 4>), grid8, nt, ring_bytes, st, static_cast<const T*>(x),               \
                     reinterpret_cast<const float2*>(coef), reinterpret_cast<const uint8_t*>(mask1), w_head, out,   \
                     logits, fov, acc, samples, iter_base, *d, trips);                                              \
    } else {                                                                                                        \
      B2U_PDL_LAUNCH((head8_kernel<T>), grid8, nt, 0, st, static_cast<const T*>(x),                                 \
                     reinterpret_cast<const float2*>(coef), reinterpret_cast<const uint8_t*>(mask1), w_head, out,   \
                     logits, fov, acc, samples, iter_base, *d, trips);                                              \
    }                                                                                                               \
  } while (0)
    if (d->dtype == B2U_F16) B2U_HEAD8_T(__half);
    else B2U_HEAD8_T(__nv_bfloat16);
#undef B2U_HEAD8_T
    B2U_LAUNCH_CHECK();
    return B2U_OK;
  }
#define B2U_HEAD_T(T)                                                                                               \
  B2U_PDL_LAUNCH((head_kernel<T>), grid, 256, 0, st, static_cast<const T*>(x), reinterpret_cast<const float2*>(coef),   \
                 reinterpret_cast<const uint8_t*>(mask1), w_head, out, logits, fov, acc, samples, iter_base, *d)
  if (d->dtype == B2U_F32) B2U_HEAD_T(float);
  else if (d->dtype == B2U_F16) B2U_HEAD_T(__half);
  else B2U_HEAD_T(__nv_bfloat16);
#undef B2U_HEAD_T
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_mc_finalize(const double* acc, float* mean, float* stdv, long long npix, long long t, void* stream) {
  B2U_REQUIRE(acc && mean && stdv && npix > 0 && t > 0, "bad arguments");
  mc_finalize_kernel<<<grid_for(npix, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(acc, mean, stdv, npix, t);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_mc_accumulate(const float* x, const float* fov, double* acc, float* samples,
                                 const long long* iter_base, int n, long long npix, int return_num, void* stream) {
  B2U_REQUIRE(x && acc && n > 0 && npix > 0, "bad arguments");
  mc_accumulate_kernel<<<grid_for(npix, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, fov, acc, samples, iter_base, n,
                                                                                              npix, return_num);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_confusion_counts(const float* seg, const float* gt, const float* mask, long long n,
                                    unsigned long long* counts4, void* stream) {
  B2U_REQUIRE(seg && gt && mask && counts4 && n > 0, "bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B2U_CHECK_CUDA(cudaMemsetAsync(counts4, 0, 4 * sizeof(unsigned long long), st));
  confusion_kernel<<<grid_for(n, 256), 256, 0, st>>>(seg, gt, mask, n, counts4);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_advance_counter(long long* counter, long long delta, void* stream) {
  B2U_REQUIRE(counter, "null counter");
  B2U_PDL_LAUNCH((advance_counter_kernel), 1, 1, 0, reinterpret_cast<cudaStream_t>(stream), counter, delta);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

// theta per torchvision functional.py:1006-1063 with angle -> -angle (functional.py:1130), evaluated in double, cast to
// fp32, then rescaled by [0.5 w, 0.5 h] in fp32 (_functional_tensor.py:598).
static b2u::RotParams rot_params_for(double angle_deg, int h, int w) {
  const double rot = -angle_deg * 3.14159265358979323846 / 180.0;   // math.radians(-angle)
  const double a = cos(rot), b = -sin(rot), cc = sin(rot), dd = cos(rot);
  const float m0 = static_cast<float>(dd), m1 = static_cast<float>(-b), m2 = 0.f;
  const float m3 = static_cast<float>(-cc), m4 = static_cast<float>(a), m5 = 0.f;
  const float sx = 0.5f * w, sy = 0.5f * h;
  // rescaled_theta = theta^T / [sx, sy]: column 0 (gx) divides by sx, column 1 (gy) by sy
  b2u::RotParams r;
  r.t00 = m0 / sx; r.t01 = m1 / sx; r.t02 = m2 / sx;
  r.t10 = m3 / sy; r.t11 = m4 / sy; r.t12 = m5 / sy;
  return r;
}

extern "C" int b2u_rotate_bilinear(const float* x, float* out, int n, int c, int h, int w, const double* angles_deg,
                                   int x_batch_stride_is_zero, void* stream) {
  B2U_REQUIRE(x && out && angles_deg && n > 0 && c > 0 && h > 0 && w > 0, "bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long img_stride = static_cast<long>(c) * h * w;
  // The per-angle coefficients travel by value as a kernel parameter: no allocation, CUDA-graph capturable.
  for (int base = 0; base < n; base += kMaxAnglesPerLaunch) {
    const int nb = n - base < kMaxAnglesPerLaunch ? n - base : kMaxAnglesPerLaunch;
    RotBatch rb;
    for (int i = 0; i < nb; ++i) rb.r[i] = rot_params_for(angles_deg[base + i], h, w);
    const long total = static_cast<long>(nb) * h * w;
    rotate_kernel<<<grid_for(total, 256), 256, 0, st>>>(x + (x_batch_stride_is_zero ? 0 : base * img_stride),
                                                       out + base * img_stride, nb, c, h, w, rb,
                                                       x_batch_stride_is_zero ? 0 : img_stride);
    B2U_LAUNCH_CHECK();
  }
  return B2U_OK;
}

extern "C" int b2u_rotation_table(const double* angles_deg, int n, int h, int w, float* table_host) {
  B2U_REQUIRE(angles_deg && table_host && n > 0 && h > 0 && w > 0, "bad arguments");
  static_assert(sizeof(b2u::RotParams) == 6 * sizeof(float), "table row = 6 floats");
  for (int i = 0; i < n; ++i) {
    const b2u::RotParams r = rot_params_for(angles_deg[i], h, w);
    memcpy(table_host + 6 * static_cast<size_t>(i), &r, sizeof(r));
  }
  return B2U_OK;
}

extern "C" int b2u_rotate_in_table(const float* x, float* out, int n, int c, int h, int w, const float* table_dev, int table_len,
                                   const long long* iter_base_dev, void* stream) {
  B2U_REQUIRE(x && out && table_dev && iter_base_dev && n > 0 && c > 0 && h > 0 && w > 0 && table_len > 0, "bad arguments");
  B2U_REQUIRE(static_cast<long>(h) * w < (1L << 31), "image too large");
  rotate_in_table_kernel<<<grid_for(static_cast<long>(n) * h * w, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, out, n, c, h, w, reinterpret_cast<const b2u::RotParams*>(table_dev), table_len, iter_base_dev);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_rotate_back_accumulate(const float* seg, const float* fov, double* acc, float* samples, int n, int h, int w,
                                          int return_num, const float* table_dev, int table_len, const long long* iter_base_dev,
                                          const long long* iter_limit_dev, void* stream) {
  B2U_REQUIRE(seg && acc && table_dev && iter_base_dev && iter_limit_dev && n > 0 && h > 0 && w > 0 && table_len > 0, "bad arguments");
  B2U_REQUIRE(static_cast<long>(h) * w < (1L << 31), "image too large");
  rotate_back_accumulate_kernel<<<grid_for(static_cast<long>(h) * w, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      seg, fov, acc, samples, n, h, w, return_num, reinterpret_cast<const b2u::RotParams*>(table_dev), table_len, iter_base_dev,
      iter_limit_dev);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}
