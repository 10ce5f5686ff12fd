// Backward of the fused elementwise tail of a conv unit (GroupNorm -> DropBlock -> ReLU), of the max-pool,
// of the concat-site DropBlock and of the output head.  Two passes per unit over the same inputs:
//
//   pass 1 (unit_bwd_stats):  dZ = g * relu'(z) * mask1 * s1   and per-(image, channel) partial sums
//                             P1 = sum dZ,  P2 = sum dZ * xhat   (xhat = (y - mean) * rstd),  [P3 = sum dlogit * act]
//   bwd_finalize           :  dgamma = sum_n P2, dbeta = sum_n P1, per-(image, group) c1 = S1/cnt, c2 = S2/cnt with
//                             S1 = sum_{c in g} gamma_c P1, S2 = sum_{c in g} gamma_c P2   (fp64, fixed order)
//   pass 2 (unit_bwd_apply):  dY = rstd * (gamma * dZ - c1 - xhat * c2)   (GroupNorm backward), bf16/fp32 NHWC or the
//                             space-to-depth layout [N, H/2, W/2, 4, C] the transposed conv's gradient GEMMs read.
//
// The upstream gradient g of a pixel/channel is the sum of up to three sources, none of which is materialised:
//   A: a dense NHWC tensor (channel window of a wider tensor), optionally times the concat-site keep mask and rescale
//      (backward of `x = dropblock(cat([up, skip]))`, reference utils_unet.py:382-383);
//   P: the max-pool backward -- the pooled-resolution gradient goes to the window element whose index equals the
//      stored argmax code (nn.MaxPool2d, utils_unet.py:265-266);
//   H: the head -- dlogit * w_head with dlogit = grad_out * out * (1 - out) inside the un-padded image
//      (Conv2d 1x1 + Sigmoid + crop, utils_unet.py:397-404,440).
#include "b2u_common.cuh"

namespace b2u {

struct UnitBwdParams {
  int n, h, w, c;
  int relu;
  // unit
  int images_per_call1;
  double numel_per_call1;
  int num_groups;
  // source A
  int a_cstride, a_coffset;
  int mask2_cstride, mask2_coffset, images_per_call2;
  double numel_per_call2;
  // source H
  int h0, w0;
  // output
  int s2d;                 // 1: write dY as [N, H/2, W/2, 4, C]
  int rows;                // partial rows per image (pass 1)
};

struct UnitBwdPtrs {
  const void* y;                          // raw conv output [N,H,W,C]
  const float2* coef;                     // [N][C] (a*s1, b*s1)
  const float2* mr;                       // [N][G] (mean, rstd)
  const float* gamma;                     // [C]
  const uint8_t* mask1;                   // own-site keep mask bytes or null
  const unsigned long long* keep1;        // own-site keep counts or null
  const void* ga;                         // source A or null
  const uint8_t* mask2;                   // concat-site mask bytes or null
  const unsigned long long* keep2;
  const void* gp;                         // source P: pooled gradient [N,H/2,W/2,C] or null
  const uint8_t* argmax;                  // [N,H/2,W/2,C]
  const float* grad_out;                  // source H: [N,1,h0,w0] or null
  const float* out;                       // [N,1,h0,w0]
  const float* w_head;                    // [C]
  const float2* gcoef;                    // pass 2: [N][G] (c1, c2)
  float* partials;                        // pass 1: [N][rows][C][3]
  void* dy;                               // pass 2 output
};

template <typename T>
__device__ __forceinline__ void load8f(const T* p, float (&f)[8]) {
  Vec8<T> v;
  v.load(p);
  v.to_float(f);
}

// Everything one (pixel, channel vector) needs from global memory, loaded in ONE phase: the three gradient
// sources and the masks are independent loads, and issuing them back to back (before the first use) is what keeps a
// latency-bound pass like this one near the HBM roofline -- the first version interleaved load and use and spent
// 60 % of its cycles in three serialised long-scoreboard waits (profiles/r01_bwd_stats_source.md).
template <typename T>
struct UnitRaw {
  Vec8<T> y, ga, gp;
  uint2 am;
  uint32_t m1, m2;
  float o, go;
  bool inside;
};

// HAS_P / HAS_H: the pool-backward / head source exists (compile time: the A-only units -- 16 of the 26 -- then carry no
// pooled-index, argmax, crop or (row, column) arithmetic at all; the first version decided all of it per pixel at run
// time and spent ~280 instructions per 8-channel vector, issue-bound at 1.2-2.0 TB/s).  `pix` is the pixel index inside
// image n; every tensor is addressed as per-image base + 32-bit offset (the host checks h * w * channels < 2^31).
template <typename T, bool HAS_P, bool HAS_H>
__device__ __forceinline__ void unit_load(const UnitBwdParams& p, const UnitBwdPtrs& q, int n, int pix, int hh, int ww, int cv, UnitRaw<T>& r) {
  const int npix = p.h * p.w;
  r.y.load(reinterpret_cast<const T*>(q.y) + static_cast<size_t>(n) * npix * p.c + (pix * p.c + cv * 8));
  r.m1 = q.mask1 ? q.mask1[static_cast<size_t>(n) * npix * (p.c >> 3) + (pix * (p.c >> 3) + cv)] : 0xFFu;
  r.m2 = 0xFFu;
  if (!HAS_H) {
    r.ga.load(reinterpret_cast<const T*>(q.ga) + static_cast<size_t>(n) * npix * p.a_cstride + (pix * p.a_cstride + p.a_coffset + cv * 8));
    if (q.mask2) r.m2 = q.mask2[static_cast<size_t>(n) * npix * (p.mask2_cstride >> 3) + (pix * (p.mask2_cstride >> 3) + (p.mask2_coffset >> 3) + cv)];
  }
  if (HAS_P) {
    const int pp = (hh >> 1) * (p.w >> 1) + (ww >> 1);
    const size_t pbase = static_cast<size_t>(n) * (npix >> 2) * p.c;
    r.gp.load(reinterpret_cast<const T*>(q.gp) + pbase + (pp * p.c + cv * 8));
    r.am = *reinterpret_cast<const uint2*>(q.argmax + pbase + (pp * p.c + cv * 8));
  }
  r.inside = false;
  r.o = r.go = 0.f;
  if (HAS_H) {
    if (hh < p.h0 && ww < p.w0) {
      const size_t op = static_cast<size_t>(n) * p.h0 * p.w0 + (hh * p.w0 + ww);
      r.o = q.out[op];
      r.go = q.grad_out[op];
      r.inside = true;
    }
  }
}

// upstream gradient g[8] -> dZ[8], raw y[8] (returned in xhat) and the activation for one (pixel, channel vector)
template <typename T, bool HAS_P, bool HAS_H>
__device__ __forceinline__ void unit_grad8(const UnitBwdParams& p, const UnitBwdPtrs& q, const UnitRaw<T>& r, int hh, int ww, int cv,
                                           const float (&a)[8], const float (&b)[8], float s1, float s2,
                                           float (&dz)[8], float (&xhat)[8], float (&act)[8], float& dlogit) {
  float yv[8];
  r.y.to_float(yv);
  float g[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) g[i] = 0.f;
  if (!HAS_H) {
    float t[8];
    r.ga.to_float(t);
    if (q.mask2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = ((r.m2 >> i) & 1u) ? t[i] * s2 : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] += t[i];
  }
  if (HAS_P) {
    float t[8];
    r.gp.to_float(t);
    const uint32_t code = ((hh & 1) << 1) | (ww & 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t a8 = ((i < 4 ? r.am.x : r.am.y) >> (8 * (i & 3))) & 0xFFu;
      if (a8 == code) g[i] += t[i];
    }
  }
  dlogit = 0.f;
  if (HAS_H) {
    if (r.inside) dlogit = r.go * r.o * (1.f - r.o);
    const float4 wa = __ldg(reinterpret_cast<const float4*>(q.w_head + cv * 8));
    const float4 wb = __ldg(reinterpret_cast<const float4*>(q.w_head + cv * 8 + 4));
    g[0] += dlogit * wa.x; g[1] += dlogit * wa.y; g[2] += dlogit * wa.z; g[3] += dlogit * wa.w;
    g[4] += dlogit * wb.x; g[5] += dlogit * wb.y; g[6] += dlogit * wb.z; g[7] += dlogit * wb.w;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float z = fmaf(yv[i], a[i], b[i]);                    // s1 * (gamma * xhat + beta): same sign as the pre-activation
    const bool keep = (r.m1 >> i) & 1u;
    const bool pass = keep && (!p.relu || z > 0.f);
    act[i] = keep ? (p.relu ? fmaxf(z, 0.f) : z) : 0.f;
    dz[i] = pass ? g[i] * s1 : 0.f;
    xhat[i] = yv[i];
  }
}

constexpr int kBwdUnroll = 2;       // pixels per thread per trip (all loads of both issued before the first use)

// grid = (rows, n); thread owns channel vector t % cvs; deterministic block reduction to partials[n][row][c][3]
// GPV = GroupNorm groups per 8-channel vector (1 when the group size is >= 8, else 8 / group size): the per-group
// (mean, rstd[, c1, c2]) live in GPV registers each instead of 8.
template <typename T, int GPV, bool HAS_P, bool HAS_H>
__global__ void __launch_bounds__(256, 2) unit_bwd_stats_kernel(UnitBwdParams p, UnitBwdPtrs q) {
  pdl_wait();       // predecessor grid complete + visible (no-op without the PDL launch attribute)
  pdl_trigger();    // the successor may be scheduled once every CTA got here
  extern __shared__ float sm[];
  const int n = blockIdx.y;
  const int cvs = p.c >> 3;
  const int cv = threadIdx.x % cvs;
  const int slot = threadIdx.x / cvs;
  const int slots = blockDim.x / cvs;
  const int gsize = p.c / p.num_groups;
  constexpr int CPG = 8 / GPV;                                  // channels of the vector per group slot
  float a[8], b[8];
  {
    const float4* cp = reinterpret_cast<const float4*>(q.coef + static_cast<size_t>(n) * p.c + cv * 8);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = __ldg(cp + i);
      a[2 * i] = t.x; b[2 * i] = t.y; a[2 * i + 1] = t.z; b[2 * i + 1] = t.w;
    }
  }
  float s1 = 1.f, s2 = 1.f;
  if (q.keep1) s1 = static_cast<float>(p.numel_per_call1 / static_cast<double>(q.keep1[n / p.images_per_call1]));
  if (q.mask2) s2 = static_cast<float>(p.numel_per_call2 / static_cast<double>(q.keep2[n / p.images_per_call2]));
  float mean8[GPV], rstd8[GPV];
#pragma unroll
  for (int i = 0; i < GPV; ++i) {
    const float2 t = __ldg(q.mr + static_cast<size_t>(n) * p.num_groups + (cv * 8 + i * CPG) / gsize);
    mean8[i] = t.x;
    rstd8[i] = t.y;
  }
  float acc1[8], acc2[8], acc3[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc1[i] = acc2[i] = acc3[i] = 0.f;
  const int npix = p.h * p.w;
  const int stride = gridDim.x * slots * kBwdUnroll;
  for (int base = blockIdx.x * slots * kBwdUnroll + slot; base < npix; base += stride) {
    UnitRaw<T> raw[kBwdUnroll];
    int hh[kBwdUnroll], ww[kBwdUnroll];
#pragma unroll
    for (int u = 0; u < kBwdUnroll; ++u) {
      const int pix = base + u * slots;
      hh[u] = ww[u] = 0;
      if (HAS_P || HAS_H) {                                   // (row, column) only where a source needs them
        hh[u] = pix / p.w;
        ww[u] = pix - hh[u] * p.w;
      }
      if (pix < npix) unit_load<T, HAS_P, HAS_H>(p, q, n, pix, hh[u], ww[u], cv, raw[u]);
    }
#pragma unroll
    for (int u = 0; u < kBwdUnroll; ++u) {
      if (base + u * slots < npix) {
        float dz[8], xhat[8], act[8], dlogit;
        unit_grad8<T, HAS_P, HAS_H>(p, q, raw[u], hh[u], ww[u], cv, a, b, s1, s2, dz, xhat, act, dlogit);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = (xhat[i] - mean8[i / CPG]) * rstd8[i / CPG];   // xhat[] holds the raw conv output
          acc1[i] += dz[i];
          acc2[i] += dz[i] * xh;
          if (HAS_H) acc3[i] += dlogit * act[i];
        }
      }
    }
  }
  // block reduction: [thread][24] -> [c][3]
  float* mine = sm + threadIdx.x * 24;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mine[i] = acc1[i];
    mine[8 + i] = acc2[i];
    mine[16 + i] = acc3[i];
  }
  __syncthreads();
  float* out_row = q.partials + (static_cast<size_t>(n) * gridDim.x + blockIdx.x) * p.c * 3;
  for (int o = threadIdx.x; o < p.c * 3; o += blockDim.x) {
    const int ch = o / 3, st = o - ch * 3;
    const int ccv = ch >> 3, ci = ch & 7;
    float acc = 0.f;
    for (int sl = 0; sl < slots; ++sl) acc += sm[(sl * cvs + ccv) * 24 + st * 8 + ci];
    out_row[o] = acc;
  }
}

// One block per GroupNorm group reduces the partial rows of every image (fp64, fixed order => deterministic):
//   per (image, group):  c1 = sum_c gamma_c P1 / cnt,  c2 = sum_c gamma_c P2 / cnt     -> gcoef[n][g]
//   per channel:         dbeta = sum_n P1, dgamma = sum_n P2, dw_head = sum_n P3
// A partial row holds the group's gsize*3 floats contiguously; a warp covers 32/(gsize*3) rows per trip when the
// group is narrow (C = 64: 6 floats per row) so the loads stay coalesced and every lane has work.
constexpr int kBwdFinWarps = 16;
__global__ void __launch_bounds__(kBwdFinWarps * 32) bwd_finalize_kernel(const float* __restrict__ partials, int n, int rows, int c,
                                                                       int num_groups, const float* __restrict__ gamma, double count,
                                                                       float2* __restrict__ gcoef, float* __restrict__ dgamma,
                                                                       float* __restrict__ dbeta, float* __restrict__ dw_head) {
  pdl_wait();       // predecessor grid complete + visible (no-op without the PDL launch attribute)
  pdl_trigger();    // the successor may be scheduled once every CTA got here
  const int g = blockIdx.x;
  const int gsize = c / num_groups;
  const int E = gsize * 3;                                   // floats of this group in one partial row
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rpw = E <= 32 ? 32 / E : 1;                      // rows per warp trip
  const int nj = E <= 32 ? 1 : (E + 31) / 32;                // 32-float chunks per row (<= 3 for gsize <= 32)
  const int rsub = E <= 32 ? lane / E : 0;
  const int e0 = E <= 32 ? lane - rsub * E : lane;
  const bool lane_on = E <= 32 ? (rsub < rpw) : true;
  __shared__ double red[kBwdFinWarps][32][3];
  __shared__ double tot[96];                                 // per (channel, stat) sums over the images
  for (int i = threadIdx.x; i < 96; i += blockDim.x) tot[i] = 0.0;
  for (int img = 0; img < n; ++img) {
    const float* base = partials + (static_cast<size_t>(img) * rows * c + static_cast<size_t>(g) * gsize) * 3;
    double acc[3] = {0.0, 0.0, 0.0};
    // four rows in flight per lane, added in the original order (the loop was bound by L2 latency per trip)
    constexpr int kRowStep = kBwdFinWarps;
    const int step = kRowStep * rpw;
    int r = warp * rpw + rsub;
    if (lane_on) {
      for (; r + 3 * step < rows; r += 4 * step) {
        float v[4][3];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float* rowp = base + static_cast<size_t>(r + u * step) * c * 3;
#pragma unroll
          for (int j = 0; j < 3; ++j) v[u][j] = (j < nj && e0 + 32 * j < E) ? __ldg(rowp + e0 + 32 * j) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int j = 0; j < 3; ++j)
            if (j < nj && e0 + 32 * j < E) acc[j] += static_cast<double>(v[u][j]);
      }
      for (; r < rows; r += step) {
        const float* rowp = base + static_cast<size_t>(r) * c * 3;
#pragma unroll
        for (int j = 0; j < 3; ++j)
          if (j < nj && e0 + 32 * j < E) acc[j] += static_cast<double>(rowp[e0 + 32 * j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) red[warp][lane][j] = acc[j];
    __syncthreads();
    // element e of the row = (channel e / 3, stat e % 3)
    if (threadIdx.x < E) {
      const int e = threadIdx.x;
      double sum = 0.0;
      if (E <= 32) {
        for (int w = 0; w < kBwdFinWarps; ++w)
          for (int rs = 0; rs < rpw; ++rs) sum += red[w][rs * E + e][0];
      } else {
        for (int w = 0; w < kBwdFinWarps; ++w) sum += red[w][e & 31][e >> 5];
      }
      red[0][0][0] = red[0][0][0];                           // (keeps the compiler from hoisting across the barrier)
      tot[e] += sum;
      // stash this image's per-element sums for the group coefficients
      reinterpret_cast<double*>(red)[kBwdFinWarps * 96 - 96 + e] = sum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const double* cur = reinterpret_cast<double*>(red) + kBwdFinWarps * 96 - 96;
      double s1 = 0.0, s2 = 0.0;
      for (int k = 0; k < gsize; ++k) {
        const double gm = static_cast<double>(gamma[g * gsize + k]);
        s1 += gm * cur[k * 3];
        s2 += gm * cur[k * 3 + 1];
      }
      gcoef[static_cast<size_t>(img) * num_groups + g] = make_float2(static_cast<float>(s1 / count), static_cast<float>(s2 / count));
    }
    __syncthreads();
  }
  if (threadIdx.x < gsize) {
    const int ch = g * gsize + threadIdx.x;
    if (dbeta) dbeta[ch] = static_cast<float>(tot[threadIdx.x * 3]);
    if (dgamma) dgamma[ch] = static_cast<float>(tot[threadIdx.x * 3 + 1]);
    if (dw_head) dw_head[ch] = static_cast<float>(tot[threadIdx.x * 3 + 2]);
  }
}

template <typename T, int GPV, bool HAS_P, bool HAS_H>
__global__ void __launch_bounds__(256, 2) unit_bwd_apply_kernel(UnitBwdParams p, UnitBwdPtrs q) {
  pdl_wait();       // predecessor grid complete + visible (no-op without the PDL launch attribute)
  pdl_trigger();    // the successor may be scheduled once every CTA got here
  const int n = blockIdx.y;
  const int cvs = p.c >> 3;
  const int cv = threadIdx.x % cvs;
  const int slot = threadIdx.x / cvs;
  const int slots = blockDim.x / cvs;
  const int gsize = p.c / p.num_groups;
  constexpr int CPG = 8 / GPV;
  float a[8], b[8], gm[8];
  {
    const float4* cp = reinterpret_cast<const float4*>(q.coef + static_cast<size_t>(n) * p.c + cv * 8);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = __ldg(cp + i);
      a[2 * i] = t.x; b[2 * i] = t.y; a[2 * i + 1] = t.z; b[2 * i + 1] = t.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) gm[i] = __ldg(q.gamma + cv * 8 + i);
  }
  float s1 = 1.f, s2 = 1.f;
  if (q.keep1) s1 = static_cast<float>(p.numel_per_call1 / static_cast<double>(q.keep1[n / p.images_per_call1]));
  if (q.mask2) s2 = static_cast<float>(p.numel_per_call2 / static_cast<double>(q.keep2[n / p.images_per_call2]));
  float mean8[GPV], rstd8[GPV], c1[GPV], c2[GPV];
#pragma unroll
  for (int i = 0; i < GPV; ++i) {
    const int g = (cv * 8 + i * CPG) / gsize;
    const float2 t = __ldg(q.mr + static_cast<size_t>(n) * p.num_groups + g);
    const float2 u = __ldg(q.gcoef + static_cast<size_t>(n) * p.num_groups + g);
    mean8[i] = t.x; rstd8[i] = t.y; c1[i] = u.x; c2[i] = u.y;
  }
  const int npix = p.h * p.w;
  T* dy = reinterpret_cast<T*>(q.dy);
  const int stride = gridDim.x * slots * kBwdUnroll;
  for (int base = blockIdx.x * slots * kBwdUnroll + slot; base < npix; base += stride) {
    UnitRaw<T> raw[kBwdUnroll];
    int hh[kBwdUnroll], ww[kBwdUnroll];
#pragma unroll
    for (int u = 0; u < kBwdUnroll; ++u) {
      const int pix = base + u * slots;
      hh[u] = ww[u] = 0;
      if (HAS_P || HAS_H || p.s2d) {                          // (row, column) only where a source or the output layout needs them
        hh[u] = pix / p.w;
        ww[u] = pix - hh[u] * p.w;
      }
      if (pix < npix) unit_load<T, HAS_P, HAS_H>(p, q, n, pix, hh[u], ww[u], cv, raw[u]);
    }
#pragma unroll
    for (int u = 0; u < kBwdUnroll; ++u) {
      if (base + u * slots < npix) {
        float dz[8], xhat[8], act[8], dlogit;
        unit_grad8<T, HAS_P, HAS_H>(p, q, raw[u], hh[u], ww[u], cv, a, b, s1, s2, dz, xhat, act, dlogit);
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = (xhat[i] - mean8[i / CPG]) * rstd8[i / CPG];
          o[i] = rstd8[i / CPG] * (gm[i] * dz[i] - c1[i / CPG] - xh * c2[i / CPG]);
        }
        int dst;                                              // offset inside image n
        if (p.s2d) {
          const int pp = (hh[u] >> 1) * (p.w >> 1) + (ww[u] >> 1);
          dst = (pp * 4 + (((hh[u] & 1) << 1) | (ww[u] & 1))) * p.c + cv * 8;
        } else {
          dst = (base + u * slots) * p.c + cv * 8;
        }
        Vec8<T> v;
        v.from_float(o);
        v.store(dy + static_cast<size_t>(n) * npix * p.c + dst);
      }
    }
  }
}

static int bwd_pick_threads(int cvs) { return cvs > 256 ? 0 : (256 / cvs) * cvs; }

static int bwd_rows(int h, int w, int c) {
  const int threads = bwd_pick_threads(c / 8);
  const int slots = threads / (c / 8);
  long b = (static_cast<long>(h) * w + slots * kBwdUnroll - 1) / (slots * kBwdUnroll);
  long cap = static_cast<long>(b2u_num_sms()) * 2;            // exactly one wave at 2 resident blocks per SM
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace b2u

using namespace b2u;

static int fill_unit(const b2u_unit_bwd_desc* d, UnitBwdParams* p, UnitBwdPtrs* q) {
  B2U_REQUIRE(d, "null descriptor");
  B2U_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0 && d->c > 0 && d->c % 8 == 0 && d->c / 8 <= 256, "bad tensor shape");
  B2U_REQUIRE(d->num_groups > 0 && d->c % d->num_groups == 0, "bad group count");
  B2U_REQUIRE(d->y && d->coef && d->mean_rstd && d->gamma, "null unit tensors");
  B2U_REQUIRE(d->grad_a || d->grad_out, "no upstream gradient source");
  // the kernels are specialised for the three source sets the backward schedule produces
  B2U_REQUIRE(!d->grad_out || (!d->grad_a && !d->grad_pool), "the head source (grad_out) stands alone");
  B2U_REQUIRE(!d->grad_pool || d->grad_a, "the pool source (grad_pool) comes with a dense source (grad_a)");
  {
    const long long widest = d->a_cstride > d->c ? d->a_cstride : d->c;
    B2U_REQUIRE(static_cast<long long>(d->h) * d->w * widest < (1ll << 31), "image too large for 32-bit in-image offsets");
  }
  B2U_REQUIRE(!d->grad_pool || (d->argmax && d->h % 2 == 0 && d->w % 2 == 0), "pool source needs argmax and even h,w");
  B2U_REQUIRE(!d->grad_out || (d->out && d->w_head), "head source needs out and w_head");
  B2U_REQUIRE(!d->mask2 || d->keep_counts2, "mask2 needs keep counts");
  B2U_REQUIRE(!d->mask1 || d->keep_counts1, "mask1 needs keep counts");
  p->n = d->n; p->h = d->h; p->w = d->w; p->c = d->c; p->relu = d->relu;
  p->images_per_call1 = d->images_per_call1 > 0 ? d->images_per_call1 : 1;
  p->numel_per_call1 = d->numel_per_call1;
  p->num_groups = d->num_groups;
  p->a_cstride = d->a_cstride; p->a_coffset = d->a_coffset;
  p->mask2_cstride = d->mask2_cstride; p->mask2_coffset = d->mask2_coffset;
  p->images_per_call2 = d->images_per_call2 > 0 ? d->images_per_call2 : 1;
  p->numel_per_call2 = d->numel_per_call2;
  p->h0 = d->h0; p->w0 = d->w0;
  p->s2d = d->s2d;
  p->rows = bwd_rows(d->h, d->w, d->c);
  q->y = d->y; q->coef = reinterpret_cast<const float2*>(d->coef); q->mr = reinterpret_cast<const float2*>(d->mean_rstd);
  q->gamma = d->gamma;
  q->mask1 = reinterpret_cast<const uint8_t*>(d->mask1); q->keep1 = d->mask1 ? d->keep_counts1 : nullptr;
  q->ga = d->grad_a; q->mask2 = reinterpret_cast<const uint8_t*>(d->mask2); q->keep2 = d->keep_counts2;
  q->gp = d->grad_pool; q->argmax = d->argmax;
  q->grad_out = d->grad_out; q->out = d->out; q->w_head = d->w_head;
  q->gcoef = nullptr; q->partials = nullptr; q->dy = nullptr;
  return B2U_OK;
}

extern "C" int b2u_unit_bwd_rows(int h, int w, int c, int* rows_per_image) {
  B2U_REQUIRE(h > 0 && w > 0 && c > 0 && c % 8 == 0 && c / 8 <= 256 && rows_per_image, "bad arguments");
  *rows_per_image = bwd_rows(h, w, c);
  return B2U_OK;
}

extern "C" int b2u_unit_bwd_stats(const b2u_unit_bwd_desc* d, float* partials, void* stream) {
  UnitBwdParams p;
  UnitBwdPtrs q;
  int rc = fill_unit(d, &p, &q);
  if (rc) return rc;
  B2U_REQUIRE(partials, "null partials");
  q.partials = partials;
  const int threads = bwd_pick_threads(d->c / 8);
  dim3 grid(p.rows, d->n);
  const size_t smem = static_cast<size_t>(threads) * 24 * sizeof(float);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int gsize = d->c / d->num_groups;
  const int gpv = gsize >= 8 ? 1 : 8 / gsize;
  B2U_REQUIRE(gpv == 1 || gpv == 2 || gpv == 4, "group size %d must be 2, 4 or a multiple of 8", gsize);
#define B2U_STATS_S(T, HP, HH)                                                                       \
  do {                                                                                               \
    if (gpv == 1) B2U_PDL_LAUNCH((unit_bwd_stats_kernel<T, 1, HP, HH>), grid, threads, smem, st, p, q);      \
    else if (gpv == 2) B2U_PDL_LAUNCH((unit_bwd_stats_kernel<T, 2, HP, HH>), grid, threads, smem, st, p, q); \
    else B2U_PDL_LAUNCH((unit_bwd_stats_kernel<T, 4, HP, HH>), grid, threads, smem, st, p, q);               \
  } while (0)
#define B2U_STATS(T)                                                                  \
  do {                                                                                \
    if (d->grad_out) B2U_STATS_S(T, false, true);                                     \
    else if (d->grad_pool) B2U_STATS_S(T, true, false);                               \
    else B2U_STATS_S(T, false, false);                                                \
  } while (0)
  if (d->dtype == B2U_F32) B2U_STATS(float);
  else B2U_STATS(__nv_bfloat16);
#undef B2U_STATS
#undef B2U_STATS_S
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_unit_bwd_finalize(const float* partials, int n, int rows_per_image, int c, int num_groups, const float* gamma,
                                     double count, float* group_coef, float* dgamma, float* dbeta, float* dw_head, void* stream) {
  B2U_REQUIRE(partials && gamma && group_coef && n > 0 && c > 0 && num_groups > 0 && c % num_groups == 0, "bad arguments");
  B2U_REQUIRE(c / num_groups <= 32, "group size %d > 32 is not supported by the backward finalise", c / num_groups);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B2U_PDL_LAUNCH((bwd_finalize_kernel), num_groups, kBwdFinWarps * 32, 0, st, partials, n, rows_per_image, c, num_groups, gamma, count, reinterpret_cast<float2*>(group_coef), dgamma, dbeta, dw_head);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_unit_bwd_apply(const b2u_unit_bwd_desc* d, const float* group_coef, void* dy, void* stream) {
  UnitBwdParams p;
  UnitBwdPtrs q;
  int rc = fill_unit(d, &p, &q);
  if (rc) return rc;
  B2U_REQUIRE(group_coef && dy, "null pointer");
  B2U_REQUIRE(!d->s2d || (d->h % 2 == 0 && d->w % 2 == 0), "space-to-depth output needs even h,w");
  q.gcoef = reinterpret_cast<const float2*>(group_coef);
  q.dy = dy;
  const int threads = bwd_pick_threads(d->c / 8);
  const int slots = threads / (d->c / 8);
  long bpi = (static_cast<long>(d->h) * d->w + slots * kBwdUnroll - 1) / (slots * kBwdUnroll);
  const long cap = (static_cast<long>(b2u_num_sms()) * 8 + d->n - 1) / d->n;
  if (bpi > cap) bpi = cap;
  dim3 grid(static_cast<unsigned>(bpi), d->n);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int gsize = d->c / d->num_groups;
  const int gpv = gsize >= 8 ? 1 : 8 / gsize;
  B2U_REQUIRE(gpv == 1 || gpv == 2 || gpv == 4, "group size %d must be 2, 4 or a multiple of 8", gsize);
#define B2U_APPLY_S(T, HP, HH)                                                                    \
  do {                                                                                            \
    if (gpv == 1) B2U_PDL_LAUNCH((unit_bwd_apply_kernel<T, 1, HP, HH>), grid, threads, 0, st, p, q);      \
    else if (gpv == 2) B2U_PDL_LAUNCH((unit_bwd_apply_kernel<T, 2, HP, HH>), grid, threads, 0, st, p, q); \
    else B2U_PDL_LAUNCH((unit_bwd_apply_kernel<T, 4, HP, HH>), grid, threads, 0, st, p, q);               \
  } while (0)
#define B2U_APPLY(T)                                                               \
  do {                                                                             \
    if (d->grad_out) B2U_APPLY_S(T, false, true);                                  \
    else if (d->grad_pool) B2U_APPLY_S(T, true, false);                            \
    else B2U_APPLY_S(T, false, false);                                             \
  } while (0)
  if (d->dtype == B2U_F32) B2U_APPLY(float);
  else B2U_APPLY(__nv_bfloat16);
#undef B2U_APPLY
#undef B2U_APPLY_S
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}
