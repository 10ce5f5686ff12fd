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

// upstream gradient g[8] and dZ[8], xhat[8] for one (pixel, channel vector)
template <typename T>
__device__ __forceinline__ void unit_grad8(const UnitBwdParams& p, const UnitBwdPtrs& q, int n, int hh, int ww, int cv,
                                           const float (&a)[8], const float (&b)[8], float mean, float rstd, float s1, float s2,
                                           const float (&wh)[8], float (&dz)[8], float (&xhat)[8], float (&act)[8], float& dlogit) {
  const long pix = (static_cast<long>(n) * p.h + hh) * p.w + ww;
  float yv[8];
  load8f(reinterpret_cast<const T*>(q.y) + pix * p.c + cv * 8, yv);
  float g[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) g[i] = 0.f;
  if (q.ga) {
    float t[8];
    load8f(reinterpret_cast<const T*>(q.ga) + pix * p.a_cstride + p.a_coffset + cv * 8, t);
    if (q.mask2) {
      const uint32_t m2 = q.mask2[pix * (p.mask2_cstride >> 3) + (p.mask2_coffset >> 3) + cv];
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = ((m2 >> i) & 1u) ? t[i] * s2 : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] += t[i];
  }
  if (q.gp) {
    const long pp = (static_cast<long>(n) * (p.h >> 1) + (hh >> 1)) * (p.w >> 1) + (ww >> 1);
    float t[8];
    load8f(reinterpret_cast<const T*>(q.gp) + pp * p.c + cv * 8, t);
    const uint2 am = *reinterpret_cast<const uint2*>(q.argmax + pp * p.c + cv * 8);
    const uint32_t code = ((hh & 1) << 1) | (ww & 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t a8 = ((i < 4 ? am.x : am.y) >> (8 * (i & 3))) & 0xFFu;
      if (a8 == code) g[i] += t[i];
    }
  }
  dlogit = 0.f;
  if (q.grad_out) {
    if (hh < p.h0 && ww < p.w0) {
      const long op = (static_cast<long>(n) * p.h0 + hh) * p.w0 + ww;
      const float o = q.out[op];
      dlogit = q.grad_out[op] * o * (1.f - o);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] += dlogit * wh[i];
  }
  const uint32_t m1 = q.mask1 ? q.mask1[pix * (p.c >> 3) + cv] : 0xFFu;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float z = fmaf(yv[i], a[i], b[i]);                    // s1 * (gamma * xhat + beta): same sign as the pre-activation
    const bool keep = (m1 >> i) & 1u;
    const bool pass = keep && (!p.relu || z > 0.f);
    act[i] = keep ? (p.relu ? fmaxf(z, 0.f) : z) : 0.f;
    dz[i] = pass ? g[i] * s1 : 0.f;
    xhat[i] = (yv[i] - mean) * rstd;
  }
}

// grid = (rows, n); thread owns channel vector t % cvs; deterministic block reduction to partials[n][row][c][3]
template <typename T>
__global__ void __launch_bounds__(256) unit_bwd_stats_kernel(UnitBwdParams p, UnitBwdPtrs q) {
  extern __shared__ float sm[];
  const int n = blockIdx.y;
  const int cvs = p.c >> 3;
  const int cv = threadIdx.x % cvs;
  const int slot = threadIdx.x / cvs;
  const int slots = blockDim.x / cvs;
  const int gsize = p.c / p.num_groups;
  float a[8], b[8], wh[8];
  {
    const float4* cp = reinterpret_cast<const float4*>(q.coef + static_cast<size_t>(n) * p.c + cv * 8);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = __ldg(cp + i);
      a[2 * i] = t.x; b[2 * i] = t.y; a[2 * i + 1] = t.z; b[2 * i + 1] = t.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) wh[i] = q.w_head ? __ldg(q.w_head + cv * 8 + i) : 0.f;
  }
  // all 8 channels of a vector lie in one group when gsize >= 8; otherwise per-channel (mean, rstd)
  float s1 = 1.f, s2 = 1.f;
  if (q.keep1) s1 = static_cast<float>(p.numel_per_call1 / static_cast<double>(q.keep1[n / p.images_per_call1]));
  if (q.mask2) s2 = static_cast<float>(p.numel_per_call2 / static_cast<double>(q.keep2[n / p.images_per_call2]));
  float mean8[8], rstd8[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 t = __ldg(q.mr + static_cast<size_t>(n) * p.num_groups + (cv * 8 + i) / gsize);
    mean8[i] = t.x;
    rstd8[i] = t.y;
  }
  float acc1[8], acc2[8], acc3[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc1[i] = acc2[i] = acc3[i] = 0.f;
  const int npix = p.h * p.w;
  for (int pix = blockIdx.x * slots + slot; pix < npix; pix += gridDim.x * slots) {
    const int hh = pix / p.w, ww = pix - hh * p.w;
    float dz[8], xhat[8], act[8], dlogit;
    // per-channel mean/rstd: pass channel 0's and fix up below (gsize < 8 only happens for C = 64, 128)
    unit_grad8<T>(p, q, n, hh, ww, cv, a, b, 0.f, 1.f, s1, s2, wh, dz, xhat, act, dlogit);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = (xhat[i] - mean8[i]) * rstd8[i];         // xhat[] holds raw y here (mean 0, rstd 1 above)
      acc1[i] += dz[i];
      acc2[i] += dz[i] * xh;
      acc3[i] += dlogit * act[i];
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

// grid = (num_groups, n) for the per-group coefficients, then a second launch shape for dgamma/dbeta; kept as two
// kernels for clarity.
__global__ void bwd_group_coef_kernel(const float* __restrict__ partials, int rows, int c, int num_groups,
                                      const float* __restrict__ gamma, double count, float2* __restrict__ gcoef) {
  const int g = blockIdx.x, n = blockIdx.y;
  const int gsize = c / num_groups;
  const float* base = partials + static_cast<size_t>(n) * rows * c * 3;
  double s1 = 0.0, s2 = 0.0;
  for (int i = threadIdx.x; i < rows * gsize; i += blockDim.x) {
    const int r = i / gsize, k = i - r * gsize;
    const int ch = g * gsize + k;
    const float* pp = base + (static_cast<size_t>(r) * c + ch) * 3;
    const double gm = static_cast<double>(gamma[ch]);
    s1 += gm * static_cast<double>(pp[0]);
    s2 += gm * static_cast<double>(pp[1]);
  }
  __shared__ double sh[2][128];
  sh[0][threadIdx.x] = s1;
  sh[1][threadIdx.x] = s2;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) gcoef[static_cast<size_t>(n) * num_groups + g] = make_float2(static_cast<float>(sh[0][0] / count), static_cast<float>(sh[1][0] / count));
}

// one block per channel: dgamma[c] = sum_{n,rows} P2, dbeta[c] = sum P1, dw_head[c] = sum P3
__global__ void bwd_param_grad_kernel(const float* __restrict__ partials, int n, int rows, int c, float* __restrict__ dgamma,
                                      float* __restrict__ dbeta, float* __restrict__ dw_head) {
  const int ch = blockIdx.x;
  double s[3] = {0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < n * rows; i += blockDim.x) {
    const float* pp = partials + (static_cast<size_t>(i) * c + ch) * 3;
    s[0] += pp[0];
    s[1] += pp[1];
    s[2] += pp[2];
  }
  __shared__ double sh[3][128];
  for (int k = 0; k < 3; ++k) sh[k][threadIdx.x] = s[k];
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int k = 0; k < 3; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (dbeta) dbeta[ch] = static_cast<float>(sh[0][0]);
    if (dgamma) dgamma[ch] = static_cast<float>(sh[1][0]);
    if (dw_head) dw_head[ch] = static_cast<float>(sh[2][0]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) unit_bwd_apply_kernel(UnitBwdParams p, UnitBwdPtrs q) {
  const int n = blockIdx.y;
  const int cvs = p.c >> 3;
  const int cv = threadIdx.x % cvs;
  const int slot = threadIdx.x / cvs;
  const int slots = blockDim.x / cvs;
  const int gsize = p.c / p.num_groups;
  float a[8], b[8], wh[8], gm[8];
  {
    const float4* cp = reinterpret_cast<const float4*>(q.coef + static_cast<size_t>(n) * p.c + cv * 8);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = __ldg(cp + i);
      a[2 * i] = t.x; b[2 * i] = t.y; a[2 * i + 1] = t.z; b[2 * i + 1] = t.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      wh[i] = q.w_head ? __ldg(q.w_head + cv * 8 + i) : 0.f;
      gm[i] = __ldg(q.gamma + cv * 8 + i);
    }
  }
  float s1 = 1.f, s2 = 1.f;
  if (q.keep1) s1 = static_cast<float>(p.numel_per_call1 / static_cast<double>(q.keep1[n / p.images_per_call1]));
  if (q.mask2) s2 = static_cast<float>(p.numel_per_call2 / static_cast<double>(q.keep2[n / p.images_per_call2]));
  float mean8[8], rstd8[8], c1[8], c2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = (cv * 8 + i) / gsize;
    const float2 t = __ldg(q.mr + static_cast<size_t>(n) * p.num_groups + g);
    const float2 u = __ldg(q.gcoef + static_cast<size_t>(n) * p.num_groups + g);
    mean8[i] = t.x; rstd8[i] = t.y; c1[i] = u.x; c2[i] = u.y;
  }
  const int npix = p.h * p.w;
  T* dy = reinterpret_cast<T*>(q.dy);
  for (int pix = blockIdx.x * slots + slot; pix < npix; pix += gridDim.x * slots) {
    const int hh = pix / p.w, ww = pix - hh * p.w;
    float dz[8], xhat[8], act[8], dlogit;
    unit_grad8<T>(p, q, n, hh, ww, cv, a, b, 0.f, 1.f, s1, s2, wh, dz, xhat, act, dlogit);
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = (xhat[i] - mean8[i]) * rstd8[i];
      o[i] = rstd8[i] * (gm[i] * dz[i] - c1[i] - xh * c2[i]);
    }
    long dst;
    if (p.s2d) {
      const long pp = (static_cast<long>(n) * (p.h >> 1) + (hh >> 1)) * (p.w >> 1) + (ww >> 1);
      dst = (pp * 4 + (((hh & 1) << 1) | (ww & 1))) * p.c + cv * 8;
    } else {
      dst = ((static_cast<long>(n) * p.h + hh) * p.w + ww) * p.c + cv * 8;
    }
    Vec8<T> v;
    v.from_float(o);
    v.store(dy + dst);
  }
}

static int bwd_pick_threads(int cvs) { return cvs > 256 ? 0 : (256 / cvs) * cvs; }

static int bwd_rows(int h, int w, int c) {
  const int threads = bwd_pick_threads(c / 8);
  const int slots = threads / (c / 8);
  long b = (static_cast<long>(h) * w + slots - 1) / slots;
  long cap = static_cast<long>(b2u_num_sms()) * 4;
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
  B2U_REQUIRE(d->grad_a || d->grad_pool || d->grad_out, "no upstream gradient source");
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
  if (d->dtype == B2U_F32) unit_bwd_stats_kernel<float><<<grid, threads, smem, st>>>(p, q);
  else unit_bwd_stats_kernel<__nv_bfloat16><<<grid, threads, smem, st>>>(p, q);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_unit_bwd_finalize(const float* partials, int n, int rows_per_image, int c, int num_groups, const float* gamma,
                                     double count, float* group_coef, float* dgamma, float* dbeta, float* dw_head, void* stream) {
  B2U_REQUIRE(partials && gamma && group_coef && n > 0 && c > 0 && num_groups > 0 && c % num_groups == 0, "bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  bwd_group_coef_kernel<<<dim3(num_groups, n), 128, 0, st>>>(partials, rows_per_image, c, num_groups, gamma, count,
                                                            reinterpret_cast<float2*>(group_coef));
  B2U_LAUNCH_CHECK();
  if (dgamma || dbeta || dw_head) {
    bwd_param_grad_kernel<<<c, 128, 0, st>>>(partials, n, rows_per_image, c, dgamma, dbeta, dw_head);
    B2U_LAUNCH_CHECK();
  }
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
  long bpi = (static_cast<long>(d->h) * d->w + slots - 1) / slots;
  const long cap = (static_cast<long>(b2u_num_sms()) * 16 + d->n - 1) / d->n;
  if (bpi > cap) bpi = cap;
  dim3 grid(static_cast<unsigned>(bpi), d->n);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (d->dtype == B2U_F32) unit_bwd_apply_kernel<float><<<grid, threads, 0, st>>>(p, q);
  else unit_bwd_apply_kernel<__nv_bfloat16><<<grid, threads, 0, st>>>(p, q);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}
