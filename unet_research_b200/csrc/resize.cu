// Fused `square_pad` + `TF.resize`.
//
// The multi-fidelity steps and the `-resize` Monte-Carlo runs pad an image to a square with zeros
// (unet_code/utils/utils_general.py:32-43: rows get top = d/2, bottom = d - d/2; columns get LEFT = d - d/2,
// right = d/2 -- the reference's left/right names are swapped) and then call torchvision's
// `TF.resize(tensor, (s, s))`, which for tensors is `F.interpolate(mode='bilinear', align_corners=False,
// antialias=True)` (torchvision 0.26).  This kernel restates ATen's anti-aliased bilinear
// filter (UpSampleBilinear2d.cu, upsample_gen2d_aa_out_frame): per output coordinate i
//     scale = in / out; support = max(scale, 1); center = scale * (i + 0.5)
//     xmin = max(int(center - support + 0.5), 0); xsize = min(int(center + support + 0.5), in) - xmin
//     w_j = max(0, 1 - |(j + xmin - center + 0.5) / max(scale, 1)|), normalised to sum 1
// applied along W for every contributing row, then along H, in fp32 -- and reads the zero padding on the fly, so the
// padded square is never materialised.  One thread per output pixel; x: [n, c, h, w] fp32 -> out: [n, c, oh, ow].
#include "b2u_common.cuh"

namespace b2u {

constexpr int kMaxTaps = 64;          // supports down-scaling factors up to ~31

__device__ __forceinline__ void aa_span(int i, int in_size, float scale, float support, int* xmin, int* xsize, float* center) {
  *center = scale * (static_cast<float>(i) + 0.5f);
  int lo = static_cast<int>(*center - support + 0.5f);
  if (lo < 0) lo = 0;
  int hi = static_cast<int>(*center + support + 0.5f);
  if (hi > in_size) hi = in_size;
  *xmin = lo;
  *xsize = hi - lo;
}

__device__ __forceinline__ float aa_weights(float* wt, int xmin, int xsize, float center, float scale) {
  const float invscale = scale >= 1.f ? 1.f / scale : 1.f;
  float total = 0.f;
  for (int j = 0; j < xsize; ++j) {
    float v = (static_cast<float>(j) + static_cast<float>(xmin) - center + 0.5f) * invscale;
    v = fabsf(v);
    const float w = v < 1.f ? 1.f - v : 0.f;
    wt[j] = w;
    total += w;
  }
  return total;
}

__global__ void __launch_bounds__(128) square_pad_resize_kernel(const float* __restrict__ x, float* __restrict__ out, int planes, int h, int w,
                                                                int pad_top, int pad_left, int sh, int sw, int oh, int ow) {
  const long total = static_cast<long>(planes) * oh * ow;
  const float scale_h = static_cast<float>(sh) / static_cast<float>(oh);
  const float scale_w = static_cast<float>(sw) / static_cast<float>(ow);
  const float support_h = scale_h >= 1.f ? scale_h : 1.f;
  const float support_w = scale_w >= 1.f ? scale_w : 1.f;
  float wy[kMaxTaps], wx[kMaxTaps];
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(idx % ow);
    const int oy = static_cast<int>((idx / ow) % oh);
    const int pl = static_cast<int>(idx / (static_cast<long>(ow) * oh));
    int ymin, ysize, xmin, xsize;
    float yc, xc;
    aa_span(oy, sh, scale_h, support_h, &ymin, &ysize, &yc);
    aa_span(ox, sw, scale_w, support_w, &xmin, &xsize, &xc);
    const float ty = aa_weights(wy, ymin, ysize, yc, scale_h);
    const float tx = aa_weights(wx, xmin, xsize, xc, scale_w);
    if (ty != 0.f) for (int j = 0; j < ysize; ++j) wy[j] /= ty;
    if (tx != 0.f) for (int j = 0; j < xsize; ++j) wx[j] /= tx;
    const float* src = x + static_cast<long>(pl) * h * w;
    float acc = 0.f;
    for (int j = 0; j < ysize; ++j) {
      const int yy = ymin + j - pad_top;                     // row in the un-padded image
      float row = 0.f;
      if (yy >= 0 && yy < h) {
        for (int i = 0; i < xsize; ++i) {
          const int xx = xmin + i - pad_left;
          const float v = (xx >= 0 && xx < w) ? __ldg(src + static_cast<long>(yy) * w + xx) : 0.f;
          row += v * wx[i];
        }
      }
      acc += row * wy[j];
    }
    out[idx] = acc;
  }
}

// Backward of the same operator (the multi-fidelity TRAINING steps resize the segmentation back up before the loss,
// MF-training-UNI.py:66-69, so the gradient flows through TF.resize): the adjoint of the separable filter,
//     grad_in[y][x] = sum_{oy, ox} Wy[oy][y + pad_top] * Wx[ox][x + pad_left] * grad_out[oy][ox],
// in gather form -- one thread per INPUT pixel walks the few output rows / columns whose support covers it (candidates
// from the inverse of the centre formula, membership and the normalised weight from the forward's own aa_span /
// aa_weights, so forward and backward use bit-identical weights).  Padding pixels receive no gradient (they are not
// part of the input).
constexpr int kMaxCand = 48;          // output coordinates that can touch one input coordinate (up-scaling factors up to ~11)

__device__ __forceinline__ int aa_adjoint(int pos, int in_size, int out_size, float scale, float support, int* first, float* wts) {
  // candidates: |scale * (o + 0.5) - pos| <= support + 1
  int lo = static_cast<int>(floorf((static_cast<float>(pos) - support - 1.f) / scale - 0.5f));
  int hi = static_cast<int>(ceilf((static_cast<float>(pos) + support + 1.f) / scale - 0.5f));
  if (lo < 0) lo = 0;
  if (hi > out_size - 1) hi = out_size - 1;
  *first = lo;
  int n = hi - lo + 1;
  if (n > kMaxCand) n = kMaxCand;
  float taps[kMaxTaps];
  for (int k = 0; k < n; ++k) {
    int xmin, xsize;
    float center;
    aa_span(lo + k, in_size, scale, support, &xmin, &xsize, &center);
    float w = 0.f;
    if (pos >= xmin && pos < xmin + xsize) {
      const float total = aa_weights(taps, xmin, xsize, center, scale);
      w = total != 0.f ? taps[pos - xmin] / total : taps[pos - xmin];
    }
    wts[k] = w;
  }
  return n > 0 ? n : 0;
}

__global__ void __launch_bounds__(128) square_pad_resize_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gin, int planes, int h,
                                                                    int w, int pad_top, int pad_left, int sh, int sw, int oh, int ow) {
  const long total = static_cast<long>(planes) * h * w;
  const float scale_h = static_cast<float>(sh) / static_cast<float>(oh);
  const float scale_w = static_cast<float>(sw) / static_cast<float>(ow);
  const float support_h = scale_h >= 1.f ? scale_h : 1.f;
  const float support_w = scale_w >= 1.f ? scale_w : 1.f;
  float wy[kMaxCand], wx[kMaxCand];
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(idx % w);
    const int y = static_cast<int>((idx / w) % h);
    const int pl = static_cast<int>(idx / (static_cast<long>(w) * h));
    int oy0, ox0;
    const int ny = aa_adjoint(y + pad_top, sh, oh, scale_h, support_h, &oy0, wy);
    const int nx = aa_adjoint(x + pad_left, sw, ow, scale_w, support_w, &ox0, wx);
    const float* g = gout + static_cast<long>(pl) * oh * ow;
    float acc = 0.f;
    for (int j = 0; j < ny; ++j) {
      if (wy[j] == 0.f) continue;
      float row = 0.f;
      for (int i = 0; i < nx; ++i) row += __ldg(g + static_cast<long>(oy0 + j) * ow + ox0 + i) * wx[i];
      acc += row * wy[j];
    }
    gin[idx] = acc;
  }
}

}  // namespace b2u

using namespace b2u;

extern "C" int b2u_square_pad_resize(const float* x, float* out, int planes, int h, int w, int square_pad, int oh, int ow,
                                     void* stream) {
  B2U_REQUIRE(x && out && planes > 0 && h > 0 && w > 0 && oh > 0 && ow > 0, "bad arguments");
  int sh = h, sw = w, top = 0, left = 0;
  if (square_pad) {
    const int size = h > w ? h : w;
    top = (size - h) / 2;                                    // utils_general.py:36-38
    const int tw = size - w;
    left = tw - tw / 2;                                      // :40-41 (the reference's `left` is total - total//2)
    sh = sw = size;
  }
  const float scale_h = static_cast<float>(sh) / oh, scale_w = static_cast<float>(sw) / ow;
  const float sup = fmaxf(fmaxf(scale_h, scale_w), 1.f);
  B2U_REQUIRE(2.f * sup + 2.f <= kMaxTaps, "down-scaling factor %.1f exceeds the %d-tap filter buffer", sup, kMaxTaps);
  const long total = static_cast<long>(planes) * oh * ow;
  long blocks = (total + 127) / 128;
  if (blocks > b2u_num_sms() * 16L) blocks = b2u_num_sms() * 16L;
  square_pad_resize_kernel<<<static_cast<int>(blocks), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, out, planes, h, w, top, left,
                                                                                                    sh, sw, oh, ow);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_square_pad_resize_bwd(const float* grad_out, float* grad_in, int planes, int h, int w, int square_pad, int oh, int ow,
                                         void* stream) {
  B2U_REQUIRE(grad_out && grad_in && planes > 0 && h > 0 && w > 0 && oh > 0 && ow > 0, "bad arguments");
  int sh = h, sw = w, top = 0, left = 0;
  if (square_pad) {
    const int size = h > w ? h : w;
    top = (size - h) / 2;
    const int tw = size - w;
    left = tw - tw / 2;
    sh = sw = size;
  }
  const float scale_h = static_cast<float>(sh) / oh, scale_w = static_cast<float>(sw) / ow;
  const float sup = fmaxf(fmaxf(scale_h, scale_w), 1.f);
  B2U_REQUIRE(2.f * sup + 2.f <= kMaxTaps, "down-scaling factor %.1f exceeds the %d-tap filter buffer", sup, kMaxTaps);
  const float smin = fminf(scale_h, scale_w);
  const float sup_min = fmaxf(smin, 1.f);
  B2U_REQUIRE((2.f * sup_min + 2.f) / smin + 3.f <= kMaxCand, "up-scaling factor %.1f exceeds the %d-candidate adjoint buffer", 1.f / smin, kMaxCand);
  const long total = static_cast<long>(planes) * h * w;
  long blocks = (total + 127) / 128;
  if (blocks > b2u_num_sms() * 16L) blocks = b2u_num_sms() * 16L;
  square_pad_resize_bwd_kernel<<<static_cast<int>(blocks), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(grad_out, grad_in, planes, h, w, top,
                                                                                                        left, sh, sw, oh, ow);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}
