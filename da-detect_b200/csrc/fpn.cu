// FPN glue kernels (maskrcnn_benchmark/modeling/backbone/fpn.py:43-85), NHWC, HBM-bound, 128-bit accesses:
//   upsample2x          F.interpolate(x, scale_factor=2, mode="nearest")                      fpn.py:62
//   upsample2x backward 2x2 sum pooling of the gradient
//   subsample2          LastLevelMaxPool: F.max_pool2d(x, 1, 2, 0) = every second pixel       fpn.py:80-82
//   subsample2 backward scatter to the even pixels, zeros elsewhere
// The lateral add `inner_lateral + inner_top_down` (fpn.py:67) is not here: the up-sampled map is handed to the
// lateral 1x1 conv as its epilogue residual, so the sum never makes a separate pass over HBM.
#include "common.cuh"

namespace {

// out[n, y, x, :] = in[n, y/2, x/2, :]; one thread per 4 channels of an OUTPUT pixel.
__global__ void upsample2x_kernel(const float4* __restrict__ in, float4* __restrict__ out, int h, int w, int c4,
                                  long long total) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % c4);
    long long p = t / c4;
    const int x = (int)(p % (2 * w)); p /= 2 * w;
    const int y = (int)(p % (2 * h));
    const long long n = p / (2 * h);
    out[t] = __ldg(in + ((n * h + (y >> 1)) * w + (x >> 1)) * c4 + c);
  }
}

// gin[n, y, x, :] = sum of the 2x2 block of gout; one thread per 4 channels of an INPUT pixel.
__global__ void sumpool2x2_kernel(const float4* __restrict__ gout, float4* __restrict__ gin, int h, int w, int c4,
                                  long long total) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % c4);
    long long p = t / c4;
    const int x = (int)(p % w); p /= w;
    const int y = (int)(p % h);
    const long long n = p / h;
    const float4* r0 = gout + ((n * 2 * h + 2 * y) * 2 * w + 2 * x) * c4 + c;
    const float4* r1 = r0 + (long long)2 * w * c4;
    const float4 a = __ldg(r0), b = __ldg(r0 + c4), d = __ldg(r1), e = __ldg(r1 + c4);
    float4 s;                                            // (a + b) + (d + e): fixed order, deterministic
    s.x = (a.x + b.x) + (d.x + e.x); s.y = (a.y + b.y) + (d.y + e.y);
    s.z = (a.z + b.z) + (d.z + e.z); s.w = (a.w + b.w) + (d.w + e.w);
    gin[t] = s;
  }
}

// forward: out[n, y, x, :] = in[n, 2y, 2x, :] (oh = (h-1)/2+1); backward (scatter = 1): the transpose, zero fill.
__global__ void subsample2_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int h, int w, int oh,
                                  int ow, int c4, int scatter, long long total) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % c4);
    long long p = t / c4;
    if (!scatter) {                                       // t enumerates the small map
      const int x = (int)(p % ow); p /= ow;
      const int y = (int)(p % oh);
      const long long n = p / oh;
      dst[t] = __ldg(src + ((n * h + 2 * y) * w + 2 * x) * c4 + c);
    } else {                                              // t enumerates the large map
      const int x = (int)(p % w); p /= w;
      const int y = (int)(p % h);
      const long long n = p / h;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!(x & 1) && !(y & 1)) v = __ldg(src + ((n * oh + (y >> 1)) * ow + (x >> 1)) * c4 + c);
      dst[t] = v;
    }
  }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" int dd_upsample2x_forward(const float* x, float* y, int N, int H, int W, int C, void* stream) {
  DD_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && aligned16(x) && aligned16(y));
  const long long total = (long long)N * 2 * H * 2 * W * (C / 4);
  upsample2x_kernel<<<dd::grid_for(total, 256), 256, 0, dd::S(stream)>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), H, W, C / 4, total);
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_upsample2x_backward(const float* gy, float* gx, int N, int H, int W, int C, void* stream) {
  DD_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && aligned16(gy) && aligned16(gx));
  const long long total = (long long)N * H * W * (C / 4);
  sumpool2x2_kernel<<<dd::grid_for(total, 256), 256, 0, dd::S(stream)>>>(
      reinterpret_cast<const float4*>(gy), reinterpret_cast<float4*>(gx), H, W, C / 4, total);
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_subsample2_forward(const float* x, float* y, int N, int H, int W, int C, void* stream) {
  DD_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && aligned16(x) && aligned16(y));
  const int oh = (H - 1) / 2 + 1, ow = (W - 1) / 2 + 1;
  const long long total = (long long)N * oh * ow * (C / 4);
  subsample2_kernel<<<dd::grid_for(total, 256), 256, 0, dd::S(stream)>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), H, W, oh, ow, C / 4, 0, total);
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_subsample2_backward(const float* gy, float* gx, int N, int H, int W, int C, void* stream) {
  DD_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && aligned16(gy) && aligned16(gx));
  const int oh = (H - 1) / 2 + 1, ow = (W - 1) / 2 + 1;
  const long long total = (long long)N * H * W * (C / 4);
  subsample2_kernel<<<dd::grid_for(total, 256), 256, 0, dd::S(stream)>>>(
      reinterpret_cast<const float4*>(gy), reinterpret_cast<float4*>(gx), H, W, oh, ow, C / 4, 1, total);
  DD_LAUNCHED();
  return 0;
}
