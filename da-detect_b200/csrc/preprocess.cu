// Input pipeline of one training sample as ONE kernel: Resize (Pillow's antialiased bilinear resampling, integer
// arithmetic, bit-exact) -> RandomHorizontalFlip -> ToTensor -> Normalize(to_bgr255) -> zero-padded slot of the
// batch tensor.  Replaces maskrcnn_benchmark/data/transforms/transforms.py:35-98 (CPU, PIL) and the padding copy of
// structures/image_list.py:66-88.  HBM-bound byte work: every source byte is read once (plus tile halos, which stay
// in L1/L2), every output float is written once, the uint8 intermediate of the two-pass resampling lives in shared
// memory only.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;   // Pillow Resample.c
constexpr int TW = 64;                       // output tile width (pixels); the tile height is chosen by the host

struct PreParams {
  const uint8_t* src; int src_h, src_w, pix_stride; long long row_stride;
  const int* xb; const int* xk; int xks;     // [out_w,2] (first, count), [out_w,xks]
  const int* yb; const int* yk; int yks;     // [out_h,2], [out_h,yks]
  int out_h, out_w, flip, bgr;
  float mean[3], stdv[3];
  float* dst; int Hp, Wp;                    // dst = this image's [3,Hp,Wp] slot
  int th, max_rows;                          // tile height, shared-memory rows available
};

__device__ __forceinline__ int clip8(int acc) {
  acc >>= PRECISION_BITS;
  return acc < 0 ? 0 : (acc > 255 ? 255 : acc);
}

// grid = (ceil(Wp/TW), ceil(Hp/th)); block = 256 threads.
__global__ void __launch_bounds__(256) preprocess_kernel(const PreParams p) {
  extern __shared__ uint8_t hbuf[];          // [rows][TW][3]: horizontally resampled source rows of this tile
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * p.th;
  const int tw = min(TW, p.Wp - x0), th = min(p.th, p.Hp - y0);
  const int vw = max(0, min(tw, p.out_w - x0)), vh = max(0, min(th, p.out_h - y0));   // part inside the image
  const size_t plane = (size_t)p.Hp * p.Wp;

  int r0 = 0, nrows = 0;
  if (vw > 0 && vh > 0) {
    r0 = __ldg(p.yb + 2 * y0);
    const int yl = y0 + vh - 1;
    nrows = __ldg(p.yb + 2 * yl) + __ldg(p.yb + 2 * yl + 1) - r0;      // bounds are monotone in y
    // pass 1 (ImagingResampleHorizontal_8bpc): source rows r0..r0+nrows-1 -> hbuf
    for (int t = threadIdx.x; t < nrows * vw; t += blockDim.x) {
      const int r = t / vw, c = t - r * vw;
      const int xo = x0 + c;
      const int xs = __ldg(p.xb + 2 * xo), n = __ldg(p.xb + 2 * xo + 1);
      const int* k = p.xk + (size_t)xo * p.xks;
      const uint8_t* s = p.src + (size_t)(r0 + r) * p.row_stride + (size_t)xs * p.pix_stride;
      int a0 = 1 << (PRECISION_BITS - 1), a1 = a0, a2 = a0;
      for (int i = 0; i < n; ++i) {
        const int w = __ldg(k + i);
        a0 += (int)__ldg(s) * w; a1 += (int)__ldg(s + 1) * w; a2 += (int)__ldg(s + 2) * w;
        s += p.pix_stride;
      }
      uint8_t* h = hbuf + ((size_t)r * TW + c) * 3;
      h[0] = (uint8_t)clip8(a0); h[1] = (uint8_t)clip8(a1); h[2] = (uint8_t)clip8(a2);
    }
  }
  __syncthreads();
  // pass 2 (ImagingResampleVertical_8bpc) + flip + ToTensor + Normalize + padding; threads run along x.
  for (int t = threadIdx.x; t < th * tw; t += blockDim.x) {
    const int r = t / tw, c = t - r * tw;
    const int yo = y0 + r, xo = x0 + c;
    float o0 = 0.f, o1 = 0.f, o2 = 0.f;     // plane order of the output
    int xd = xo;
    if (r < vh && c < vw) {
      const int ys = __ldg(p.yb + 2 * yo) - r0, n = __ldg(p.yb + 2 * yo + 1);
      const int* k = p.yk + (size_t)yo * p.yks;
      const uint8_t* h = hbuf + ((size_t)ys * TW + c) * 3;
      int a0 = 1 << (PRECISION_BITS - 1), a1 = a0, a2 = a0;
      for (int i = 0; i < n; ++i) {
        const int w = __ldg(k + i);
        a0 += (int)h[0] * w; a1 += (int)h[1] * w; a2 += (int)h[2] * w;
        h += TW * 3;
      }
      // torch order of operations, every step a correctly rounded fp32 op: u8 -> f32, / 255, (* 255), - mean, / std
      float v[3] = {(float)clip8(a0), (float)clip8(a1), (float)clip8(a2)};
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        v[ch] = __fdiv_rn(v[ch], 255.f);
        if (p.bgr) v[ch] = __fmul_rn(v[ch], 255.f);
      }
      const float c0 = p.bgr ? v[2] : v[0], c2 = p.bgr ? v[0] : v[2];
      o0 = __fdiv_rn(__fsub_rn(c0, p.mean[0]), p.stdv[0]);
      o1 = __fdiv_rn(__fsub_rn(v[1], p.mean[1]), p.stdv[1]);
      o2 = __fdiv_rn(__fsub_rn(c2, p.mean[2]), p.stdv[2]);
      if (p.flip) xd = p.out_w - 1 - xo;
    }
    float* d = p.dst + (size_t)yo * p.Wp + xd;
    d[0] = o0; d[plane] = o1; d[2 * plane] = o2;
  }
}

double bilinear(double x) {
  if (x < 0.0) x = -x;
  return x < 1.0 ? 1.0 - x : 0.0;
}

}  // namespace

// Pillow Resample.c precompute_coeffs (whole-axis box, bilinear) + normalize_coeffs_8bpc.  HOST function.
extern "C" int dd_resample_ksize(int in_size, int out_size) {
  if (in_size <= 0 || out_size <= 0) return -1;
  double filterscale = (double)in_size / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  return (int)ceil(1.0 * filterscale) * 2 + 1;
}

extern "C" int dd_resample_coeffs(int in_size, int out_size, int* h_bounds, int* h_kk) {
  DD_CHECK_ARG(in_size > 0 && out_size > 0 && h_bounds != nullptr && h_kk != nullptr);
  const double scale = (double)in_size / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 1.0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  const double ss = 1.0 / filterscale;
  double* w = new double[ksize];
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      w[x] = bilinear((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    int* k = h_kk + (size_t)xx * ksize;
    for (int x = 0; x < ksize; ++x) {
      double v = 0.0;
      if (x < xmax) v = ww != 0.0 ? w[x] / ww : w[x];
      k[x] = v < 0.0 ? (int)(-0.5 + v * (1 << PRECISION_BITS)) : (int)(0.5 + v * (1 << PRECISION_BITS));
    }
    h_bounds[2 * xx] = xmin;
    h_bounds[2 * xx + 1] = xmax;
  }
  delete[] w;
  return 0;
}

extern "C" int dd_preprocess_image(const uint8_t* src, int src_h, int src_w, int pixel_stride, long long row_stride,
                                   const int* xbounds, const int* xk, int xksize, const int* ybounds, const int* yk,
                                   int yksize, int out_h, int out_w, int flip, int to_bgr255, const float* h_mean,
                                   const float* h_std, float* dst, int Hp, int Wp, void* stream) {
  DD_CHECK_ARG(src != nullptr && dst != nullptr && xbounds != nullptr && xk != nullptr && ybounds != nullptr &&
               yk != nullptr && h_mean != nullptr && h_std != nullptr);
  DD_CHECK_ARG(src_h > 0 && src_w > 0 && (pixel_stride == 3 || pixel_stride == 4) &&
               row_stride >= (long long)src_w * pixel_stride);
  DD_CHECK_ARG(out_h > 0 && out_w > 0 && Hp >= out_h && Wp >= out_w);
  DD_CHECK_ARG(xksize == dd_resample_ksize(src_w, out_w) && yksize == dd_resample_ksize(src_h, out_h));
  PreParams p = {};
  p.src = src; p.src_h = src_h; p.src_w = src_w; p.pix_stride = pixel_stride; p.row_stride = row_stride;
  p.xb = xbounds; p.xk = xk; p.xks = xksize; p.yb = ybounds; p.yk = yk; p.yks = yksize;
  p.out_h = out_h; p.out_w = out_w; p.flip = flip ? 1 : 0; p.bgr = to_bgr255 ? 1 : 0;
  for (int i = 0; i < 3; ++i) { p.mean[i] = h_mean[i]; p.stdv[i] = h_std[i]; }
  p.dst = dst; p.Hp = Hp; p.Wp = Wp;
  // Tile height: th output rows read at most ceil((th-1)*scale) + yksize source rows; keep the shared-memory
  // intermediate ([rows][TW][3] bytes) under 96 KB.
  const double yscale = (double)src_h / out_h;
  int th = 16;
  for (;; th >>= 1) {
    p.max_rows = (int)ceil((th - 1) * yscale) + yksize + 1;
    if ((size_t)p.max_rows * TW * 3 <= 96 * 1024 || th == 1) break;
  }
  const size_t smem = (size_t)p.max_rows * TW * 3;
  DD_CHECK_ARG(smem <= 200 * 1024);          // a >300x vertical reduction is not an input-pipeline case
  p.th = th;
  if (smem > 48 * 1024)
    DD_CUDA(cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((Wp + TW - 1) / TW, (Hp + th - 1) / th);
  preprocess_kernel<<<grid, 256, smem, dd::S(stream)>>>(p);
  DD_LAUNCHED();
  return 0;
}
