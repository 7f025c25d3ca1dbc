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

// ---------------------------------------------------------------------------------------------------------------
// Staged variant (the one that runs for every ordinary resize ratio): the v1 kernel above spends its time on
// per-byte global loads (76 % excess L1 sectors under ncu), a runtime division per element and six IEEE divisions
// per pixel.  Here
//   * the source window of the tile is landed in shared memory by 16-byte loads (warp per row, lane per chunk),
//   * thread (column, row group) keeps its horizontal coefficients in registers and walks rows without any division,
//   * ToTensor + Normalize is a 3 x 256 entry table built once per CTA with the SAME correctly rounded fp32 steps
//     (bit-identical by construction), so the epilogue is three shared-memory reads per pixel.
// Shared memory: lut[3][256] f32 | hbuf[rows][TW*3] u8 | raw[rows][pitch] u8.
constexpr int kStagedThreads = 256;
constexpr int kRowGroups = kStagedThreads / TW;          // 4
constexpr int kRegTaps = 8;

struct StagedExtra { int pitch; long long total_bytes; };

__global__ void __launch_bounds__(kStagedThreads) preprocess_staged_kernel(const PreParams p, const StagedExtra e) {
  extern __shared__ __align__(16) uint8_t smem[];
  float* lut = reinterpret_cast<float*>(smem);
  uint8_t* hbuf = smem + 3 * 256 * sizeof(float);
  uint8_t* raw = hbuf + (size_t)p.max_rows * TW * 3;
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * p.th;
  const int tw = min(TW, p.Wp - x0), th = min(p.th, p.Hp - y0);
  const int vw = max(0, min(tw, p.out_w - x0)), vh = max(0, min(th, p.out_h - y0));
  const size_t plane = (size_t)p.Hp * p.Wp;
  const int c = tid & (TW - 1), rg = tid / TW;

  if (vw > 0 && vh > 0) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {                        // output plane j, source value tid
      float v = __fdiv_rn((float)tid, 255.f);
      if (p.bgr) v = __fmul_rn(v, 255.f);
      lut[j * 256 + tid] = __fdiv_rn(__fsub_rn(v, p.mean[j]), p.stdv[j]);
    }
    const int r0 = __ldg(p.yb + 2 * y0);
    const int yl = y0 + vh - 1;
    const int nrows = __ldg(p.yb + 2 * yl) + __ldg(p.yb + 2 * yl + 1) - r0;
    const int c_first = __ldg(p.xb + 2 * x0);
    const int xl = x0 + vw - 1;
    const int ncols = __ldg(p.xb + 2 * xl) + __ldg(p.xb + 2 * xl + 1) - c_first;
    // stage: row r of the window starts at source byte g0(r); its 16-byte aligned chunks go to raw[r][...]
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < nrows; r += kStagedThreads / 32) {
      const long long g0 = (long long)(r0 + r) * p.row_stride + (long long)c_first * p.pix_stride;
      const long long a0 = g0 & ~15LL;
      const int nch = (int)((g0 + (long long)ncols * p.pix_stride - a0 + 15) >> 4);
      uint8_t* dst = raw + (size_t)r * e.pitch;
      for (int j = lane; j < nch; j += 32) {
        const long long a = a0 + 16LL * j;
        if (a + 16 <= e.total_bytes) {
          *reinterpret_cast<uint4*>(dst + 16 * j) = __ldg(reinterpret_cast<const uint4*>(p.src + a));
        } else {                                          // the last chunk of the image: never read past its end
          for (int b = 0; b < 16; ++b) dst[16 * j + b] = a + b < e.total_bytes ? __ldg(p.src + a + b) : (uint8_t)0;
        }
      }
    }
    __syncthreads();
    // pass 1 (horizontal): thread = (output column c, row group rg)
    if (c < vw) {
      const int xo = x0 + c;
      const int xs = __ldg(p.xb + 2 * xo) - c_first, n = __ldg(p.xb + 2 * xo + 1);
      const int* k = p.xk + (size_t)xo * p.xks;
      int kw[kRegTaps];
#pragma unroll
      for (int i = 0; i < kRegTaps; ++i) kw[i] = i < n ? __ldg(k + i) : 0;
      for (int r = rg; r < nrows; r += kRowGroups) {
        const int off = (int)(((long long)(r0 + r) * p.row_stride + (long long)c_first * p.pix_stride) & 15);
        const uint8_t* s = raw + (size_t)r * e.pitch + off + xs * p.pix_stride;
        int a0 = 1 << (PRECISION_BITS - 1), a1 = a0, a2 = a0;
        if (n <= kRegTaps) {
#pragma unroll
          for (int i = 0; i < kRegTaps; ++i) {
            if (i < n) {
              a0 += (int)s[0] * kw[i]; a1 += (int)s[1] * kw[i]; a2 += (int)s[2] * kw[i];
              s += p.pix_stride;
            }
          }
        } else {
          for (int i = 0; i < n; ++i) {
            const int w = __ldg(k + i);
            a0 += (int)s[0] * w; a1 += (int)s[1] * w; a2 += (int)s[2] * w;
            s += p.pix_stride;
          }
        }
        uint8_t* h = hbuf + ((size_t)r * TW + c) * 3;
        h[0] = (uint8_t)clip8(a0); h[1] = (uint8_t)clip8(a1); h[2] = (uint8_t)clip8(a2);
      }
    }
    __syncthreads();
    // pass 2 (vertical) + flip + table look-up; a warp shares one output row, so its coefficients are uniform
    for (int r = rg; r < th; r += kRowGroups) {
      if (c >= tw) continue;
      const int yo = y0 + r, xo = x0 + c;
      float o0 = 0.f, o1 = 0.f, o2 = 0.f;
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
        const int v0 = clip8(a0), v1 = clip8(a1), v2 = clip8(a2);
        o0 = lut[p.bgr ? v2 : v0];
        o1 = lut[256 + v1];
        o2 = lut[512 + (p.bgr ? v0 : v2)];
        if (p.flip) xd = p.out_w - 1 - xo;
      }
      float* d = p.dst + (size_t)yo * p.Wp + xd;
      d[0] = o0; d[plane] = o1; d[2 * plane] = o2;
    }
  } else {                                                // a tile of pure padding
    for (int r = rg; r < th; r += kRowGroups) {
      if (c >= tw) continue;
      float* d = p.dst + (size_t)(y0 + r) * p.Wp + x0 + c;
      d[0] = 0.f; d[plane] = 0.f; d[2 * plane] = 0.f;
    }
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
  // Tile height: th output rows read at most ceil((th-1)*scale) + yksize + 1 source rows (and TW columns
  // ceil((TW-1)*xscale) + xksize + 1 source columns).
  const double yscale = (double)src_h / out_h, xscale = (double)src_w / out_w;
  const int max_cols = (int)ceil((TW - 1) * xscale) + xksize + 1;
  const int pitch = ((max_cols * pixel_stride + 15 + 15) / 16) * 16 + 16;
  // staged kernel: lut + hbuf + raw window must fit comfortably (several CTAs per SM); needs a 16-byte aligned source
  for (int th = 16; th >= 4 && (reinterpret_cast<uintptr_t>(src) & 15) == 0; th >>= 1) {
    const int rows = (int)ceil((th - 1) * yscale) + yksize + 1;
    const size_t smem = 3 * 256 * sizeof(float) + (size_t)rows * (TW * 3 + pitch);
    if (smem > 64 * 1024) continue;
    p.th = th; p.max_rows = rows;
    StagedExtra e = {pitch, (long long)(src_h - 1) * row_stride + (long long)src_w * pixel_stride};
    if (smem > 48 * 1024)
      DD_CUDA(cudaFuncSetAttribute(preprocess_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((Wp + TW - 1) / TW, (Hp + th - 1) / th);
    preprocess_staged_kernel<<<grid, kStagedThreads, smem, dd::S(stream)>>>(p, e);
    DD_LAUNCHED();
    return 0;
  }
  // extreme reductions (or an unaligned source): the direct kernel, intermediate rows only in shared memory
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
