// ROIAlign forward/backward, NHWC, for sm_100a.
//
// Semantics follow the reference op exactly (maskrcnn_benchmark/csrc/cuda/ROIAlign_cuda.cu:15-254):
// no half-pixel shift, ROI min size 1, adaptive sampling grid ceil(roi/pooled) when
// sampling_ratio <= 0, samples outside [-1, size] contribute 0, bilinear taps clamped to the map.
// Design differences (B200-first): channels are the contiguous dimension, one thread owns 4 channels
// of one output bin (128-bit loads/stores, tap geometry computed once per thread and shared by its
// 4 channels); a CTA covers one (roi, bin) pair across all channels so every global access of a
// warp is a fully coalesced 512 B segment.  `bin_step` = 2 evaluates only the even bins.
// Backward scatters with 128-bit vector reductions (red.global.add.v4.f32).
#include "common.cuh"

namespace {

struct RoiGeom {
  float start_w, start_h, bin_w, bin_h;
  int grid_h, grid_w, batch;
};

__device__ __forceinline__ RoiGeom roi_geom(const float* __restrict__ roi, float scale, int PH, int PW,
                                            int sampling_ratio) {
  RoiGeom g;
  g.batch = (int)roi[0];
  g.start_w = __fmul_rn(roi[1], scale);
  g.start_h = __fmul_rn(roi[2], scale);
  const float end_w = __fmul_rn(roi[3], scale), end_h = __fmul_rn(roi[4], scale);
  const float rw = fmaxf(__fsub_rn(end_w, g.start_w), 1.0f);
  const float rh = fmaxf(__fsub_rn(end_h, g.start_h), 1.0f);
  g.bin_h = __fdiv_rn(rh, (float)PH);
  g.bin_w = __fdiv_rn(rw, (float)PW);
  g.grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rh, (float)PH));
  g.grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rw, (float)PW));
  return g;
}

struct Taps {
  int o1, o2, o3, o4;  // pixel offsets (y*W + x)
  float w1, w2, w3, w4;
  bool valid;
};

__device__ __forceinline__ Taps make_taps(int H, int W, float y, float x) {
  Taps t;
  t.valid = !(y < -1.0f || y > (float)H || x < -1.0f || x > (float)W);
  if (!t.valid) return t;
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
  if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
  const float ly = __fsub_rn(y, (float)yl), lx = __fsub_rn(x, (float)xl);
  const float hy = __fsub_rn(1.0f, ly), hx = __fsub_rn(1.0f, lx);
  t.w1 = __fmul_rn(hy, hx); t.w2 = __fmul_rn(hy, lx); t.w3 = __fmul_rn(ly, hx); t.w4 = __fmul_rn(ly, lx);
  t.o1 = yl * W + xl; t.o2 = yl * W + xh; t.o3 = yh * W + xl; t.o4 = yh * W + xh;
  return t;
}

__device__ __forceinline__ float sample_coord(float start, int p, float bin, int i, int grid) {
  // start + p*bin + (i + .5)*bin/grid, evaluated left to right like the reference expression
  const float a = __fadd_rn(start, __fmul_rn((float)p, bin));
  const float b = __fdiv_rn(__fmul_rn((float)i + 0.5f, bin), (float)grid);
  return __fadd_rn(a, b);
}

// grid = (OPH*OPW, K); block = C/4 threads (capped at 256, looping over channel groups)
__global__ void __launch_bounds__(256) roi_align_fwd_nhwc(const float* __restrict__ feat,
                                                          const float* __restrict__ rois, float* __restrict__ out,
                                                          int H, int W, int C, float scale, int PH, int PW,
                                                          int sampling_ratio, int bin_step, int OPW,
                                                          const int* __restrict__ roi_level, int level) {
  const int k = blockIdx.y;
  if (roi_level != nullptr && __ldg(roi_level + k) != level) return;   // multi-level pooling: not this map's ROI
  const int ob = blockIdx.x;
  const int ph = (ob / OPW) * bin_step, pw = (ob % OPW) * bin_step;
  const RoiGeom g = roi_geom(rois + 5 * k, scale, PH, PW, sampling_ratio);
  const float inv_count = 1.0f / (float)(g.grid_h * g.grid_w);
  const float* base = feat + (size_t)g.batch * H * W * C;
  float* dst = out + ((size_t)k * gridDim.x + ob) * C;
  for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const float y = sample_coord(g.start_h, ph, g.bin_h, iy, g.grid_h);
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const float x = sample_coord(g.start_w, pw, g.bin_w, ix, g.grid_w);
        const Taps t = make_taps(H, W, y, x);
        if (!t.valid) continue;
        const float4 v1 = dd::ldg4(base + (size_t)t.o1 * C + c);
        const float4 v2 = dd::ldg4(base + (size_t)t.o2 * C + c);
        const float4 v3 = dd::ldg4(base + (size_t)t.o3 * C + c);
        const float4 v4 = dd::ldg4(base + (size_t)t.o4 * C + c);
        acc.x += t.w1 * v1.x + t.w2 * v2.x + t.w3 * v3.x + t.w4 * v4.x;
        acc.y += t.w1 * v1.y + t.w2 * v2.y + t.w3 * v3.y + t.w4 * v4.y;
        acc.z += t.w1 * v1.z + t.w2 * v2.z + t.w3 * v3.z + t.w4 * v4.z;
        acc.w += t.w1 * v1.w + t.w2 * v2.w + t.w3 * v3.w + t.w4 * v4.w;
      }
    }
    acc.x *= inv_count; acc.y *= inv_count; acc.z *= inv_count; acc.w *= inv_count;
    *reinterpret_cast<float4*>(dst + c) = acc;
  }
}

__device__ __forceinline__ void red_add4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__global__ void __launch_bounds__(256) roi_align_bwd_nhwc(const float* __restrict__ gout,
                                                          const float* __restrict__ rois, float* __restrict__ gfeat,
                                                          int H, int W, int C, float scale, int PH, int PW,
                                                          int sampling_ratio, int bin_step, int OPW,
                                                          const int* __restrict__ roi_level, int level) {
  const int k = blockIdx.y;
  if (roi_level != nullptr && __ldg(roi_level + k) != level) return;   // multi-level pooling: not this map's ROI
  const int ob = blockIdx.x;
  const int ph = (ob / OPW) * bin_step, pw = (ob % OPW) * bin_step;
  const RoiGeom g = roi_geom(rois + 5 * k, scale, PH, PW, sampling_ratio);
  const float count = (float)(g.grid_h * g.grid_w);
  float* base = gfeat + (size_t)g.batch * H * W * C;
  const float* src = gout + ((size_t)k * gridDim.x + ob) * C;
  for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4) {
    const float4 go = dd::ldg4(src + c);
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const float y = sample_coord(g.start_h, ph, g.bin_h, iy, g.grid_h);
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const float x = sample_coord(g.start_w, pw, g.bin_w, ix, g.grid_w);
        const Taps t = make_taps(H, W, y, x);
        if (!t.valid) continue;
        const float s1 = t.w1 / count, s2 = t.w2 / count, s3 = t.w3 / count, s4 = t.w4 / count;
        red_add4(base + (size_t)t.o1 * C + c, make_float4(go.x * s1, go.y * s1, go.z * s1, go.w * s1));
        red_add4(base + (size_t)t.o2 * C + c, make_float4(go.x * s2, go.y * s2, go.z * s2, go.w * s2));
        red_add4(base + (size_t)t.o3 * C + c, make_float4(go.x * s3, go.y * s3, go.z * s3, go.w * s3));
        red_add4(base + (size_t)t.o4 * C + c, make_float4(go.x * s4, go.y * s4, go.z * s4, go.w * s4));
      }
    }
  }
}

// Scalar-channel variants for C % 4 != 0 (tests with odd channel counts; never on the R-50 path).
__global__ void roi_align_fwd_nhwc_c1(const float* __restrict__ feat, const float* __restrict__ rois,
                                      float* __restrict__ out, int H, int W, int C, float scale, int PH, int PW,
                                      int sampling_ratio, int bin_step, int OPW, const int* __restrict__ roi_level,
                                      int level) {
  const int k = blockIdx.y, ob = blockIdx.x;
  if (roi_level != nullptr && __ldg(roi_level + k) != level) return;
  const int ph = (ob / OPW) * bin_step, pw = (ob % OPW) * bin_step;
  const RoiGeom g = roi_geom(rois + 5 * k, scale, PH, PW, sampling_ratio);
  const float inv_count = 1.0f / (float)(g.grid_h * g.grid_w);
  const float* base = feat + (size_t)g.batch * H * W * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const float y = sample_coord(g.start_h, ph, g.bin_h, iy, g.grid_h);
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const float x = sample_coord(g.start_w, pw, g.bin_w, ix, g.grid_w);
        const Taps t = make_taps(H, W, y, x);
        if (!t.valid) continue;
        acc += t.w1 * base[(size_t)t.o1 * C + c] + t.w2 * base[(size_t)t.o2 * C + c] +
               t.w3 * base[(size_t)t.o3 * C + c] + t.w4 * base[(size_t)t.o4 * C + c];
      }
    }
    out[((size_t)k * gridDim.x + ob) * C + c] = acc * inv_count;
  }
}

__global__ void roi_align_bwd_nhwc_c1(const float* __restrict__ gout, const float* __restrict__ rois,
                                      float* __restrict__ gfeat, int H, int W, int C, float scale, int PH, int PW,
                                      int sampling_ratio, int bin_step, int OPW, const int* __restrict__ roi_level,
                                      int level) {
  const int k = blockIdx.y, ob = blockIdx.x;
  if (roi_level != nullptr && __ldg(roi_level + k) != level) return;
  const int ph = (ob / OPW) * bin_step, pw = (ob % OPW) * bin_step;
  const RoiGeom g = roi_geom(rois + 5 * k, scale, PH, PW, sampling_ratio);
  const float count = (float)(g.grid_h * g.grid_w);
  float* base = gfeat + (size_t)g.batch * H * W * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float go = gout[((size_t)k * gridDim.x + ob) * C + c];
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const float y = sample_coord(g.start_h, ph, g.bin_h, iy, g.grid_h);
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const float x = sample_coord(g.start_w, pw, g.bin_w, ix, g.grid_w);
        const Taps t = make_taps(H, W, y, x);
        if (!t.valid) continue;
        atomicAdd(base + (size_t)t.o1 * C + c, go * t.w1 / count);
        atomicAdd(base + (size_t)t.o2 * C + c, go * t.w2 / count);
        atomicAdd(base + (size_t)t.o3 * C + c, go * t.w3 / count);
        atomicAdd(base + (size_t)t.o4 * C + c, go * t.w4 / count);
      }
    }
  }
}

// [planes, rows, cols] -> [planes, cols, rows] tiled transpose (NCHW <-> NHWC with rows/cols = C, H*W)
__global__ void transpose_tiles(const float* __restrict__ x, float* __restrict__ y, int rows, int cols) {
  __shared__ float tile[32][33];
  const size_t plane = (size_t)blockIdx.z * rows * cols;
  int r = blockIdx.y * 32 + threadIdx.y, c = blockIdx.x * 32 + threadIdx.x;
  for (int i = 0; i < 32; i += 8)
    if (r + i < rows && c < cols) tile[threadIdx.y + i][threadIdx.x] = x[plane + (size_t)(r + i) * cols + c];
  __syncthreads();
  r = blockIdx.x * 32 + threadIdx.y;
  c = blockIdx.y * 32 + threadIdx.x;
  for (int i = 0; i < 32; i += 8)
    if (r + i < cols && c < rows) y[plane + (size_t)(r + i) * rows + c] = tile[threadIdx.x][threadIdx.y + i];
}

int transpose_launch(const float* x, float* y, int planes, int rows, int cols, cudaStream_t s) {
  if (planes == 0 || rows == 0 || cols == 0) return 0;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, planes), block(32, 8);
  DD_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535);
  transpose_tiles<<<grid, block, 0, s>>>(x, y, rows, cols);
  DD_LAUNCHED();
  return 0;
}

}  // namespace

extern "C" int dd_nchw_to_nhwc(const float* x, float* y, int N, int C, int H, int W, void* stream) {
  return transpose_launch(x, y, N, C, H * W, dd::S(stream));
}
extern "C" int dd_nhwc_to_nchw(const float* x, float* y, int N, int C, int H, int W, void* stream) {
  return transpose_launch(x, y, N, H * W, C, dd::S(stream));
}

namespace {

int roi_align_launch(bool fwd, const float* a, const float* rois, float* b, int N, int H, int W, int C, int K,
                     float spatial_scale, int PH, int PW, int sampling_ratio, int bin_step, const int* roi_level,
                     int level, cudaStream_t s) {
  DD_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && K >= 0 && PH > 0 && PW > 0 && bin_step >= 1);
  if (K == 0) return 0;
  const int OPH = (PH + bin_step - 1) / bin_step, OPW = (PW + bin_step - 1) / bin_step;
  DD_CHECK_ARG(K <= 65535);
  dim3 grid(OPH * OPW, K);
  if (C % 4 == 0) {
    int threads = C / 4 < 256 ? ((C / 4 + 31) / 32) * 32 : 256;
    if (fwd)
      roi_align_fwd_nhwc<<<grid, threads, 0, s>>>(a, rois, b, H, W, C, spatial_scale, PH, PW, sampling_ratio,
                                                  bin_step, OPW, roi_level, level);
    else
      roi_align_bwd_nhwc<<<grid, threads, 0, s>>>(a, rois, b, H, W, C, spatial_scale, PH, PW, sampling_ratio,
                                                  bin_step, OPW, roi_level, level);
  } else {
    if (fwd)
      roi_align_fwd_nhwc_c1<<<grid, 128, 0, s>>>(a, rois, b, H, W, C, spatial_scale, PH, PW, sampling_ratio,
                                                 bin_step, OPW, roi_level, level);
    else
      roi_align_bwd_nhwc_c1<<<grid, 128, 0, s>>>(a, rois, b, H, W, C, spatial_scale, PH, PW, sampling_ratio,
                                                 bin_step, OPW, roi_level, level);
  }
  DD_LAUNCHED();
  return 0;
}

// LevelMapper (poolers.py:11-42): floor(lvl0 + log2(sqrt(area) / s0 + eps)) clamped to [k_min, k_max], minus k_min;
// area with the +1 convention (bounding_box.py:227-230).  Correctly rounded fp32 steps, like the torch expression.
__global__ void fpn_level_map_kernel(const float* __restrict__ rois, int K, int k_min, int k_max, float s0, int lvl0,
                                     float eps, int* __restrict__ levels) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const float* r = rois + 5 * k;
  const float w = __fadd_rn(__fsub_rn(r[3], r[1]), 1.0f), h = __fadd_rn(__fsub_rn(r[4], r[2]), 1.0f);
  const float s = __fsqrt_rn(__fmul_rn(w, h));
  float t = floorf(__fadd_rn((float)lvl0, log2f(__fadd_rn(__fdiv_rn(s, s0), eps))));
  t = fminf(fmaxf(t, (float)k_min), (float)k_max);        // NaN (negative area) -> k_min, like torch.clamp's min first
  levels[k] = (int)t - k_min;
}

}  // namespace

extern "C" int dd_roi_align_forward(const float* feat, const float* rois, float* out, int N, int H, int W, int C,
                                    int K, float spatial_scale, int PH, int PW, int sampling_ratio, int bin_step,
                                    void* stream) {
  return roi_align_launch(true, feat, rois, out, N, H, W, C, K, spatial_scale, PH, PW, sampling_ratio, bin_step,
                          nullptr, 0, dd::S(stream));
}

extern "C" int dd_roi_align_backward(const float* grad_out, const float* rois, float* grad_feat, int N, int H,
                                     int W, int C, int K, float spatial_scale, int PH, int PW, int sampling_ratio,
                                     int bin_step, void* stream) {
  return roi_align_launch(false, grad_out, rois, grad_feat, N, H, W, C, K, spatial_scale, PH, PW, sampling_ratio,
                          bin_step, nullptr, 0, dd::S(stream));
}

extern "C" int dd_fpn_level_map(const float* rois, int K, int k_min, int k_max, float canonical_scale,
                                int canonical_level, float eps, int* levels, void* stream) {
  DD_CHECK_ARG(K >= 0 && k_min <= k_max && canonical_scale > 0.f);
  if (K == 0) return 0;
  fpn_level_map_kernel<<<(K + 255) / 256, 256, 0, dd::S(stream)>>>(rois, K, k_min, k_max, canonical_scale,
                                                                     canonical_level, eps, levels);
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_roi_align_level_forward(const float* feat, const float* rois, const int* roi_level, int level,
                                          float* out, int N, int H, int W, int C, int K, float spatial_scale,
                                          int PH, int PW, int sampling_ratio, void* stream) {
  DD_CHECK_ARG(roi_level != nullptr);
  return roi_align_launch(true, feat, rois, out, N, H, W, C, K, spatial_scale, PH, PW, sampling_ratio, 1, roi_level,
                          level, dd::S(stream));
}

extern "C" int dd_roi_align_level_backward(const float* grad_out, const float* rois, const int* roi_level, int level,
                                           float* grad_feat, int N, int H, int W, int C, int K, float spatial_scale,
                                           int PH, int PW, int sampling_ratio, void* stream) {
  DD_CHECK_ARG(roi_level != nullptr);
  return roi_align_launch(false, grad_out, rois, grad_feat, N, H, W, C, K, spatial_scale, PH, PW, sampling_ratio, 1,
                          roi_level, level, dd::S(stream));
}

extern "C" int dd_roi_align_forward_nchw(const float* feat, const float* rois, float* out, int N, int C, int H,
                                         int W, int K, float spatial_scale, int PH, int PW, int sampling_ratio,
                                         void* stream) {
  cudaStream_t s = dd::S(stream);
  float *f = nullptr, *o = nullptr;
  DD_CUDA(cudaMallocAsync(&f, sizeof(float) * (size_t)N * C * H * W, s));
  DD_CUDA(cudaMallocAsync(&o, sizeof(float) * ((size_t)K * C * PH * PW + 1), s));
  int rc = dd_nchw_to_nhwc(feat, f, N, C, H, W, stream);
  if (!rc) rc = dd_roi_align_forward(f, rois, o, N, H, W, C, K, spatial_scale, PH, PW, sampling_ratio, 1, stream);
  if (!rc) rc = dd_nhwc_to_nchw(o, out, K, C, PH, PW, stream);
  cudaFreeAsync(f, s);
  cudaFreeAsync(o, s);
  return rc;
}

extern "C" int dd_roi_align_backward_nchw(const float* grad_out, const float* rois, float* grad_feat, int N, int C,
                                          int H, int W, int K, float spatial_scale, int PH, int PW,
                                          int sampling_ratio, void* stream) {
  cudaStream_t s = dd::S(stream);
  float *g = nullptr, *f = nullptr;
  DD_CUDA(cudaMallocAsync(&g, sizeof(float) * ((size_t)K * C * PH * PW + 1), s));
  DD_CUDA(cudaMallocAsync(&f, sizeof(float) * (size_t)N * C * H * W, s));
  DD_CUDA(cudaMemsetAsync(f, 0, sizeof(float) * (size_t)N * C * H * W, s));
  int rc = dd_nchw_to_nhwc(grad_out, g, K, C, PH, PW, stream);
  if (!rc) rc = dd_roi_align_backward(g, rois, f, N, H, W, C, K, spatial_scale, PH, PW, sampling_ratio, 1, stream);
  if (!rc) rc = dd_nhwc_to_nchw(f, grad_feat, N, C, H, W, stream);
  cudaFreeAsync(g, s);
  cudaFreeAsync(f, s);
  return rc;
}
