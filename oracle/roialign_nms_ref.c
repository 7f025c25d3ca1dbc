/* TEST INFRASTRUCTURE ONLY — CPU restatement (plain C, scalar, single thread) of the
 * three native ops on the reference's training path.  Never linked into, imported by
 * or executed from the product path; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may use it.
 *
 * Pinned (tests/test_oracle_pins.py) against
 *   - the reference's own compiled CPU kernels (oracle/_ref/refcpu_C.so:
 *     csrc/cpu/ROIAlign_cpu.cpp, csrc/cpu/nms_cpu.cpp) — forward and `>=` NMS;
 *   - the reference's NMS golden vectors (tests/test_nms.py:11-58, :60-217);
 *   - torchvision.ops.roi_align(aligned=False) autograd for the backward, because
 *     the reference has no CPU backward (csrc/ROIAlign.h:44) and its CUDA file does
 *     not build against this torch (THC removed).
 *
 * Semantics restated from:
 *   RoIAlignForward            maskrcnn_benchmark/csrc/cuda/ROIAlign_cuda.cu:64-122
 *   bilinear_interpolate       ...ROIAlign_cuda.cu:15-62
 *   RoIAlignBackwardFeature    ...ROIAlign_cuda.cu:177-254 (gradient weights :125-175)
 *   nms (GPU: strict >)        maskrcnn_benchmark/csrc/cuda/nms.cu:13-21,52-63,112-123
 *   nms (CPU: >=)              maskrcnn_benchmark/csrc/cpu/nms_cpu.cpp:41-62
 * Layout here is the reference's: NCHW float32, rois [K,5] = (batch, x1, y1, x2, y2).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int yl, yh, xl, xh; float w1, w2, w3, w4; int valid; } tap4;

/* One bilinear sample point -> 4 taps + weights; valid=0 when the point is outside
 * [-1, size] (contributes 0).  ROIAlign_cuda.cu:22-57 / :133-172. */
static tap4 sample_taps(int height, int width, float y, float x) {
  tap4 t; memset(&t, 0, sizeof t);
  if (y < -1.0f || y > (float)height || x < -1.0f || x > (float)width) return t;
  if (y <= 0) y = 0;
  if (x <= 0) x = 0;
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= height - 1) { yh = yl = height - 1; y = (float)yl; } else yh = yl + 1;
  if (xl >= width - 1)  { xh = xl = width - 1;  x = (float)xl; } else xh = xl + 1;
  float ly = y - yl, lx = x - xl, hy = 1.0f - ly, hx = 1.0f - lx;
  t.yl = yl; t.yh = yh; t.xl = xl; t.xh = xh;
  t.w1 = hy * hx; t.w2 = hy * lx; t.w3 = ly * hx; t.w4 = ly * lx; t.valid = 1;
  return t;
}

typedef struct { float sw, sh, bw, bh; int gh, gw; int b; } roigeom;

static roigeom roi_geometry(const float* r, float scale, int ph, int pw, int sampling_ratio) {
  roigeom g;
  g.b = (int)r[0];
  g.sw = r[1] * scale; g.sh = r[2] * scale;
  float ew = r[3] * scale, eh = r[4] * scale;
  float rw = fmaxf(ew - g.sw, 1.0f), rh = fmaxf(eh - g.sh, 1.0f);   /* min ROI size 1 */
  g.bh = rh / (float)ph; g.bw = rw / (float)pw;
  g.gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)ph);
  g.gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)pw);
  return g;
}

void ref_roi_align_forward(const float* in, const float* rois, int K, int C, int H, int W,
                           float scale, int PH, int PW, int sampling_ratio, float* out) {
  for (int k = 0; k < K; ++k) {
    roigeom g = roi_geometry(rois + 5 * k, scale, PH, PW, sampling_ratio);
    float count = (float)(g.gh * g.gw);
    for (int c = 0; c < C; ++c) {
      const float* plane = in + ((size_t)g.b * C + c) * H * W;
      for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw) {
          float acc = 0.f;
          for (int iy = 0; iy < g.gh; ++iy) {
            float y = g.sh + ph * g.bh + ((float)iy + .5f) * g.bh / (float)g.gh;
            for (int ix = 0; ix < g.gw; ++ix) {
              float x = g.sw + pw * g.bw + ((float)ix + .5f) * g.bw / (float)g.gw;
              tap4 t = sample_taps(H, W, y, x);
              if (!t.valid) continue;
              acc += t.w1 * plane[t.yl * W + t.xl] + t.w2 * plane[t.yl * W + t.xh] +
                     t.w3 * plane[t.yh * W + t.xl] + t.w4 * plane[t.yh * W + t.xh];
            }
          }
          out[(((size_t)k * C + c) * PH + ph) * PW + pw] = acc / count;
        }
    }
  }
}

/* grad_in [N,C,H,W] must be zero-initialised by the caller (ROIAlign_cuda.cu:316). */
void ref_roi_align_backward(const float* gout, const float* rois, int K, int C, int H, int W,
                            float scale, int PH, int PW, int sampling_ratio, float* gin) {
  for (int k = 0; k < K; ++k) {
    roigeom g = roi_geometry(rois + 5 * k, scale, PH, PW, sampling_ratio);
    float count = (float)(g.gh * g.gw);
    for (int c = 0; c < C; ++c) {
      float* plane = gin + ((size_t)g.b * C + c) * H * W;
      for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw) {
          float go = gout[(((size_t)k * C + c) * PH + ph) * PW + pw];
          for (int iy = 0; iy < g.gh; ++iy) {
            float y = g.sh + ph * g.bh + ((float)iy + .5f) * g.bh / (float)g.gh;
            for (int ix = 0; ix < g.gw; ++ix) {
              float x = g.sw + pw * g.bw + ((float)ix + .5f) * g.bw / (float)g.gw;
              tap4 t = sample_taps(H, W, y, x);
              if (!t.valid) continue;
              plane[t.yl * W + t.xl] += go * t.w1 / count;
              plane[t.yl * W + t.xh] += go * t.w2 / count;
              plane[t.yh * W + t.xl] += go * t.w3 / count;
              plane[t.yh * W + t.xh] += go * t.w4 / count;
            }
          }
        }
    }
  }
}

static float iou_plus1(const float* a, const float* b) {
  float l = fmaxf(a[0], b[0]), r = fminf(a[2], b[2]);
  float t = fmaxf(a[1], b[1]), d = fminf(a[3], b[3]);
  float w = fmaxf(r - l + 1.f, 0.f), h = fmaxf(d - t + 1.f, 0.f);
  float inter = w * h;
  float sa = (a[2] - a[0] + 1.f) * (a[3] - a[1] + 1.f);
  float sb = (b[2] - b[0] + 1.f) * (b[3] - b[1] + 1.f);
  return inter / (sa + sb - inter);
}

/* Greedy NMS over boxes visited in `order` (score-descending indices supplied by the
 * caller so that tie order is explicit).  strict=1: suppress when IoU > thresh (GPU
 * reference); strict=0: IoU >= thresh (CPU reference).  Writes kept ORIGINAL indices in
 * ascending order (nms.cu:127-130 / nms_cpu.cpp:64) and returns how many. */
int ref_nms(const float* boxes, const int64_t* order, int n, float thresh, int strict,
            int64_t* keep_out) {
  unsigned char* dead = (unsigned char*)calloc((size_t)(n > 0 ? n : 1), 1);
  for (int i = 0; i < n; ++i) {
    int64_t bi = order[i];
    if (dead[bi]) continue;
    for (int j = i + 1; j < n; ++j) {
      int64_t bj = order[j];
      if (dead[bj]) continue;
      float o = iou_plus1(boxes + 4 * bi, boxes + 4 * bj);
      if (strict ? (o > thresh) : (o >= thresh)) dead[bj] = 1;
    }
  }
  int m = 0;
  for (int i = 0; i < n; ++i) if (!dead[i]) keep_out[m++] = i;
  free(dead);
  return m;
}
