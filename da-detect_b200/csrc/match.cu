// Fused IoU + Matcher, and BoxCoder encode/decode.
//
// Reference semantics: structures/boxlist_ops.py:56-91 (IoU with '+1' extents), modeling/matcher.py:42-112
// (argmax over GT with first-index ties, thresholds -> -1 / -2, low-quality restore by float equality
// with the per-GT maximum), modeling/box_coder.py:22-95.  The [M,N] IoU matrix the reference
// materialises (three [M,N,2] temporaries at N = 122 880 anchors) never exists here: each thread owns
// one prediction and streams the GT boxes from shared memory; the per-GT maxima are reduced with warp
// shuffles + one atomicMax per warp.  All IoU arithmetic uses explicit round-to-nearest intrinsics in the
// reference's operation order so that equality tests and threshold comparisons are bit-exact.
#include "common.cuh"

namespace {

constexpr int kMaxGT = 1024;

__device__ __forceinline__ float iou_ref_order(const float4 g, const float ga, const float4 p, const float pa) {
  // lt = max(box1[:, None, :2], box2[:, :2]); rb = min(...); wh = (rb - lt + 1).clamp(min=0)
  const float ltx = fmaxf(g.x, p.x), lty = fmaxf(g.y, p.y);
  const float rbx = fminf(g.z, p.z), rby = fminf(g.w, p.w);
  const float w = fmaxf(__fadd_rn(__fsub_rn(rbx, ltx), 1.0f), 0.0f);
  const float h = fmaxf(__fadd_rn(__fsub_rn(rby, lty), 1.0f), 0.0f);
  const float inter = __fmul_rn(w, h);
  // iou = inter / (area1[:, None] + area2 - inter)
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(ga, pa), inter));
}

__device__ __forceinline__ float area_plus1(const float4 b) {
  return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.0f), __fadd_rn(__fsub_rn(b.w, b.y), 1.0f));
}

// Pass 1: per prediction argmax/max; per GT max (atomicMax on the int view: IoU >= 0).
// m_dev (optional): the number of valid GT rows lives on the device (GT padded to the capacity M, so that one
// captured step graph serves batches with any number of boxes); rows at and beyond it do not exist for the Matcher.
__device__ __forceinline__ int live_rows(int M, const int32_t* __restrict__ m_dev) {
  return m_dev ? max(min(__ldg(m_dev), M), 0) : M;
}

__global__ void __launch_bounds__(256) match_pass1(const float4* __restrict__ gt, int Mcap,
                                                   const int32_t* __restrict__ m_dev,
                                                   const float4* __restrict__ pred, int N, float high, float low,
                                                   int64_t* __restrict__ matches, int32_t* __restrict__ argmax_out,
                                                   float* __restrict__ vals_out, float* __restrict__ gt_best) {
  extern __shared__ float4 sgt[];          // Mcap boxes, then Mcap areas
  float* sarea = reinterpret_cast<float*>(sgt + Mcap);
  const int M = live_rows(Mcap, m_dev);
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const float4 g = gt[i];
    sgt[i] = g;
    sarea[i] = area_plus1(g);
  }
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = j < N;
  float4 p = make_float4(0.f, 0.f, -1.f, -1.f);
  if (live) p = pred[j];
  const float pa = area_plus1(p);
  float best = -1.0f;
  int arg = 0;
  for (int i = 0; i < M; ++i) {
    float v = live ? iou_ref_order(sgt[i], sarea[i], p, pa) : 0.0f;
    if (live && v > best) { best = v; arg = i; }
    if (gt_best != nullptr) {
      float m = v;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(gt_best) + i, __float_as_int(m));
    }
  }
  if (live) {
    long long m = arg;
    if (best < low) m = -1;
    else if (best < high) m = -2;
    matches[j] = m;
    argmax_out[j] = arg;
    if (vals_out) vals_out[j] = best;
  }
}

// Pass 2 (allow_low_quality_matches): a prediction whose IoU with some GT equals that GT's maximum gets
// its argmax back (matcher.py:92-112).
__global__ void __launch_bounds__(256) match_pass2(const float4* __restrict__ gt, int Mcap,
                                                   const int32_t* __restrict__ m_dev,
                                                   const float4* __restrict__ pred, int N,
                                                   const float* __restrict__ gt_best,
                                                   const int32_t* __restrict__ argmax_in,
                                                   int64_t* __restrict__ matches) {
  extern __shared__ float4 sgt[];
  float* sarea = reinterpret_cast<float*>(sgt + Mcap);
  float* sbest = sarea + Mcap;
  const int M = live_rows(Mcap, m_dev);
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const float4 g = gt[i];
    sgt[i] = g;
    sarea[i] = area_plus1(g);
    sbest[i] = gt_best[i];
  }
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const float4 p = pred[j];
  const float pa = area_plus1(p);
  bool restore = false;
  for (int i = 0; i < M; ++i) restore |= (iou_ref_order(sgt[i], sarea[i], p, pa) == sbest[i]);
  if (restore) matches[j] = argmax_in[j];
}

__global__ void box_encode_kernel(const float4* __restrict__ gt, int Mcap, const int32_t* __restrict__ m_dev,
                                  const float4* __restrict__ pred,
                                  const int64_t* __restrict__ matches, int N, float wx, float wy, float ww, float wh,
                                  int wrap_negative, float4* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const int M = max(live_rows(Mcap, m_dev), 1);
  long long m = matches[j];
  if (m < 0) m = wrap_negative ? m + M : 0;
  if (m < 0) m = 0;
  const float4 g = gt[m], p = pred[j];
  const float ew = __fadd_rn(__fsub_rn(p.z, p.x), 1.f), eh = __fadd_rn(__fsub_rn(p.w, p.y), 1.f);
  const float ex = __fadd_rn(p.x, __fmul_rn(0.5f, ew)), ey = __fadd_rn(p.y, __fmul_rn(0.5f, eh));
  const float gw = __fadd_rn(__fsub_rn(g.z, g.x), 1.f), gh = __fadd_rn(__fsub_rn(g.w, g.y), 1.f);
  const float gx = __fadd_rn(g.x, __fmul_rn(0.5f, gw)), gy = __fadd_rn(g.y, __fmul_rn(0.5f, gh));
  float4 t;
  t.x = __fdiv_rn(__fmul_rn(wx, __fsub_rn(gx, ex)), ew);
  t.y = __fdiv_rn(__fmul_rn(wy, __fsub_rn(gy, ey)), eh);
  t.z = __fmul_rn(ww, logf(__fdiv_rn(gw, ew)));
  t.w = __fmul_rn(wh, logf(__fdiv_rn(gh, eh)));
  out[j] = t;
}

__global__ void box_decode_kernel(const float* __restrict__ codes, const float4* __restrict__ boxes, int R, int k,
                                  float wx, float wy, float ww, float wh, float* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= R * k) return;
  const int r = t / k;
  const float4 a = boxes[r];
  const float4 d = reinterpret_cast<const float4*>(codes)[t];
  const float clip = 4.135166556742356f;
  const float w = __fadd_rn(__fsub_rn(a.z, a.x), 1.0f), h = __fadd_rn(__fsub_rn(a.w, a.y), 1.0f);
  const float cx = __fadd_rn(a.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(a.y, __fmul_rn(0.5f, h));
  const float dx = __fdiv_rn(d.x, wx), dy = __fdiv_rn(d.y, wy);
  const float dw = fminf(__fdiv_rn(d.z, ww), clip), dh = fminf(__fdiv_rn(d.w, wh), clip);
  const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
  const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
  float4 o;
  o.x = __fsub_rn(pcx, __fmul_rn(0.5f, pw));
  o.y = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
  o.z = __fsub_rn(__fadd_rn(pcx, __fmul_rn(0.5f, pw)), 1.0f);
  o.w = __fsub_rn(__fadd_rn(pcy, __fmul_rn(0.5f, ph)), 1.0f);
  reinterpret_cast<float4*>(out)[t] = o;
}

}  // namespace

extern "C" int dd_match(const float* gt, int M, const int32_t* m_dev, const float* pred, int N, float high, float low,
                        int allow_low_quality, int64_t* matches, float* matched_vals, float* gt_best, void* stream) {
  DD_CHECK_ARG(M > 0 && M <= kMaxGT && N > 0);   // the reference raises on empty GT / proposals (matcher.py:53-62)
  DD_CHECK_ARG(!allow_low_quality || gt_best != nullptr);
  cudaStream_t s = dd::S(stream);
  int32_t* argmax = nullptr;
  DD_CUDA(cudaMallocAsync(&argmax, sizeof(int32_t) * (size_t)N, s));
  if (allow_low_quality) DD_CUDA(cudaMemsetAsync(gt_best, 0, sizeof(float) * M, s));
  const int blocks = (N + 255) / 256;
  match_pass1<<<blocks, 256, M * 20, s>>>(reinterpret_cast<const float4*>(gt), M, m_dev,
                                          reinterpret_cast<const float4*>(pred), N, high, low, matches, argmax,
                                          matched_vals, allow_low_quality ? gt_best : nullptr);
  DD_LAUNCHED();
  if (allow_low_quality) {
    match_pass2<<<blocks, 256, M * 24, s>>>(reinterpret_cast<const float4*>(gt), M, m_dev,
                                            reinterpret_cast<const float4*>(pred), N, gt_best, argmax, matches);
    DD_LAUNCHED();
  }
  DD_CUDA(cudaFreeAsync(argmax, s));
  return 0;
}

extern "C" int dd_box_encode(const float* gt, int M, const int32_t* m_dev, const float* pred, const int64_t* matches,
                             int N, float wx,
                             float wy, float ww, float wh, int wrap_negative, float* targets, void* stream) {
  DD_CHECK_ARG(M > 0 && N >= 0);
  if (N == 0) return 0;
  box_encode_kernel<<<(N + 255) / 256, 256, 0, dd::S(stream)>>>(
      reinterpret_cast<const float4*>(gt), M, m_dev, reinterpret_cast<const float4*>(pred), matches, N, wx, wy, ww, wh,
      wrap_negative, reinterpret_cast<float4*>(targets));
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_box_decode(const float* codes, const float* boxes, int R, int k, float wx, float wy, float ww,
                             float wh, float* out, void* stream) {
  DD_CHECK_ARG(R >= 0 && k > 0);
  if (R == 0) return 0;
  box_decode_kernel<<<(R * k + 255) / 256, 256, 0, dd::S(stream)>>>(codes, reinterpret_cast<const float4*>(boxes), R,
                                                                    k, wx, wy, ww, wh, out);
  DD_LAUNCHED();
  return 0;
}
