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
  extern __shared__ float4 sgt[];          // Mcap boxes, then Mcap areas, then Mcap per-CTA maxima
  float* sarea = reinterpret_cast<float*>(sgt + Mcap);
  int* sbest = reinterpret_cast<int*>(sarea + Mcap);
  const int M = live_rows(Mcap, m_dev);
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const float4 g = gt[i];
    sgt[i] = g;
    sarea[i] = area_plus1(g);
    sbest[i] = 0;
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
      if ((threadIdx.x & 31) == 0) atomicMax(sbest + i, __float_as_int(m));     // per CTA first: 480 CTAs x 8 warps
    }                                                                            // on M addresses serialise in L2
  }
  if (gt_best != nullptr) {
    __syncthreads();
    for (int i = threadIdx.x; i < M; i += blockDim.x) atomicMax(reinterpret_cast<int*>(gt_best) + i, sbest[i]);
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

// BoxCoder.encode of one (ground truth, proposal) pair in the reference's operation order (box_coder.py:22-52)
__device__ __forceinline__ float4 encode_ref_order(const float4 g, const float4 p, float wx, float wy, float ww,
                                                   float wh) {
  const float ew = __fadd_rn(__fsub_rn(p.z, p.x), 1.f), eh = __fadd_rn(__fsub_rn(p.w, p.y), 1.f);
  const float ex = __fadd_rn(p.x, __fmul_rn(0.5f, ew)), ey = __fadd_rn(p.y, __fmul_rn(0.5f, eh));
  const float gw = __fadd_rn(__fsub_rn(g.z, g.x), 1.f), gh = __fadd_rn(__fsub_rn(g.w, g.y), 1.f);
  const float gx = __fadd_rn(g.x, __fmul_rn(0.5f, gw)), gy = __fadd_rn(g.y, __fmul_rn(0.5f, gh));
  float4 t;
  t.x = __fdiv_rn(__fmul_rn(wx, __fsub_rn(gx, ex)), ew);
  t.y = __fdiv_rn(__fmul_rn(wy, __fsub_rn(gy, ey)), eh);
  t.z = __fmul_rn(ww, logf(__fdiv_rn(gw, ew)));
  t.w = __fmul_rn(wh, logf(__fdiv_rn(gh, eh)));
  return t;
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
  out[j] = encode_ref_order(gt[m], pred[j], wx, wy, ww, wh);
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


// RPN anchor labels from the Matcher's result and the visibility mask (rpn/loss.py:57-89): 1 = matched, 0 = below the
// low threshold, -1 = between the thresholds or straddling the image border (ignored by the sampler).
__global__ void rpn_anchor_labels_kernel(const int64_t* __restrict__ matches, const uint8_t* __restrict__ vis, int N,
                                         int32_t* __restrict__ labels) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const long long m = matches[j];
  int lab = m >= 0 ? 1 : 0;
  if (m == -2 || !vis[j]) lab = -1;
  labels[j] = lab;
}

// RPNLossComputation's loss tail (rpn/loss.py:118-141) over the sampled anchors of all source images, forward and
// gradient in one launch of one CTA: BCE-with-logits over the sampled anchors and smooth-L1 (beta) over the sampled
// positives against BoxCoder(1,1,1,1) targets, both divided by the number of sampled anchors.  sel [S,B] / counts [S,2]
// come from dd_balanced_sample; row s*A + a of logits / deltas belongs to anchor a of source image s (source images
// come first in a batch).  dlogits / ddeltas (dense, zero-filled by the caller) receive d loss / d input.
__global__ void __launch_bounds__(256) rpn_sampled_losses_kernel(
    const float* __restrict__ logits, const float4* __restrict__ deltas, const float4* __restrict__ anchors, int A,
    int S, int B, const int64_t* __restrict__ sel, const int32_t* __restrict__ counts,
    const int32_t* __restrict__ labels, const int64_t* __restrict__ matches, const float4* __restrict__ gt_cat,
    const int32_t* __restrict__ gt_off, const int32_t* __restrict__ src_img, float beta, float* __restrict__ losses,
    float* __restrict__ dlogits, float4* __restrict__ ddeltas) {
  __shared__ float red[32];
  int total = 0;
  for (int s = 0; s < S; ++s) total += min(max(counts[2 * s + 1], 0), B);
  const float inv = 1.0f / (float)total;
  float bce = 0.f, l1 = 0.f;
  for (int idx = threadIdx.x; idx < S * B; idx += blockDim.x) {
    const int s = idx / B, r = idx - s * B;
    if (r >= counts[2 * s + 1]) continue;
    const int a = (int)sel[idx];
    const size_t row = (size_t)s * A + a;
    const int lab = labels[row];
    const float x = logits[row];
    const float t = lab == 1 ? 1.f : 0.f;
    bce += fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
    dlogits[row] = (1.0f / (1.0f + expf(-x)) - t) * inv;
    if (lab == 1) {
      long long m = matches[row];
      if (m < 0) m = 0;
      const float4 tg = encode_ref_order(gt_cat[gt_off[src_img[s]] + m], anchors[a], 1.f, 1.f, 1.f, 1.f);
      const float4 d4 = deltas[row];
      const float dv[4] = {d4.x - tg.x, d4.y - tg.y, d4.z - tg.z, d4.w - tg.w};
      float gv[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float d = dv[c], ad = fabsf(d);
        if (ad < beta) {
          l1 += 0.5f * ad * ad / beta;
          gv[c] = d / beta * inv;
        } else {
          l1 += ad - 0.5f * beta;
          gv[c] = (d > 0.f ? 1.f : -1.f) * inv;
        }
      }
      ddeltas[row] = make_float4(gv[0], gv[1], gv[2], gv[3]);
    }
  }
  bce = dd::block_sum(bce, red);
  l1 = dd::block_sum(l1, red);
  if (threadIdx.x == 0) {
    losses[0] = bce * inv;
    losses[1] = l1 * inv;
  }
}


// Box-head proposal labels of one image (box_head/loss.py:55-99): the matched ground-truth class, 0 below the low
// threshold, -1 (ignored by the sampler) between the thresholds and for buffer rows beyond the image's proposal
// count; every proposal of a target-domain image is background (:84-85).
__global__ void roi_labels_kernel(const int64_t* __restrict__ matches, const int64_t* __restrict__ gt_labels,
                                  int is_source, const int32_t* __restrict__ n_prop, int cap,
                                  int32_t* __restrict__ labels) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cap) return;
  const long long m = matches[j];
  int lab = 0;
  if (is_source) lab = m >= 0 ? (int)gt_labels[m] : (m == -2 ? -1 : 0);
  if (j >= *n_prop) lab = -1;
  labels[j] = lab;
}

// The sampled ROIs of a batch (box_head/loss.py:100-130 after the sampler): one thread per ROI slot gathers the
// proposal, its objectness, class label (0 in slots beyond the sampled count) and BoxCoder regression target
// (negative match indices wrap for target-domain images like the reference's tensor indexing, :47-51).
__global__ void roi_gather_sampled_kernel(
    const float4* __restrict__ boxes, const float* __restrict__ objectness, const int64_t* __restrict__ sel,
    const int32_t* __restrict__ counts, const int32_t* __restrict__ labels, const int64_t* __restrict__ matches,
    const float4* __restrict__ gt_cat, const int32_t* __restrict__ gt_off, const int32_t* __restrict__ gt_counts,
    const uint8_t* __restrict__ is_source, int n_img, int cap, int B, float wx, float wy, float ww, float wh,
    float* __restrict__ rois, int64_t* __restrict__ out_labels, float4* __restrict__ reg_targets,
    uint8_t* __restrict__ domain, uint8_t* __restrict__ valid, float* __restrict__ out_obj) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_img * B) return;
  const int img = t / B, r = t - img * B;
  const int a = (int)sel[t];
  const size_t row = (size_t)img * cap + a;
  const float4 bx = boxes[row];
  const bool ok = r < counts[2 * img + 1];
  rois[(size_t)t * 5 + 0] = (float)img;
  rois[(size_t)t * 5 + 1] = bx.x;
  rois[(size_t)t * 5 + 2] = bx.y;
  rois[(size_t)t * 5 + 3] = bx.z;
  rois[(size_t)t * 5 + 4] = bx.w;
  out_labels[t] = ok ? (int64_t)labels[row] : 0;
  const int g0 = gt_off[img], mcap = gt_off[img + 1] - g0;
  const int live = gt_counts ? min(max(gt_counts[img], 0), mcap) : mcap;
  const int M = max(live, 1);
  long long m = matches[row];
  if (m < 0) m = is_source[img] ? 0 : m + M;
  if (m < 0) m = 0;
  reg_targets[t] = encode_ref_order(gt_cat[g0 + m], bx, wx, wy, ww, wh);
  domain[t] = is_source[img];
  valid[t] = ok ? 1 : 0;
  out_obj[t] = objectness[row];
}

}  // namespace

extern "C" int dd_roi_labels(const int64_t* matches, const int64_t* gt_labels, int is_source, const int32_t* n_prop,
                             int cap, int32_t* labels, void* stream) {
  DD_CHECK_ARG(cap > 0);
  roi_labels_kernel<<<(cap + 255) / 256, 256, 0, dd::S(stream)>>>(matches, gt_labels, is_source, n_prop, cap, labels);
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_roi_gather_sampled(const float* boxes, const float* objectness, const int64_t* sel,
                                     const int32_t* counts, const int32_t* labels, const int64_t* matches,
                                     const float* gt_cat, const int32_t* gt_offsets, const int32_t* gt_counts,
                                     const uint8_t* is_source, int n_img, int cap, int B, float wx, float wy, float ww,
                                     float wh, float* rois, int64_t* out_labels, float* reg_targets, uint8_t* domain,
                                     uint8_t* valid, float* out_objectness, void* stream) {
  DD_CHECK_ARG(n_img > 0 && cap > 0 && B > 0);
  roi_gather_sampled_kernel<<<(n_img * B + 255) / 256, 256, 0, dd::S(stream)>>>(
      reinterpret_cast<const float4*>(boxes), objectness, sel, counts, labels, matches,
      reinterpret_cast<const float4*>(gt_cat), gt_offsets, gt_counts, is_source, n_img, cap, B, wx, wy, ww, wh, rois,
      out_labels, reinterpret_cast<float4*>(reg_targets), domain, valid, out_objectness);
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_rpn_anchor_labels(const int64_t* matches, const uint8_t* visibility, int N, int32_t* labels,
                                    void* stream) {
  DD_CHECK_ARG(N > 0);
  rpn_anchor_labels_kernel<<<(N + 255) / 256, 256, 0, dd::S(stream)>>>(matches, visibility, N, labels);
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_rpn_sampled_losses(const float* logits, const float* deltas, const float* anchors, int A, int S, int B,
                                     const int64_t* sel, const int32_t* counts, const int32_t* labels,
                                     const int64_t* matches, const float* gt_cat, const int32_t* gt_offsets,
                                     const int32_t* src_img, float beta, float* losses, float* dlogits, float* ddeltas,
                                     void* stream) {
  DD_CHECK_ARG(A > 0 && S > 0 && B > 0 && beta > 0.f);
  rpn_sampled_losses_kernel<<<1, 256, 0, dd::S(stream)>>>(
      logits, reinterpret_cast<const float4*>(deltas), reinterpret_cast<const float4*>(anchors), A, S, B, sel, counts,
      labels, matches, reinterpret_cast<const float4*>(gt_cat), gt_offsets, src_img, beta, losses, dlogits,
      reinterpret_cast<float4*>(ddeltas));
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_match(const float* gt, int M, const int32_t* m_dev, const float* pred, int N, float high, float low,
                        int allow_low_quality, int64_t* matches, float* matched_vals, float* gt_best, void* stream) {
  DD_CHECK_ARG(M > 0 && M <= kMaxGT && N > 0);   // the reference raises on empty GT / proposals (matcher.py:53-62)
  DD_CHECK_ARG(!allow_low_quality || gt_best != nullptr);
  cudaStream_t s = dd::S(stream);
  int32_t* argmax = nullptr;
  DD_CUDA(cudaMallocAsync(&argmax, sizeof(int32_t) * (size_t)N, s));
  if (allow_low_quality) DD_CUDA(cudaMemsetAsync(gt_best, 0, sizeof(float) * M, s));
  const int blocks = (N + 255) / 256;
  match_pass1<<<blocks, 256, M * 24, s>>>(reinterpret_cast<const float4*>(gt), M, m_dev,
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
