// Loss reductions of the DA Faster R-CNN path, each producing the scalar loss AND dL/d(input) in the
// same pass (the reference spends ~25 tiny ATen kernels per loss).
//
// Reference semantics: F.binary_cross_entropy_with_logits (da_heads/loss.py:95-100,165-173;
// rpn/loss.py:139-141), F.cross_entropy + smooth-L1 (box_head/loss.py:200-219, layers/smooth_l1_loss.py),
// consistency_loss (layers/consistency_loss.py:3-27), nn.TripletMarginLoss(p=2) over the last
// dimension (da_heads/loss.py:198-200,220-222).  fp32 sums; block tree reductions, one float atomic per CTA.
#include "common.cuh"

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(256) bce_kernel(const float* __restrict__ x, const float* __restrict__ targets,
                                                  const uint8_t* __restrict__ seg_labels, long long seg_len,
                                                  long long n, float inv_n, float* __restrict__ loss,
                                                  float* __restrict__ grad) {
  __shared__ float red[32];
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float t = targets ? targets[i] : (float)seg_labels[i / seg_len];
    acc += fmaxf(v, 0.f) - v * t + log1pf(expf(-fabsf(v)));
    if (grad) grad[i] = (sigmoidf_(v) - t) * inv_n;
  }
  acc = dd::block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(loss, acc * inv_n);
}

__global__ void __launch_bounds__(256) smooth_l1_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                        long long n, float beta, float inv_div,
                                                        float* __restrict__ loss, float* __restrict__ grad) {
  __shared__ float red[32];
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = x[i] - t[i];
    const float a = fabsf(d);
    if (a < beta) {
      acc += 0.5f * a * a / beta;
      if (grad) grad[i] = d / beta * inv_div;
    } else {
      acc += a - 0.5f * beta;
      if (grad) grad[i] = (d > 0.f ? 1.f : -1.f) * inv_div;
    }
  }
  acc = dd::block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(loss, acc * inv_div);
}

// single CTA
__global__ void __launch_bounds__(256) softmax_ce_kernel(const float* __restrict__ logits,
                                                         const int64_t* __restrict__ labels,
                                                         const uint8_t* __restrict__ mask, int rows, int C,
                                                         float* __restrict__ loss, float* __restrict__ grad) {
  __shared__ float red[32];
  __shared__ float s_cnt;
  float cnt = 0.f;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) cnt += (!mask || mask[r]) ? 1.f : 0.f;
  cnt = dd::block_sum(cnt, red);
  if (threadIdx.x == 0) s_cnt = cnt;
  __syncthreads();
  const float inv = 1.0f / s_cnt;
  float acc = 0.f;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    const float* z = logits + (size_t)r * C;
    float* g = grad ? grad + (size_t)r * C : nullptr;
    if (mask && !mask[r]) {
      if (g) for (int c = 0; c < C; ++c) g[c] = 0.f;
      continue;
    }
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, z[c]);
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(z[c] - m);
    const float lse = m + logf(se);
    const int y = (int)labels[r];
    acc += lse - z[y];
    if (g) for (int c = 0; c < C; ++c) g[c] = (expf(z[c] - lse) - (c == y ? 1.f : 0.f)) * inv;
  }
  acc = dd::block_sum(acc, red);
  if (threadIdx.x == 0) *loss = acc * inv;
}

// single CTA
__global__ void __launch_bounds__(256) box_reg_kernel(const float* __restrict__ box_reg,
                                                      const float* __restrict__ targets,
                                                      const int64_t* __restrict__ labels,
                                                      const uint8_t* __restrict__ mask, int rows, int C,
                                                      float* __restrict__ loss, float* __restrict__ grad) {
  __shared__ float red[32];
  __shared__ float s_cnt;
  float cnt = 0.f;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) cnt += (!mask || mask[r]) ? 1.f : 0.f;
  cnt = dd::block_sum(cnt, red);
  if (threadIdx.x == 0) s_cnt = cnt;
  __syncthreads();
  const float inv = 1.0f / s_cnt;
  float acc = 0.f;
  const int rowlen = 4 * C;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    float* g = grad ? grad + (size_t)r * rowlen : nullptr;
    if (g) for (int c = 0; c < rowlen; ++c) g[c] = 0.f;
    const int y = (int)labels[r];
    if ((mask && !mask[r]) || y <= 0) continue;
    for (int c = 0; c < 4; ++c) {
      const float d = box_reg[(size_t)r * rowlen + 4 * y + c] - targets[(size_t)r * 4 + c];
      const float a = fabsf(d);
      if (a < 1.0f) {
        acc += 0.5f * a * a;
        if (g) g[4 * y + c] = d * inv;
      } else {
        acc += a - 0.5f;
        if (g) g[4 * y + c] = (d > 0.f ? 1.f : -1.f) * inv;
      }
    }
  }
  acc = dd::block_sum(acc, red);
  if (threadIdx.x == 0) *loss = acc * inv;
}

// ---- consistency: ws = [sum_sig0, sum_sig1, dmean0, dmean1]
__global__ void __launch_bounds__(256) cst_img_mean_kernel(const float* __restrict__ img_logits, long long hw,
                                                           float* __restrict__ ws) {
  __shared__ float red[32];
  const int img = blockIdx.y;
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x)
    acc += sigmoidf_(img_logits[(size_t)img * hw + i]);
  acc = dd::block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(ws + img, acc);
}

// single CTA over the K ROIs
__global__ void __launch_bounds__(256) cst_ins_kernel(const float* __restrict__ ins_logits, int K, int n_src,
                                                      const uint8_t* __restrict__ row_valid, float inv_hw,
                                                      float* __restrict__ ws, float* __restrict__ loss,
                                                      float* __restrict__ grad_ins) {
  __shared__ float red[32];
  __shared__ float s_cnt;
  const float mean0 = ws[0] * inv_hw, mean1 = ws[1] * inv_hw;
  float cnt = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) cnt += (!row_valid || row_valid[k]) ? 1.f : 0.f;
  cnt = dd::block_sum(cnt, red);
  if (threadIdx.x == 0) s_cnt = cnt;
  __syncthreads();
  const float invK = 1.0f / s_cnt;                 // mean over the ROIs that exist (padding rows are skipped)
  float acc = 0.f, d0 = 0.f, d1 = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    if (row_valid && !row_valid[k]) {
      if (grad_ins) grad_ins[k] = 0.f;
      continue;
    }
    const float s = sigmoidf_(ins_logits[k]);
    const bool src = k < n_src;
    const float diff = (src ? mean0 : mean1) - s;
    acc += fabsf(diff);
    const float sg = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
    if (grad_ins) grad_ins[k] = -sg * s * (1.f - s) * invK;
    if (src) d0 += sg * invK; else d1 += sg * invK;
  }
  acc = dd::block_sum(acc, red);
  if (threadIdx.x == 0) *loss = acc * invK;
  d0 = dd::block_sum(d0, red);
  if (threadIdx.x == 0) ws[2] = d0;
  d1 = dd::block_sum(d1, red);
  if (threadIdx.x == 0) ws[3] = d1;
}

__global__ void __launch_bounds__(256) cst_img_grad_kernel(const float* __restrict__ img_logits, long long hw,
                                                           float inv_hw, const float* __restrict__ ws,
                                                           float* __restrict__ grad_img) {
  const int img = blockIdx.y;
  const float dmean = ws[2 + img] * inv_hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    const float s = sigmoidf_(img_logits[(size_t)img * hw + i]);
    grad_img[(size_t)img * hw + i] = dmean * s * (1.f - s);
  }
}

// ---- triplet margin loss
__device__ __forceinline__ size_t tri_addr(long long r, int d, int D, long long inner) {
  return (size_t)(r / inner) * D * inner + (size_t)d * inner + (size_t)(r % inner);
}

// inner > 1: one thread per distance vector (threads along `inner` are coalesced)
__global__ void __launch_bounds__(256) triplet_strided_kernel(const float* __restrict__ a, const float* __restrict__ p,
                                                              const float* __restrict__ n, long long rows, int D,
                                                              long long inner, float margin_host,
                                                              const float* __restrict__ margin_dev, float inv_rows,
                                                              float* __restrict__ loss, float* __restrict__ ga,
                                                              float* __restrict__ gp, float* __restrict__ gn) {
  const float margin = margin_dev ? __ldg(margin_dev) : margin_host;
  __shared__ float red[32];
  float acc = 0.f;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    float sp = 0.f, sn = 0.f;
    for (int d = 0; d < D; ++d) {
      const size_t o = tri_addr(r, d, D, inner);
      const float av = a[o];
      const float dp = av - p[o] + 1e-6f, dn = av - n[o] + 1e-6f;
      sp += dp * dp;
      sn += dn * dn;
    }
    const float np_ = sqrtf(sp), nn_ = sqrtf(sn);
    const float h = margin + np_ - nn_;
    const bool on = h > 0.f;
    if (on) acc += h;
    if (ga) {
      for (int d = 0; d < D; ++d) {
        const size_t o = tri_addr(r, d, D, inner);
        const float av = a[o];
        const float dp = av - p[o] + 1e-6f, dn = av - n[o] + 1e-6f;
        const float up = on ? dp / np_ * inv_rows : 0.f, un = on ? dn / nn_ * inv_rows : 0.f;
        ga[o] = up - un;
        gp[o] = -up;
        gn[o] = un;
      }
    }
  }
  acc = dd::block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(loss, acc * inv_rows);
}

// inner == 1: one warp per row, lanes along D
__global__ void __launch_bounds__(256) triplet_rows_kernel(const float* __restrict__ a, const float* __restrict__ p,
                                                           const float* __restrict__ n, long long rows, int D,
                                                           float margin_host, const float* __restrict__ margin_dev,
                                                           float inv_rows, float* __restrict__ loss,
                                                           float* __restrict__ ga, float* __restrict__ gp,
                                                           float* __restrict__ gn) {
  const float margin = margin_dev ? __ldg(margin_dev) : margin_host;
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float acc = 0.f;
  for (long long r = warp; r < rows; r += nwarps) {
    const size_t base = (size_t)r * D;
    float sp = 0.f, sn = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float av = a[base + d];
      const float dp = av - p[base + d] + 1e-6f, dn = av - n[base + d] + 1e-6f;
      sp += dp * dp;
      sn += dn * dn;
    }
    sp = dd::warp_sum(sp);
    sn = dd::warp_sum(sn);
    const float np_ = sqrtf(sp), nn_ = sqrtf(sn);
    const float h = margin + np_ - nn_;
    const bool on = h > 0.f;
    if (on && lane == 0) acc += h;
    if (ga) {
      for (int d = lane; d < D; d += 32) {
        const float av = a[base + d];
        const float dp = av - p[base + d] + 1e-6f, dn = av - n[base + d] + 1e-6f;
        const float up = on ? dp / np_ * inv_rows : 0.f, un = on ? dn / nn_ * inv_rows : 0.f;
        ga[base + d] = up - un;
        gp[base + d] = -up;
        gn[base + d] = un;
      }
    }
  }
  acc = dd::warp_sum(acc);
  if (lane == 0 && acc != 0.f) atomicAdd(loss, acc * inv_rows);
}

}  // namespace

extern "C" int dd_bce_logits_mean(const float* x, const float* targets, const uint8_t* seg_labels, long long seg_len,
                                  long long n, float* loss, float* grad, void* stream) {
  DD_CHECK_ARG(n > 0 && (targets != nullptr || (seg_labels != nullptr && seg_len > 0)));
  cudaStream_t s = dd::S(stream);
  DD_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), s));
  bce_kernel<<<dd::grid_for(n, 256, 2), 256, 0, s>>>(x, targets, seg_labels, seg_len, n, 1.0f / (float)n, loss, grad);
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_smooth_l1_sum(const float* x, const float* t, long long n, float beta, float divisor, float* loss,
                                float* grad, void* stream) {
  DD_CHECK_ARG(n >= 0 && beta > 0.f && divisor != 0.f);
  cudaStream_t s = dd::S(stream);
  DD_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), s));
  if (n == 0) return 0;
  smooth_l1_kernel<<<dd::grid_for(n, 256, 2), 256, 0, s>>>(x, t, n, beta, 1.0f / divisor, loss, grad);
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_softmax_ce_mean(const float* logits, const int64_t* labels, const uint8_t* row_mask, int rows, int C,
                                  float* loss, float* grad, void* stream) {
  DD_CHECK_ARG(rows > 0 && C > 0);
  softmax_ce_kernel<<<1, 256, 0, dd::S(stream)>>>(logits, labels, row_mask, rows, C, loss, grad);
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_box_reg_loss(const float* box_reg, const float* reg_targets, const int64_t* labels,
                               const uint8_t* row_mask, int rows, int C, float* loss, float* grad, void* stream) {
  DD_CHECK_ARG(rows > 0 && C > 0);
  box_reg_kernel<<<1, 256, 0, dd::S(stream)>>>(box_reg, reg_targets, labels, row_mask, rows, C, loss, grad);
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_consistency_loss(const float* img_logits, long long hw, const float* ins_logits, int K, int n_src,
                                   const uint8_t* row_valid, float* loss, float* grad_img, float* grad_ins,
                                   float* workspace2, void* stream) {
  DD_CHECK_ARG(hw > 0 && K > 0 && n_src >= 0 && n_src <= K && workspace2 != nullptr);
  cudaStream_t s = dd::S(stream);
  DD_CUDA(cudaMemsetAsync(workspace2, 0, 4 * sizeof(float), s));
  dim3 grid(dd::grid_for(hw, 256, 1), 2);
  cst_img_mean_kernel<<<grid, 256, 0, s>>>(img_logits, hw, workspace2);
  DD_LAUNCHED();
  cst_ins_kernel<<<1, 256, 0, s>>>(ins_logits, K, n_src, row_valid, 1.0f / (float)hw, workspace2, loss, grad_ins);
  DD_LAUNCHED();
  if (grad_img) {
    cst_img_grad_kernel<<<grid, 256, 0, s>>>(img_logits, hw, 1.0f / (float)hw, workspace2, grad_img);
    DD_LAUNCHED();
  }
  return 0;
}

// The adaptive image-level margin of da_heads/loss.py:182-200 as device state (no host read of the previous loss):
//   if state == 0: state = margin_cfg;  if prev_loss == 0 and int(state) != int(max_margin): state += lr
// state is a double like the reference's Python float; margin_out receives the float the loss kernel consumes.
__global__ void adaptive_margin_kernel(double* state, const float* __restrict__ prev_loss, double margin_cfg, double lr,
                                       double max_margin, float* margin_out) {
  double m = *state;
  if (m == 0.0) m = margin_cfg;
  if (prev_loss != nullptr && *prev_loss == 0.0f && (long long)m != (long long)max_margin) m += lr;
  *state = m;
  *margin_out = (float)m;
}

extern "C" int dd_adaptive_margin_update(double* state, const float* prev_loss, double margin_cfg, double lr,
                                         double max_margin, float* margin_out, void* stream) {
  DD_CHECK_ARG(state != nullptr && margin_out != nullptr);
  adaptive_margin_kernel<<<1, 1, 0, dd::S(stream)>>>(state, prev_loss, margin_cfg, lr, max_margin, margin_out);
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_triplet_margin_loss(const float* a, const float* p, const float* n, long long rows, int D,
                                      long long inner, float margin, const float* margin_dev, float* loss,
                                      float* grad_a, float* grad_p, float* grad_n, void* stream) {
  DD_CHECK_ARG(rows > 0 && D > 0 && inner >= 1 && rows % inner == 0);
  DD_CHECK_ARG((grad_a == nullptr) == (grad_p == nullptr) && (grad_a == nullptr) == (grad_n == nullptr));
  cudaStream_t s = dd::S(stream);
  DD_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), s));
  const float inv_rows = 1.0f / (float)rows;
  if (inner > 1) {
    triplet_strided_kernel<<<dd::grid_for(rows, 256, 2), 256, 0, s>>>(a, p, n, rows, D, inner, margin, margin_dev, inv_rows, loss,
                                                                      grad_a, grad_p, grad_n);
  } else {
    triplet_rows_kernel<<<dd::grid_for(rows * 32, 256, 2), 256, 0, s>>>(a, p, n, rows, D, margin, margin_dev, inv_rows, loss,
                                                                        grad_a, grad_p, grad_n);
  }
  DD_LAUNCHED();
  return 0;
}
