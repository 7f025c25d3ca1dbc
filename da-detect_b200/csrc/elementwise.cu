// Bandwidth-bound helpers of the path: stem max-pool, 7x7 average pool (+backward), ReLU backward,
// gradient-reversal scale, dropout apply, SGD-momentum update.  All NHWC, 128-bit accesses where the
// channel count allows, grid-stride loops sized in multiples of the SM count.
#include "common.cuh"

namespace {

// F.max_pool2d(kernel=3, stride=2, padding=1) (resnet.py:335).  One thread per 4 output channels.
__global__ void maxpool3x3s2_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C,
                                    int OH, int OW) {
  const int c4n = C / 4;
  const long long total = (long long)N * OH * OW * c4n;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(t % c4n);
    const int ow = (int)((t / c4n) % OW);
    const int oh = (int)((t / ((long long)c4n * OW)) % OH);
    const int n = (int)(t / ((long long)c4n * OW * OH));
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int ih = oh * 2 - 1 + dy;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int iw = ow * 2 - 1 + dx;
        if (iw < 0 || iw >= W) continue;
        const float4 v = dd::ldg4(x + (((size_t)n * H + ih) * W + iw) * C + c4 * 4);
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    }
    *reinterpret_cast<float4*>(y + (((size_t)n * OH + oh) * OW + ow) * C + c4 * 4) = m;
  }
}

__global__ void avgpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int K, int HW, int C) {
  const long long total = (long long)K * C;
  const float inv = 1.0f / (float)HW;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    const long long k = t / C;
    const float* p = x + (size_t)k * HW * C + c;
    float acc = 0.f;
    int i = 0;
    for (; i + 7 <= HW; i += 7) {                 // seven independent loads in flight, added in the original order
      float v[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) v[j] = __ldg(p + (size_t)(i + j) * C);
#pragma unroll
      for (int j = 0; j < 7; ++j) acc += v[j];
    }
    for (; i < HW; ++i) acc += __ldg(p + (size_t)i * C);
    y[t] = acc * inv;
  }
}

__global__ void avgpool_bwd_kernel(const float* __restrict__ gy, float* __restrict__ gx, int K, int HW, int C) {
  const long long total = (long long)K * HW * C;
  const float inv = 1.0f / (float)HW;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    const long long k = t / ((long long)HW * C);
    gx[t] = gy[k * C + c] * inv;
  }
}

// gx[k, i, c] = act[k, i, c] > 0 ? gy[k, c] / HW : 0   (AvgPool2d backward fused with the ReLU mask of its input)
__global__ void avgpool_relu_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ act,
                                        float* __restrict__ gx, int K, int HW, int C4) {
  const long long total = (long long)K * HW * C4;
  const float inv = 1.0f / (float)HW;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(t % C4);
    const long long k = t / ((long long)HW * C4);
    const float4 g = dd::ldg4(gy + (k * C4 + c4) * 4), a = dd::ldg4(act + t * 4);
    float4 o;
    o.x = a.x > 0.f ? g.x * inv : 0.f; o.y = a.y > 0.f ? g.y * inv : 0.f;
    o.z = a.z > 0.f ? g.z * inv : 0.f; o.w = a.w > 0.f ? g.w * inv : 0.f;
    *reinterpret_cast<float4*>(gx + t * 4) = o;
  }
}

__global__ void relu_bwd_kernel(const float* __restrict__ g, const float* __restrict__ act, float* __restrict__ out,
                                long long n4, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n4; t += stride) {
    const float4 gv = dd::ldg4(g + t * 4), av = dd::ldg4(act + t * 4);
    float4 o;
    o.x = av.x > 0.f ? gv.x : 0.f; o.y = av.y > 0.f ? gv.y : 0.f;
    o.z = av.z > 0.f ? gv.z : 0.f; o.w = av.w > 0.f ? gv.w : 0.f;
    *reinterpret_cast<float4*>(out + t * 4) = o;
  }
  for (long long t = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += stride)
    out[t] = act[t] > 0.f ? g[t] : 0.f;
}

__global__ void scale_kernel(const float* __restrict__ g, float w, float* __restrict__ out, long long n4, long long n,
                             int accumulate) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n4; t += stride) {
    const float4 gv = dd::ldg4(g + t * 4);
    float4 o = make_float4(w * gv.x, w * gv.y, w * gv.z, w * gv.w);
    if (accumulate) {
      const float4 p = *reinterpret_cast<const float4*>(out + t * 4);
      o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
    }
    *reinterpret_cast<float4*>(out + t * 4) = o;
  }
  for (long long t = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += stride)
    out[t] = accumulate ? out[t] + w * g[t] : w * g[t];
}

__global__ void scale_dev_kernel(const float* __restrict__ g, const float* __restrict__ w_dev, float* __restrict__ out,
                                 long long n4, long long n, int accumulate) {
  const float w = __ldg(w_dev);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n4; t += stride) {
    const float4 gv = dd::ldg4(g + t * 4);
    float4 o = make_float4(w * gv.x, w * gv.y, w * gv.z, w * gv.w);
    if (accumulate) {
      const float4 p = *reinterpret_cast<const float4*>(out + t * 4);
      o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
    }
    *reinterpret_cast<float4*>(out + t * 4) = o;
  }
  for (long long t = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += stride)
    out[t] = accumulate ? out[t] + w * g[t] : w * g[t];
}

// AdvGRL weight on device (da_heads.py:173-195, evident intent — SURVEY §9.2):
// w = loss <= bce ? -lam_adv * min(threshold, 1/loss) : -lam
__global__ void adv_grl_weight_kernel(const float* __restrict__ loss, float bce, float lam, float lam_adv,
                                      float threshold, float* __restrict__ w_out) {
  const float L = *loss;
  *w_out = (L <= bce) ? -1.0f * lam_adv * fminf(threshold, 1.0f / L) : -1.0f * lam;
}

__global__ void dropout_kernel(const float* __restrict__ x, const float* __restrict__ keep, float* __restrict__ out,
                               long long n) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
    out[t] = x[t] * keep[t] * 2.0f;
}

__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, long long n,
                           float lr, const float* __restrict__ lr_dev, float momentum, float wd, float grad_scale,
                           int first_step) {
  if (lr_dev) lr = lr * __ldg(lr_dev);     // learning rate kept on the device (a captured graph replays it)
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const float pv = p[t];
    float gv = g[t] * grad_scale;
    if (wd != 0.f) gv = gv + wd * pv;          // d_p = d_p.add(p, alpha=weight_decay)
    float b;
    if (first_step) b = gv;                    // buf = clone(d_p)
    else b = momentum * buf[t] + gv;           // buf.mul_(momentum).add_(d_p)
    buf[t] = b;
    p[t] = pv - lr * b;                        // p.add_(buf, alpha=-lr)
  }
}

}  // namespace

extern "C" int dd_maxpool3x3s2(const float* x, float* y, int N, int H, int W, int C, void* stream) {
  DD_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0);
  const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  const long long total = (long long)N * OH * OW * (C / 4);
  maxpool3x3s2_kernel<<<dd::grid_for(total, 256), 256, 0, dd::S(stream)>>>(x, y, N, H, W, C, OH, OW);
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_avgpool_forward(const float* x, float* y, int K, int HW, int C, void* stream) {
  DD_CHECK_ARG(K >= 0 && HW > 0 && C > 0);
  if (K == 0) return 0;
  avgpool_fwd_kernel<<<dd::grid_for((long long)K * C, 256), 256, 0, dd::S(stream)>>>(x, y, K, HW, C);
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_avgpool_backward(const float* gy, float* gx, int K, int HW, int C, void* stream) {
  DD_CHECK_ARG(K >= 0 && HW > 0 && C > 0);
  if (K == 0) return 0;
  avgpool_bwd_kernel<<<dd::grid_for((long long)K * HW * C, 256), 256, 0, dd::S(stream)>>>(gy, gx, K, HW, C);
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_avgpool_relu_backward(const float* gy, const float* act, float* gx, int K, int HW, int C,
                                        void* stream) {
  DD_CHECK_ARG(K >= 0 && HW > 0 && C > 0 && C % 4 == 0);
  if (K == 0) return 0;
  avgpool_relu_bwd_kernel<<<dd::grid_for((long long)K * HW * (C / 4), 256), 256, 0, dd::S(stream)>>>(gy, act, gx, K,
                                                                                                  HW, C / 4);
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_relu_backward(const float* g, const float* act, float* out, long long n, void* stream) {
  DD_CHECK_ARG(n >= 0);
  if (n == 0) return 0;
  relu_bwd_kernel<<<dd::grid_for(n / 4 + 1, 256), 256, 0, dd::S(stream)>>>(g, act, out, n / 4, n);
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_grl_backward(const float* g, float w, float* out, long long n, int accumulate, void* stream) {
  DD_CHECK_ARG(n >= 0);
  if (n == 0) return 0;
  scale_kernel<<<dd::grid_for(n / 4 + 1, 256), 256, 0, dd::S(stream)>>>(g, w, out, n / 4, n, accumulate);
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_grl_backward_dev(const float* g, const float* w_dev, float* out, long long n, int accumulate,
                                   void* stream) {
  DD_CHECK_ARG(n >= 0 && w_dev != nullptr);
  if (n == 0) return 0;
  scale_dev_kernel<<<dd::grid_for(n / 4 + 1, 256), 256, 0, dd::S(stream)>>>(g, w_dev, out, n / 4, n, accumulate);
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_adv_grl_weight(const float* loss, float bce, float lam, float lam_adv, float threshold, float* w_out,
                                 void* stream) {
  adv_grl_weight_kernel<<<1, 1, 0, dd::S(stream)>>>(loss, bce, lam, lam_adv, threshold, w_out);
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_dropout_apply(const float* x, const float* keep, float* out, long long n, void* stream) {
  DD_CHECK_ARG(n >= 0);
  if (n == 0) return 0;
  dropout_kernel<<<dd::grid_for(n, 256), 256, 0, dd::S(stream)>>>(x, keep, out, n);
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_sgd_momentum(float* p, const float* g, float* buf, long long n, float lr, float momentum, float wd,
                               float grad_scale, int first_step, void* stream) {
  DD_CHECK_ARG(n >= 0);
  if (n == 0) return 0;
  sgd_kernel<<<dd::grid_for(n, 256), 256, 0, dd::S(stream)>>>(p, g, buf, n, lr, nullptr, momentum, wd, grad_scale,
                                                              first_step);
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_sgd_momentum_dev(float* p, const float* g, float* buf, long long n, const float* lr_dev,
                                   float lr_factor, float momentum, float wd, float grad_scale, void* stream) {
  DD_CHECK_ARG(n >= 0 && lr_dev != nullptr);
  if (n == 0) return 0;
  sgd_kernel<<<dd::grid_for(n, 256), 256, 0, dd::S(stream)>>>(p, g, buf, n, lr_factor, lr_dev, momentum, wd,
                                                              grad_scale, 0);
  DD_LAUNCHED();
  return 0;
}
