// fp32 implicit-GEMM convolution (forward / data-gradient / weight-gradient), NHWC x OHWI, on the
// CUDA cores.  This is the bit-faithful fp32 arm of the dense tier (impl = DD_IMPL_SIMT): fp32
// multiply-accumulate in the same precision as the reference's cuDNN/cuBLAS fp32 path (TF32 off),
// used for the parity tests and as the numerical yardstick of the tcgen05 arm (conv_tc.cu).
//
// One kernel template covers the three GEMMs of a convolution; FrozenBatchNorm scale/bias, residual
// add, ReLU (forward), ReLU-mask and gradient fan-in add (dgrad), BN-scale and split-K reduction
// (wgrad) are fused so no activation-sized tensor is re-read by an elementwise pass
// (reference: resnet.py:294-314 runs conv, x*scale+bias, relu, += identity as separate kernels).
//
//   forward : Y[m, co]  = sum_{tap,ci} X[pix(m)+tap, ci] * W[co, tap, ci]        M = N*OH*OW, K = T*Cin
//   dgrad   : GX[m, ci] = sum_{tap,co} GY[opix(m,tap), co] * s[co] * W[co,tap,ci] M = N*H*W (or N*OH*OW for 1x1/s>1)
//   wgrad   : GW[co, (tap,ci)] = s[co] * sum_{pix} GY[pix, co] * X[pix+tap, ci]   split over pixels
//
// Tiling: 128x128x8 CTA tile, 256 threads, 8x8 register micro-tile per thread, operands staged through
// shared memory with 128-bit global loads along the contiguous (channel) dimension.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 8, NT = 256;
enum Mode { FWD = 0, DGRAD = 1, WGRAD = 2 };

struct ConvP {
  const float* a;        // FWD: x, DGRAD: gy, WGRAD: gy
  const float* b;        // FWD: w, DGRAD: w,  WGRAD: x
  float* out;            // FWD: y, DGRAD: gx, WGRAD: partials
  const float* scale;    // per-Cout scale (may be null)
  const float* bias;     // FWD only
  const float* extra;    // FWD: residual, DGRAD: addend
  const float* mask;     // DGRAD: activation whose >0 gates the result
  int N, H, W, Cin, Cout, KH, KW, stride, pad, OH, OW;
  int M, Ncols, K;       // GEMM sizes
  int act;
  int compact;           // DGRAD 1x1 stride>1: rows enumerate output pixels, stores are strided
  int k_per_split;       // WGRAD
};

__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

template <int MODE>
__global__ void __launch_bounds__(NT) conv_gemm_kernel(const ConvP p) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  int k_begin = 0, k_end = p.K;
  if (MODE == WGRAD) {
    k_begin = blockIdx.z * p.k_per_split;
    k_end = min(p.K, k_begin + p.k_per_split);
  }
  const int T = p.KH * p.KW;

  // ---------------- per-thread load coordinates (fixed across the K loop) ----------------
  // K-contiguous operands (FWD A and B, DGRAD A): thread -> (row = tid/2, 4 consecutive k at (tid%2)*4)
  // N/M-contiguous operands (DGRAD B, WGRAD A and B): thread -> (k row = tid/32, 4 consecutive cols at (tid%32)*4)
  const int lrow = tid >> 1, lkq = (tid & 1) * 4;
  const int lk = tid >> 5, lcol = (tid & 31) * 4;

  // A row -> pixel decomposition (FWD: output pixel; DGRAD: input pixel or compact output pixel)
  int a_n = 0, a_h = 0, a_w = 0;
  bool a_row_ok = false;
  if (MODE == FWD || MODE == DGRAD) {
    const int m = m0 + lrow;
    a_row_ok = m < p.M;
    if (a_row_ok) {
      const int ww = (MODE == FWD || p.compact) ? p.OW : p.W;
      const int hh = (MODE == FWD || p.compact) ? p.OH : p.H;
      a_w = m % ww;
      a_h = (m / ww) % hh;
      a_n = m / (ww * hh);
    }
  }
  // WGRAD B column -> (tap, ci)
  int b_tap = 0, b_ci = 0;
  bool b_col_ok = false;
  if (MODE == WGRAD) {
    const int n = n0 + lcol;
    b_col_ok = n < p.Ncols;
    if (b_col_ok) { b_tap = n / p.Cin; b_ci = n % p.Cin; }
  }
  const bool vecA = (MODE == FWD) ? (p.Cin % 4 == 0) : (p.Cout % 4 == 0);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int ty = tid >> 4, tx = tid & 15;

  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    // ------------------------------ stage A ------------------------------
    if (MODE == FWD) {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      const int k = k0 + lkq;
      if (a_row_ok && k < p.K) {
        if (vecA) {
          const int tap = k / p.Cin, ci = k % p.Cin;
          const int ih = a_h * p.stride + tap / p.KW - p.pad, iw = a_w * p.stride + tap % p.KW - p.pad;
          if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) {
            const float4 t = dd::ldg4(p.a + (((size_t)a_n * p.H + ih) * p.W + iw) * p.Cin + ci);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int kk = k + e;
            if (kk < p.K) {
              const int tap = kk / p.Cin, ci = kk % p.Cin;
              const int ih = a_h * p.stride + tap / p.KW - p.pad, iw = a_w * p.stride + tap % p.KW - p.pad;
              if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W)
                v[e] = __ldg(p.a + (((size_t)a_n * p.H + ih) * p.W + iw) * p.Cin + ci);
            }
          }
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) As[lkq + e][lrow] = v[e];
    } else if (MODE == DGRAD) {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      const int k = k0 + lkq;
      if (a_row_ok && k < p.K) {
        // k = tap*Cout + co; with Cout % 4 == 0 the 4 consecutive k share one tap (vector load)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (vecA && e > 0) break;
          const int kk = k + e;
          if (kk >= p.K) break;
          const int tap = kk / p.Cout, co = kk % p.Cout;
          bool ok = true;
          int oh = a_h, ow = a_w;
          if (!p.compact) {
            const int hn = a_h + p.pad - tap / p.KW, wn = a_w + p.pad - tap % p.KW;
            ok = hn >= 0 && wn >= 0 && (hn % p.stride == 0) && (wn % p.stride == 0);
            oh = hn / p.stride; ow = wn / p.stride;
            ok = ok && oh < p.OH && ow < p.OW;
          }
          if (!ok) continue;
          const float* src = p.a + (((size_t)a_n * p.OH + oh) * p.OW + ow) * p.Cout + co;
          if (vecA) { const float4 t = dd::ldg4(src); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
          else v[e] = __ldg(src);
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) As[lkq + e][lrow] = v[e];
    } else {  // WGRAD: A(m = co, k = pixel) = gy[pixel*Cout + co], contiguous along m
      float4 t = zero4();
      const int k = k0 + lk, m = m0 + lcol;
      if (k < k_end && m < p.M) {
        const float* src = p.a + (size_t)k * p.Cout + m;
        if (p.Cout % 4 == 0) t = dd::ldg4(src);
        else {
          t.x = __ldg(src);
          if (m + 1 < p.M) t.y = __ldg(src + 1);
          if (m + 2 < p.M) t.z = __ldg(src + 2);
          if (m + 3 < p.M) t.w = __ldg(src + 3);
        }
      }
      *reinterpret_cast<float4*>(&As[lk][lcol]) = t;
    }
    // ------------------------------ stage B ------------------------------
    if (MODE == FWD) {  // B(k, n) = w[n*K + k], contiguous along k
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      const int k = k0 + lkq, n = n0 + lrow;
      if (n < p.Ncols && k < p.K) {
        const float* src = p.b + (size_t)n * p.K + k;
        if (p.K % 4 == 0) { const float4 t = dd::ldg4(src); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
        else {
#pragma unroll
          for (int e = 0; e < 4; ++e) if (k + e < p.K) v[e] = __ldg(src + e);
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) Bs[lkq + e][lrow] = v[e];
    } else if (MODE == DGRAD) {  // B(k = (tap,co), n = ci) = s[co] * w[co, tap, ci], contiguous along n
      float4 t = zero4();
      const int k = k0 + lk, n = n0 + lcol;
      if (k < p.K && n < p.Ncols) {
        const int tap = k / p.Cout, co = k % p.Cout;
        const float* src = p.b + ((size_t)co * T + tap) * p.Cin + n;
        if (p.Cin % 4 == 0) t = dd::ldg4(src);
        else {
          t.x = __ldg(src);
          if (n + 1 < p.Ncols) t.y = __ldg(src + 1);
          if (n + 2 < p.Ncols) t.z = __ldg(src + 2);
          if (n + 3 < p.Ncols) t.w = __ldg(src + 3);
        }
        if (p.scale) { const float s = __ldg(p.scale + co); t.x *= s; t.y *= s; t.z *= s; t.w *= s; }
      }
      *reinterpret_cast<float4*>(&Bs[lk][lcol]) = t;
    } else {  // WGRAD: B(k = pixel, n = (tap,ci)) = x[pix+tap, ci], contiguous along n (Cin % 4 == 0 required)
      float4 t = zero4();
      const int k = k0 + lk;
      if (k < k_end && b_col_ok) {
        const int ow = k % p.OW, oh = (k / p.OW) % p.OH, nn = k / (p.OW * p.OH);
        const int ih = oh * p.stride + b_tap / p.KW - p.pad, iw = ow * p.stride + b_tap % p.KW - p.pad;
        if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W)
          t = dd::ldg4(p.b + (((size_t)nn * p.H + ih) * p.W + iw) * p.Cin + b_ci);
      }
      *reinterpret_cast<float4*>(&Bs[lk][lcol]) = t;
    }
    __syncthreads();
    // ------------------------------ 8x8 outer products ------------------------------
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ------------------------------ epilogue ------------------------------
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
    size_t row_off;
    if (MODE == FWD) row_off = (size_t)m * p.Cout;
    else if (MODE == DGRAD) {
      if (p.compact) {
        const int ow = m % p.OW, oh = (m / p.OW) % p.OH, nn = m / (p.OW * p.OH);
        row_off = (((size_t)nn * p.H + (size_t)oh * p.stride) * p.W + (size_t)ow * p.stride) * p.Cin;
      } else row_off = (size_t)m * p.Cin;
    } else row_off = ((size_t)blockIdx.z * p.M + m) * p.Ncols;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = n0 + jh * 64 + tx * 4;
      if (n >= p.Ncols) continue;
      float r[4] = {acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (n + e >= p.Ncols) continue;
        float v = r[e];
        if (MODE == FWD) {
          if (p.scale) v *= __ldg(p.scale + n + e);
          if (p.bias) v += __ldg(p.bias + n + e);
          if (p.extra) v += __ldg(p.extra + row_off + n + e);
          if (p.act == DD_ACT_RELU) v = fmaxf(v, 0.f);
        } else if (MODE == DGRAD) {
          if (p.extra) v += __ldg(p.extra + row_off + n + e);
          if (p.mask) v = __ldg(p.mask + row_off + n + e) > 0.f ? v : 0.f;
        }
        r[e] = v;
      }
      float* dst = p.out + row_off + n;
      if (n + 3 < p.Ncols && ((row_off + n) & 3) == 0) *reinterpret_cast<float4*>(dst) = make_float4(r[0], r[1], r[2], r[3]);
      else {
#pragma unroll
        for (int e = 0; e < 4; ++e) if (n + e < p.Ncols) dst[e] = r[e];
      }
    }
  }
}

// gw = (accumulate ? gw : 0) + scale[co] * sum_splits partial[split][co][n]
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int M, int Ncols,
                                    const float* __restrict__ scale, float* __restrict__ gw, int accumulate) {
  const long long total = (long long)M * Ncols;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += partial[(size_t)s * total + t];
    if (scale) acc *= __ldg(scale + t / Ncols);
    gw[t] = accumulate ? gw[t] + acc : acc;
  }
}

// dgrad with stride > 1 in compact mode: positions the strided conv never read get addend*mask or 0
__global__ void dgrad_fill_kernel(const float* __restrict__ addend, const float* __restrict__ mask,
                                  float* __restrict__ gx, long long n) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    float v = addend ? addend[t] : 0.f;
    if (mask) v = mask[t] > 0.f ? v : 0.f;
    gx[t] = v;
  }
}

// Column sums of gy [rows, C].  grid (channel groups of 32 x `vec`, row slices): coalesced (vector) reads along
// channels, in-CTA tree over the 8 row lanes, one atomicAdd per channel per CTA into the zero-initialised /
// accumulated gb.
template <int VEC>
__global__ void bias_grad_kernel(const float* __restrict__ gy, float* __restrict__ gb, int rows, int C,
                                 int rows_per_cta) {
  __shared__ float part[8][32 * VEC + 1];
  const int c = (blockIdx.x * 32 + threadIdx.x) * VEC;
  const int r_begin = blockIdx.y * rows_per_cta, r_end = min(rows, r_begin + rows_per_cta);
  float acc[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
  if (c < C) {
    int r = r_begin + threadIdx.y;
    if (VEC == 4) {
      // four independent 128-bit loads in flight per thread (the kernel is a pure stream: 67 MB for the RPN conv)
      for (; r + 24 < r_end; r += 32) {
        const float4 v0 = dd::ldg4(gy + (size_t)r * C + c), v1 = dd::ldg4(gy + (size_t)(r + 8) * C + c);
        const float4 v2 = dd::ldg4(gy + (size_t)(r + 16) * C + c), v3 = dd::ldg4(gy + (size_t)(r + 24) * C + c);
        acc[0] += (v0.x + v1.x) + (v2.x + v3.x); acc[1] += (v0.y + v1.y) + (v2.y + v3.y);
        acc[2] += (v0.z + v1.z) + (v2.z + v3.z); acc[3] += (v0.w + v1.w) + (v2.w + v3.w);
      }
    }
    for (; r < r_end; r += 8) {
      if (VEC == 4) {
        const float4 v = dd::ldg4(gy + (size_t)r * C + c);
        acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
      } else {
        acc[0] += gy[(size_t)r * C + c];
      }
    }
  }
#pragma unroll
  for (int e = 0; e < VEC; ++e) part[threadIdx.y][threadIdx.x * VEC + e] = acc[e];
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += part[i][threadIdx.x * VEC + e];
      atomicAdd(gb + c + e, s);
    }
  }
}

int out_dim(int in, int k, int s, int pad) { return (in + 2 * pad - k) / s + 1; }

int wgrad_splits(int M, int Ncols, int K) {
  const int tiles = ((M + BM - 1) / BM) * ((Ncols + BN - 1) / BN);
  int splits = (4 * dd::kNumSMs + tiles - 1) / tiles;
  const int max_by_k = (K + 255) / 256;
  if (splits > max_by_k) splits = max_by_k;
  if (splits > 128) splits = 128;
  if (splits < 1) splits = 1;
  return splits;
}

}  // namespace

int dd_simt_wgrad_splits(int M, int Ncols, int K) { return wgrad_splits(M, Ncols, K); }

int dd_wgrad_reduce(const float* partial, int splits, int M, int Ncols, const float* scale, float* gw, int accumulate,
                    cudaStream_t s) {
  const long long total = (long long)M * Ncols;
  wgrad_reduce_kernel<<<dd::grid_for(total, 256), 256, 0, s>>>(partial, splits, M, Ncols, scale, gw, accumulate);
  DD_LAUNCHED();
  return 0;
}

// ---- entry points for the SIMT arm (dispatch from conv_dispatch.cu) -------------------------------------
int dd_simt_conv2d_forward(const float* x, const float* w, const float* scale, const float* bias,
                           const float* residual, float* y, int N, int H, int W, int Cin, int Cout, int KH, int KW,
                           int stride, int pad, int act, cudaStream_t s) {
  ConvP p = {};
  p.a = x; p.b = w; p.out = y; p.scale = scale; p.bias = bias; p.extra = residual;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.KH = KH; p.KW = KW; p.stride = stride; p.pad = pad;
  p.OH = out_dim(H, KH, stride, pad); p.OW = out_dim(W, KW, stride, pad);
  p.M = N * p.OH * p.OW; p.Ncols = Cout; p.K = KH * KW * Cin; p.act = act;
  DD_CHECK_ARG(p.OH > 0 && p.OW > 0 && (long long)N * p.OH * p.OW < (1ll << 31));
  dim3 grid((p.Ncols + BN - 1) / BN, (p.M + BM - 1) / BM);
  DD_CHECK_ARG(grid.y <= 65535);
  conv_gemm_kernel<FWD><<<grid, NT, 0, s>>>(p);
  DD_LAUNCHED();
  return 0;
}

int dd_simt_conv2d_dgrad(const float* gy, const float* w, const float* scale, const float* addend,
                         const float* mask_act, float* gx, int N, int H, int W, int Cin, int Cout, int KH, int KW,
                         int stride, int pad, cudaStream_t s) {
  ConvP p = {};
  p.a = gy; p.b = w; p.out = gx; p.scale = scale; p.extra = addend; p.mask = mask_act;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.KH = KH; p.KW = KW; p.stride = stride; p.pad = pad;
  p.OH = out_dim(H, KH, stride, pad); p.OW = out_dim(W, KW, stride, pad);
  p.compact = (KH == 1 && KW == 1 && pad == 0 && stride > 1) ? 1 : 0;
  p.M = p.compact ? N * p.OH * p.OW : N * H * W;
  p.Ncols = Cin; p.K = KH * KW * Cout;
  DD_CHECK_ARG(p.OH > 0 && p.OW > 0 && (long long)N * H * W < (1ll << 31));
  if (p.compact) {
    const long long n = (long long)N * H * W * Cin;
    dgrad_fill_kernel<<<dd::grid_for(n, 256), 256, 0, s>>>(addend, mask_act, gx, n);
    DD_LAUNCHED();
  }
  dim3 grid((p.Ncols + BN - 1) / BN, (p.M + BM - 1) / BM);
  DD_CHECK_ARG(grid.y <= 65535);
  conv_gemm_kernel<DGRAD><<<grid, NT, 0, s>>>(p);
  DD_LAUNCHED();
  return 0;
}

size_t dd_simt_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad) {
  const int OH = out_dim(H, KH, stride, pad), OW = out_dim(W, KW, stride, pad);
  const int M = Cout, Ncols = KH * KW * Cin, K = N * OH * OW;
  return (size_t)wgrad_splits(M, Ncols, K) * M * Ncols * sizeof(float) + 256;
}

int dd_simt_conv2d_wgrad(const float* gy, const float* x, const float* scale, float* gw, int N, int H, int W, int Cin,
                         int Cout, int KH, int KW, int stride, int pad, int accumulate, void* workspace,
                         cudaStream_t s) {
  DD_CHECK_ARG(Cin % 4 == 0 && workspace != nullptr);
  ConvP p = {};
  p.a = gy; p.b = x; p.out = (float*)workspace; p.scale = nullptr;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.KH = KH; p.KW = KW; p.stride = stride; p.pad = pad;
  p.OH = out_dim(H, KH, stride, pad); p.OW = out_dim(W, KW, stride, pad);
  p.M = Cout; p.Ncols = KH * KW * Cin; p.K = N * p.OH * p.OW;
  DD_CHECK_ARG(p.OH > 0 && p.OW > 0 && (long long)N * p.OH * p.OW < (1ll << 31));
  const int splits = wgrad_splits(p.M, p.Ncols, p.K);
  p.k_per_split = ((p.K + splits - 1) / splits + BK - 1) / BK * BK;
  dim3 grid((p.Ncols + BN - 1) / BN, (p.M + BM - 1) / BM, splits);
  conv_gemm_kernel<WGRAD><<<grid, NT, 0, s>>>(p);
  DD_LAUNCHED();
  const long long total = (long long)p.M * p.Ncols;
  wgrad_reduce_kernel<<<dd::grid_for(total, 256), 256, 0, s>>>(p.out, splits, p.M, p.Ncols, scale, gw, accumulate);
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_bias_grad(const float* gy, float* gb, int rows, int C, int accumulate, void* stream) {
  DD_CHECK_ARG(rows > 0 && C > 0);
  cudaStream_t s = dd::S(stream);
  if (!accumulate) DD_CUDA(cudaMemsetAsync(gb, 0, sizeof(float) * (size_t)C, s));
  dim3 block(32, 8);
  const bool vec = C % 4 == 0 && (reinterpret_cast<uintptr_t>(gy) & 15) == 0;
  const int groups = vec ? (C + 127) / 128 : (C + 31) / 32;
  int slices = (16 * dd::kNumSMs + groups - 1) / groups;
  if (slices > (rows + 63) / 64) slices = (rows + 63) / 64;
  if (slices < 1) slices = 1;
  const int rows_per_cta = (rows + slices - 1) / slices;
  dim3 grid(groups, (rows + rows_per_cta - 1) / rows_per_cta);
  if (vec) bias_grad_kernel<4><<<grid, block, 0, s>>>(gy, gb, rows, C, rows_per_cta);
  else bias_grad_kernel<1><<<grid, block, 0, s>>>(gy, gb, rows, C, rows_per_cta);
  DD_LAUNCHED();
  return 0;
}
