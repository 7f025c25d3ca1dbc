// C-ABI entry points of the dense tier; selects the fp32 SIMT arm (conv_simt.cu) or the tcgen05 arm
// (conv_tc.cu).  There is no CPU or library fallback: an unsupported (impl, shape) pair is an error.
#include "common.cuh"

int dd_simt_conv2d_forward(const float*, const float*, const float*, const float*, const float*, float*, int, int,
                           int, int, int, int, int, int, int, int, cudaStream_t);
int dd_simt_conv2d_dgrad(const float*, const float*, const float*, const float*, const float*, float*, int, int, int,
                         int, int, int, int, int, int, cudaStream_t);
size_t dd_simt_wgrad_workspace_bytes(int, int, int, int, int, int, int, int, int);
int dd_simt_conv2d_wgrad(const float*, const float*, const float*, float*, int, int, int, int, int, int, int, int, int,
                         int, void*, cudaStream_t);

int dd_tc_conv2d_forward(const float*, const float*, const float*, const float*, const float*, float*, int, int, int,
                         int, int, int, int, int, int, int, bool, float*, bool, cudaStream_t);
int dd_tc_forward_prepare_batch(int, const float* const*, float* const*, const int*, const int*, const int*, const int*,
                                cudaStream_t);
int dd_tc_conv2d_dgrad(const float*, const float*, const float*, const float*, const float*, float*, int, int, int,
                       int, int, int, int, int, int, float*, int, bool, cudaStream_t);
int tc_rows_pad_public(int ncols);
int dd_tc_dgrad_prepare_batch(int, const float* const*, const float* const*, float* const*, const int*, const int*,
                              const int*, const int*, bool, cudaStream_t);
int dd_tc_conv2d_wgrad(const float*, const float*, const float*, float*, int, int, int, int, int, int, int, int, int,
                       int, void*, bool, cudaStream_t);
bool dd_tc_supports(int mode, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad);

static bool is_tc(int impl) { return impl == DD_IMPL_TCGEN05 || impl == DD_IMPL_TCGEN05_X3; }

extern "C" size_t dd_conv2d_forward_workspace_bytes(int Cin, int Cout, int KH, int KW, int impl) {
  if (impl != DD_IMPL_TCGEN05_X3) return 0;
  return sizeof(float) * 2 * (size_t)tc_rows_pad_public(Cout) * KH * KW * Cin;
}

static int conv2d_forward_impl(const float* x, const float* w, const float* scale, const float* bias,
                               const float* residual, float* y, int N, int H, int W, int Cin, int Cout, int KH, int KW,
                               int stride, int pad, int act, int impl, void* workspace, bool prepared, void* stream) {
  DD_CHECK_ARG(N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0);
  if (is_tc(impl) && dd_tc_supports(0, N, H, W, Cin, Cout, KH, KW, stride, pad)) {
    const bool x3 = impl == DD_IMPL_TCGEN05_X3;
    DD_CHECK_ARG(!x3 || (workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0));
    return dd_tc_conv2d_forward(x, w, scale, bias, residual, y, N, H, W, Cin, Cout, KH, KW, stride, pad, act, x3,
                                (float*)workspace, prepared, dd::S(stream));
  }
  DD_CHECK_ARG(!prepared);
  return dd_simt_conv2d_forward(x, w, scale, bias, residual, y, N, H, W, Cin, Cout, KH, KW, stride, pad, act,
                                dd::S(stream));
}

extern "C" int dd_conv2d_forward(const float* x, const float* w, const float* scale, const float* bias,
                                 const float* residual, float* y, int N, int H, int W, int Cin, int Cout, int KH,
                                 int KW, int stride, int pad, int act, int impl, void* workspace, void* stream) {
  return conv2d_forward_impl(x, w, scale, bias, residual, y, N, H, W, Cin, Cout, KH, KW, stride, pad, act, impl,
                             workspace, false, stream);
}

extern "C" int dd_conv2d_forward_prepared(const float* x, const float* w, const float* scale, const float* bias,
                                          const float* residual, float* y, int N, int H, int W, int Cin, int Cout,
                                          int KH, int KW, int stride, int pad, int act, int impl, void* workspace,
                                          void* stream) {
  DD_CHECK_ARG(impl == DD_IMPL_TCGEN05_X3);
  return conv2d_forward_impl(x, w, scale, bias, residual, y, N, H, W, Cin, Cout, KH, KW, stride, pad, act, impl,
                             workspace, true, stream);
}

extern "C" int dd_conv2d_forward_prepare_batch(int n, const float* const* w, void* const* workspaces, const int* Cin,
                                               const int* Cout, const int* KH, const int* KW, int impl, void* stream) {
  DD_CHECK_ARG(n >= 0 && impl == DD_IMPL_TCGEN05_X3);
  for (int i = 0; i < n; ++i)
    DD_CHECK_ARG(w[i] != nullptr && workspaces[i] != nullptr && (reinterpret_cast<uintptr_t>(workspaces[i]) & 15) == 0 &&
                 Cin[i] > 0 && Cout[i] > 0 && KH[i] > 0 && KW[i] > 0);
  return dd_tc_forward_prepare_batch(n, w, reinterpret_cast<float* const*>(workspaces), Cin, Cout, KH, KW, dd::S(stream));
}

extern "C" size_t dd_conv2d_dgrad_workspace_bytes(int Cin, int Cout, int KH, int KW) {
  // W' [Cin rows, padded to whole tiles][taps*Cout], twice (hi / lo planes of the 3xTF32 mode)
  return sizeof(float) * 2 * (size_t)tc_rows_pad_public(Cin) * Cout * KH * KW;
}

extern "C" int dd_conv2d_dgrad_prepare_batch(int n, const float* const* w, const float* const* scale,
                                             void* const* workspaces, const int* Cin, const int* Cout, const int* KH,
                                             const int* KW, int impl, void* stream) {
  DD_CHECK_ARG(n >= 0);
  if (n == 0 || !is_tc(impl)) return 0;               // the SIMT arm consumes the weights as they are
  for (int i = 0; i < n; ++i)
    DD_CHECK_ARG(w[i] != nullptr && workspaces[i] != nullptr && (reinterpret_cast<uintptr_t>(workspaces[i]) & 15) == 0 &&
                 Cin[i] > 0 && Cout[i] > 0 && KH[i] > 0 && KW[i] > 0);
  return dd_tc_dgrad_prepare_batch(n, w, scale, reinterpret_cast<float* const*>(workspaces), Cin, Cout, KH, KW,
                                   impl == DD_IMPL_TCGEN05_X3, dd::S(stream));
}

extern "C" int dd_conv2d_dgrad(const float* gy, const float* w, const float* scale, const float* addend,
                               const float* mask_act, float* gx, int N, int H, int W, int Cin, int Cout, int KH,
                               int KW, int stride, int pad, int impl, void* workspace, int prepared, void* stream) {
  DD_CHECK_ARG(N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0);
  if (is_tc(impl) && dd_tc_supports(1, N, H, W, Cin, Cout, KH, KW, stride, pad)) {
    DD_CHECK_ARG(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0);
    return dd_tc_conv2d_dgrad(gy, w, scale, addend, mask_act, gx, N, H, W, Cin, Cout, KH, KW, stride, pad,
                              (float*)workspace, prepared, impl == DD_IMPL_TCGEN05_X3, dd::S(stream));
  }
  return dd_simt_conv2d_dgrad(gy, w, scale, addend, mask_act, gx, N, H, W, Cin, Cout, KH, KW, stride, pad,
                              dd::S(stream));
}

extern "C" size_t dd_conv2d_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride,
                                                  int pad) {
  return dd_simt_wgrad_workspace_bytes(N, H, W, Cin, Cout, KH, KW, stride, pad);
}

extern "C" int dd_conv2d_wgrad(const float* gy, const float* x, const float* scale, float* gw, int N, int H, int W,
                               int Cin, int Cout, int KH, int KW, int stride, int pad, int accumulate, int impl,
                               void* workspace, void* stream) {
  DD_CHECK_ARG(N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0);
  if (is_tc(impl) && dd_tc_supports(2, N, H, W, Cin, Cout, KH, KW, stride, pad))
    return dd_tc_conv2d_wgrad(gy, x, scale, gw, N, H, W, Cin, Cout, KH, KW, stride, pad, accumulate, workspace,
                              impl == DD_IMPL_TCGEN05_X3, dd::S(stream));
  return dd_simt_conv2d_wgrad(gy, x, scale, gw, N, H, W, Cin, Cout, KH, KW, stride, pad, accumulate, workspace,
                              dd::S(stream));
}
