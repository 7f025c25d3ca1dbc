// tcgen05 / TMA arm of the dense tier — placeholder until the kernels land: reports "unsupported" so the
// dispatcher uses the fp32 SIMT arm.
#include "common.cuh"

extern "C" int dd_tcgen05_built(void) { return 0; }
bool dd_tc_supports(int, int, int, int, int, int, int, int, int, int) { return false; }
int dd_tc_conv2d_forward(const float*, const float*, const float*, const float*, const float*, float*, int, int, int,
                         int, int, int, int, int, int, int, cudaStream_t) { return dd::fail(-1, "tcgen05 arm not built", __FILE__, __LINE__); }
int dd_tc_conv2d_dgrad(const float*, const float*, const float*, const float*, const float*, float*, int, int, int,
                       int, int, int, int, int, int, cudaStream_t) { return dd::fail(-1, "tcgen05 arm not built", __FILE__, __LINE__); }
int dd_tc_conv2d_wgrad(const float*, const float*, const float*, float*, int, int, int, int, int, int, int, int, int,
                       int, void*, cudaStream_t) { return dd::fail(-1, "tcgen05 arm not built", __FILE__, __LINE__); }
