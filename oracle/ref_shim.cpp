// TEST INFRASTRUCTURE ONLY (oracle/): builds the reference's own CPU kernels from
// where they lie under /root/reference (never copied into this repo) into
// oracle/_ref/.  See oracle/build_ref.py for the recipe.
//
// The reference sources (maskrcnn_benchmark/csrc/cpu/ROIAlign_cpu.cpp:239-257,
// nms_cpu.cpp:67-75) dispatch on `tensor.type()`, which modern ATen no longer
// converts to a ScalarType.  Instead of patching them we give the dispatch macro
// the overload it looks up (::detail::scalar_type) for the deprecated type object.
#include <torch/extension.h>
namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
}  // namespace detail
#include "cpu/ROIAlign_cpu.cpp"
#include "cpu/nms_cpu.cpp"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("roi_align_forward", &ROIAlign_forward_cpu, "reference csrc/cpu/ROIAlign_cpu.cpp");
  m.def("nms", &nms_cpu, "reference csrc/cpu/nms_cpu.cpp");
}
