"""TEST INFRASTRUCTURE ONLY.  Recipe that compiles the reference's own CPU kernels
(csrc/cpu/ROIAlign_cpu.cpp, csrc/cpu/nms_cpu.cpp) from /root/reference into
oracle/_ref/refcpu_C*.so.  Nothing is copied into the repo; oracle/_ref/ is
git-ignored but travels to the GPU box.  Used (a) to pin oracle/ restatements and
(b) as maskrcnn_benchmark._C when oracle/make_golden.py imports the reference.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DADETECT_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def build(verbose=False):
    csrc = os.path.join(REF, "maskrcnn_benchmark", "csrc")
    if not os.path.isdir(csrc):
        return None
    from torch.utils.cpp_extension import load
    os.makedirs(OUT, exist_ok=True)
    return load(
        name="refcpu_C",
        sources=[os.path.join(HERE, "ref_shim.cpp")],
        extra_include_paths=[csrc],
        extra_cflags=["-O2", "-w"],
        build_directory=OUT,
        verbose=verbose,
    )


def load_prebuilt():
    """Import oracle/_ref/refcpu_C.so if it was built earlier (GPU box: no /root/reference)."""
    import importlib.util
    import torch  # noqa: F401  (the extension links against libtorch)
    for f in sorted(os.listdir(OUT)) if os.path.isdir(OUT) else []:
        if f.startswith("refcpu_C") and f.endswith(".so"):
            spec = importlib.util.spec_from_file_location("refcpu_C", os.path.join(OUT, f))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
    return None


if __name__ == "__main__":
    m = build(verbose=True)
    print("built" if m is not None else "reference not present; skipped", file=sys.stderr)
