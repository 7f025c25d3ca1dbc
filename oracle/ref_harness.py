"""TEST INFRASTRUCTURE ONLY — imports the REAL reference package from /root/reference
in the build container (it does not exist on the GPU box) so that oracle/ can be
pinned against it and golden fixtures can be generated (oracle/make_golden.py).

Nothing under /root/reference is modified or copied.  The incompatibilities listed in
SURVEY.md §8c are bridged from the outside:
  1. `yacs` is absent          -> a `yacs.config` module exposing this repo's CfgNode;
  2. `np.float` removed        -> alias to float before the reference is imported;
  3. `maskrcnn_benchmark._C`   -> a synthetic module: roi_align_forward and nms are the
     reference's own CPU kernels compiled by oracle/build_ref.py (oracle/_ref/);
     roi_align_backward (no CPU implementation in the reference, csrc/ROIAlign.h:44)
     is oracle/roialign_nms_ref.c; the unused ops raise;
  4. `torch.cuda.FloatTensor` cast in da_heads/loss.py:99,172 -> aliased to
     torch.FloatTensor while running on CPU.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("DADETECT_REFERENCE", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)


def available():
    return os.path.isdir(os.path.join(REF, "maskrcnn_benchmark"))


def _product_cfgnode():
    sys.path.insert(0, _ROOT)
    from dadetect_b200.config import CfgNode
    return CfgNode


def install(nms_strict=False):
    """Make `import maskrcnn_benchmark` work.  nms_strict=True swaps the reference's CPU NMS
    (IoU >= thresh) for the oracle's strict-> variant, i.e. the reference's CUDA semantics
    (SURVEY §9.12), for goldens that a GPU implementation must match."""
    if "maskrcnn_benchmark" in sys.modules:
        return sys.modules["maskrcnn_benchmark"]
    if not hasattr(np, "float"):
        np.float = float
    yacs = types.ModuleType("yacs")
    yacs_config = types.ModuleType("yacs.config")
    yacs_config.CfgNode = _product_cfgnode()
    yacs.config = yacs_config
    sys.modules.setdefault("yacs", yacs)
    sys.modules.setdefault("yacs.config", yacs_config)
    torch.cuda.FloatTensor = torch.FloatTensor

    sys.path.insert(0, _HERE)
    import build_ref
    import da_frcnn_ref as orc
    refc = build_ref.load_prebuilt() or build_ref.build()

    C = types.ModuleType("maskrcnn_benchmark._C")
    C.roi_align_forward = refc.roi_align_forward

    def roi_align_backward(grad, rois, scale, ph, pw, bs, ch, h, w, sampling_ratio):
        gin = torch.zeros((bs, ch, h, w))
        g = grad.contiguous().float()
        r = rois.contiguous().float()
        orc.lib().ref_roi_align_backward(orc._fp(g), orc._fp(r), r.shape[0], ch, h, w, float(scale),
                                         int(ph), int(pw), int(sampling_ratio), orc._fp(gin))
        return gin

    C.roi_align_backward = roi_align_backward
    if nms_strict:
        C.nms = lambda dets, scores, thr: orc.nms(dets, scores, thr, strict=True)
    else:
        C.nms = refc.nms

    def _unused(*a, **k):
        raise RuntimeError("op not on the DA Faster R-CNN path")

    for name in ("roi_pool_forward", "roi_pool_backward", "sigmoid_focalloss_forward",
                 "sigmoid_focalloss_backward"):
        setattr(C, name, _unused)
    sys.modules["maskrcnn_benchmark._C"] = C
    sys.path.insert(0, REF)
    import maskrcnn_benchmark
    maskrcnn_benchmark._C = C
    return maskrcnn_benchmark


def reference_cfg(yaml_name=None, opts=()):
    """The reference's global default cfg (cloned), optionally merged with one of its YAMLs."""
    install()
    from maskrcnn_benchmark.config import cfg
    c = cfg.clone()
    if yaml_name:
        c.merge_from_file(os.path.join(REF, "configs", yaml_name))
    c.merge_from_list(list(opts))
    c.merge_from_list(["MODEL.DEVICE", "cpu"])
    return c


def build_reference_model(cfg, state_dict):
    install()
    import contextlib
    import io
    from maskrcnn_benchmark.modeling.detector import build_detection_model
    with contextlib.redirect_stdout(io.StringIO()):     # the reference prints anchor_stride
        model = build_detection_model(cfg)
    missing = model.load_state_dict(state_dict, strict=False)
    bad = [k for k in missing.missing_keys if "cell_anchors" not in k]
    assert not bad and not missing.unexpected_keys, (bad, missing.unexpected_keys)
    return model


def to_reference_targets(targets, image_hw):
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    out = []
    h, w = image_hw
    for t in targets:
        b = BoxList(t["boxes"].clone(), (w, h), mode="xyxy")
        b.add_field("labels", t["labels"].clone())
        b.add_field("is_source", torch.full((len(t["labels"]),), bool(t["is_source"]), dtype=torch.bool))
        out.append(b)
    return out
