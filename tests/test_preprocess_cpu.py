"""CPU: the input-pipeline oracle (oracle/preprocess_ref.py) is pinned against Pillow itself and against the
REAL reference transform chain + BatchCollator (tests/golden/preprocess_ref.pt, made by
`python oracle/make_golden.py preprocess`); the product's host logic (output-size rule, draw order, coefficient
tables from the C library, box transforms) is checked against both.  No compute call needs a GPU here —
dd_resample_coeffs / dd_resample_ksize are host functions of the C ABI."""
import os
import random

import numpy as np
import pytest
import torch

import preprocess_ref as pr


def cfg_for(opts):
    from dadetect_b200.config import get_cfg_defaults
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = get_cfg_defaults()
    cfg.merge_from_file(os.path.join(root, "configs", "da_faster_rcnn",
                                     "e2e_da_faster_rcnn_R_50_C4_cityscapes_to_foggy_cityscapes.yaml"))
    cfg.merge_from_list(list(opts))
    return cfg


@pytest.fixture(scope="module")
def cases(golden_dir):
    return torch.load(os.path.join(golden_dir, "preprocess_ref.pt"), weights_only=False)


@pytest.mark.parametrize("h,w,oh,ow", [(37, 53, 20, 31), (37, 53, 37, 31), (37, 53, 50, 53), (37, 53, 80, 120),
                                       (64, 128, 38, 75), (33, 77, 12, 200), (5, 7, 1, 1), (3, 3, 9, 9),
                                       (256, 512, 150, 300)])
def test_oracle_resampling_is_bit_exact_pillow(h, w, oh, ow):
    from PIL import Image
    img = np.random.default_rng(h * w + oh).integers(0, 256, (h, w, 3), dtype=np.uint8)
    want = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
    assert np.array_equal(pr.pil_resize_bilinear(img, oh, ow), want)


def _replay_plans(case, cfg):
    """The reference's draw order per sample: random.choice(min_size), random.random() (transforms.py:45,77)."""
    random.seed(case["seed"])
    plans = []
    for raw in case["raw"]:
        size = random.choice(tuple(cfg.INPUT.MIN_SIZE_TRAIN))
        out_hw = pr.get_size((raw.shape[1], raw.shape[0]), size, cfg.INPUT.MAX_SIZE_TRAIN)
        plans.append((out_hw, random.random() < 0.5))
    return plans


def test_oracle_matches_real_reference_pipeline(cases):
    assert len(cases) >= 5
    flips = 0
    for name, case in cases.items():
        cfg = cfg_for(case["opts"])
        plans = _replay_plans(case, cfg)
        flips += sum(f for _, f in plans)
        tensors = [pr.transform_image(raw.numpy(), hw, flip, cfg.INPUT.PIXEL_MEAN, cfg.INPUT.PIXEL_STD,
                                      cfg.INPUT.TO_BGR255) for raw, (hw, flip) in zip(case["raw"], plans)]
        batch, sizes = pr.collate(tensors, cfg.DATALOADER.SIZE_DIVISIBILITY)
        assert [tuple(s) for s in sizes] == case["image_sizes"], name
        assert np.array_equal(batch, case["batch"].numpy()), name          # bit-exact, floats included
        for raw, bx, (hw, flip), want, wsize in zip(case["raw"], case["boxes"], plans, case["target_boxes"],
                                                     case["target_sizes"]):
            b = pr.resize_boxes(bx.numpy(), (raw.shape[1], raw.shape[0]), (hw[1], hw[0]))
            if flip:
                b = pr.hflip_boxes(b, hw[1])
            assert tuple(wsize) == (hw[1], hw[0])
            assert np.array_equal(b, want.numpy()), name
    assert flips >= 2                                                       # the fixtures do exercise the flip


def test_host_plan_and_targets_match_reference(cases):
    from dadetect_b200.data import build_transforms
    from dadetect_b200.structures import BoxList
    for name, case in cases.items():
        cfg = cfg_for(case["opts"])
        tf = build_transforms(cfg, is_train=True, device="cpu")          # host decisions only; no kernel is run
        random.seed(case["seed"])
        for raw, bx, want_hw, want_boxes in zip(case["raw"], case["boxes"], case["image_sizes"], case["target_boxes"]):
            out_hw, flip = tf.plan((raw.shape[1], raw.shape[0]))
            assert tuple(out_hw) == tuple(want_hw), name
            t = BoxList(bx.clone(), (raw.shape[1], raw.shape[0]), mode="xyxy")
            t.add_field("labels", torch.arange(1, 6))
            t = tf.transform_target(t, out_hw, flip)
            assert t.size == (want_hw[1], want_hw[0])
            assert torch.equal(t.bbox, want_boxes), name
            assert torch.equal(t.get_field("labels"), torch.arange(1, 6))


@pytest.mark.parametrize("a,b", [(53, 31), (53, 53), (53, 120), (2048, 1200), (1024, 600), (7, 1), (3, 9), (1000, 3)])
def test_c_abi_coefficient_tables_equal_oracle(a, b):
    from dadetect_b200.data import resample_coeffs
    bounds, kk = resample_coeffs(a, b)
    rb, rk, ks = pr.precompute_coeffs(a, b)
    assert kk.shape[1] == ks
    assert np.array_equal(bounds.numpy(), rb) and np.array_equal(kk.numpy(), rk)


def test_eval_transform_never_flips_and_rejects_host_images():
    from dadetect_b200.data import build_transforms
    cfg = cfg_for(["INPUT.MIN_SIZE_TEST", 600, "INPUT.MAX_SIZE_TEST", 1200])
    tf = build_transforms(cfg, is_train=False, device="cpu")
    random.seed(0)
    for _ in range(20):
        out_hw, flip = tf.plan((2048, 1024))
        assert out_hw == (600, 1200) and flip is False
    with pytest.raises(RuntimeError):                                       # no CPU path for the arithmetic
        tf.run(torch.zeros((4, 4, 3), dtype=torch.uint8), (4, 4), False, torch.zeros((3, 4, 4)))


def test_triplet_collator_draws_sample_by_sample():
    """BatchCollator_triplet's datasets transform (image, image_p, image_n) of one sample before the next sample:
    the plans of the device collator must consume Python's `random` in exactly that order."""
    from dadetect_b200.data import DeviceBatchCollatorTriplet, build_transforms
    cfg = cfg_for(["INPUT.MIN_SIZE_TRAIN", (30, 44, 52), "INPUT.MAX_SIZE_TRAIN", 90])
    coll = DeviceBatchCollatorTriplet(build_transforms(cfg, True, device="cpu"), 32)
    shapes = [(61, 97), (97, 61), (50, 50)]
    batch = [tuple(x for j in range(3) for x in (torch.zeros(shapes[(i + j) % 3] + (3,), dtype=torch.uint8), None))
             + (i, i, i) for i in range(2)]
    random.seed(5)
    plans = coll.plan_batch(batch)
    random.seed(5)
    want = []
    for i in range(2):
        row = []
        for j in range(3):
            h, w = shapes[(i + j) % 3]
            size = random.choice((30, 44, 52))
            row.append((pr.get_size((w, h), size, 90), random.random() < 0.5))
        want.append(row)
    assert plans == want
