"""GPU parity of the input pipeline kernel (dd_preprocess_image through dadetect_b200.data): bit-exact — bytes of
the resampled image AND the normalised floats — against the real reference pipeline's golden batches and against
the CPU oracle (oracle/preprocess_ref.py, pinned to Pillow and to the reference by tests/test_preprocess_cpu.py)."""
import os
import random

import numpy as np
import pytest
import torch

import preprocess_ref as pr
from test_preprocess_cpu import cfg_for

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def cases(golden_dir):
    return torch.load(os.path.join(golden_dir, "preprocess_ref.pt"), weights_only=False)


def test_collator_matches_real_reference_batches(cases):
    from dadetect_b200.data import DeviceBatchCollator, build_transforms
    from dadetect_b200.structures import BoxList
    for name, case in cases.items():
        cfg = cfg_for(case["opts"])
        collate = DeviceBatchCollator(build_transforms(cfg, is_train=True, device=DEV), cfg.DATALOADER.SIZE_DIVISIBILITY)
        samples = []
        for i, (raw, bx) in enumerate(zip(case["raw"], case["boxes"])):
            t = BoxList(bx.clone(), (raw.shape[1], raw.shape[0]), mode="xyxy")
            t.add_field("labels", torch.arange(1, 6))
            samples.append((raw.numpy() if i % 2 else raw, t, i))         # numpy and torch host images
        random.seed(case["seed"])
        images, targets, ids = collate(samples)
        torch.cuda.synchronize()
        assert [tuple(s) for s in images.image_sizes] == case["image_sizes"], name
        assert images.tensors.is_cuda and tuple(images.tensors.shape) == tuple(case["batch"].shape), name
        assert torch.equal(images.tensors.cpu(), case["batch"]), name     # bit-exact floats, padding included
        for t, want in zip(targets, case["target_boxes"]):
            assert torch.equal(t.bbox, want), name
        assert ids == tuple(range(len(samples)))
        assert collate.h2d_bytes == sum(r.numel() for r in case["raw"])


@pytest.mark.parametrize("h,w,oh,ow,flip,ps,bgr", [
    (1024, 2048, 600, 1200, False, 3, True),      # the DA YAMLs' Cityscapes case (MIN_SIZE_TRAIN 600, MAX 1200)
    (1024, 2048, 600, 1200, True, 4, True),       # RGBX pixels (PIL's internal layout), mirrored
    (1024, 2048, 1024, 2048, True, 3, True),      # BASELINE's synthetic size: no resampling, flip + normalise + pad
    (333, 517, 800, 1242, False, 3, False),       # upsampling
    (900, 64, 31, 64, True, 3, True),             # 29x vertical reduction: tall shared-memory tiles, th < 16
    (8, 8, 1, 1, False, 3, True),
])
def test_kernel_matches_oracle_full_size(h, w, oh, ow, flip, ps, bgr):
    from dadetect_b200.data import DeviceTransform
    mean, std = (102.9801, 115.9465, 122.7717), (1.0, 1.0, 1.0)
    if not bgr:
        mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    rng = np.random.default_rng(h + w + oh)
    img = rng.integers(0, 256, (h, w, ps), dtype=np.uint8)
    want = pr.transform_image(np.ascontiguousarray(img[:, :, :3]), (oh, ow), flip, mean, std, bgr)
    hp, wp = -(-oh // 32) * 32, -(-ow // 32) * 32 + 32
    tf = DeviceTransform((oh,), None, 0.0, mean, std, bgr, DEV)
    dst = torch.full((3, hp, wp), float("nan"), device=DEV)
    tf.run(torch.from_numpy(img).to(DEV), (oh, ow), flip, dst)
    got = dst.cpu().numpy()
    assert np.array_equal(got[:, :oh, :ow], want)
    pad = got.copy()
    pad[:, :oh, :ow] = 0
    assert not np.isnan(got).any() and not pad.any()                        # every padding element written as zero


def test_flip_is_the_mirror_and_resize_commutes_with_it():
    """Size-independent property: Pillow's coefficient tables are symmetric, so transform(flip) of an image equals
    the mirrored transform of it; checked at the full Cityscapes size without the oracle."""
    from dadetect_b200.data import DeviceTransform
    tf = DeviceTransform((600,), 1200, 0.0, (102.9801, 115.9465, 122.7717), (1.0, 1.0, 1.0), True, DEV)
    g = torch.Generator().manual_seed(9)
    img = torch.randint(0, 256, (1024, 2048, 3), generator=g, dtype=torch.uint8).to(DEV)
    a = torch.empty((3, 608, 1216), device=DEV)
    b = torch.empty((3, 608, 1216), device=DEV)
    tf.run(img, (600, 1200), False, a)
    tf.run(img, (600, 1200), True, b)
    assert torch.equal(a[:, :600, :1200].flip(2), b[:, :600, :1200])
    c = torch.empty((3, 608, 1216), device=DEV)
    tf.run(img.flip(1).contiguous(), (600, 1200), False, c)
    assert torch.equal(c, b)


@pytest.mark.parametrize("min_size,max_size,padded", [(96, 192, (96, 192)), (90, 180, (96, 192))])
def test_pipeline_feeds_the_model(min_size, max_size, padded):
    """Raw uint8 samples -> DeviceBatchCollator -> GeneralizedRCNN.forward in training mode: finite losses.  The
    second case is the real-data situation of the DA YAMLs (600x1200 padded to 608x1216): the un-padded image size
    (anchor visibility, clipping) differs from the size of the batch tensor (feature-map extent)."""
    from dadetect_b200 import ops
    from dadetect_b200.data import DeviceBatchCollator, build_transforms
    from dadetect_b200.modeling import build_detection_model
    from dadetect_b200.structures import BoxList
    cfg = cfg_for(["INPUT.MIN_SIZE_TRAIN", (min_size,), "INPUT.MAX_SIZE_TRAIN", max_size,
                   "DATALOADER.SIZE_DIVISIBILITY", 32])
    ops.set_default_impl(ops.IMPL_TCGEN05)
    try:
        torch.manual_seed(3)
        model = build_detection_model(cfg).to(DEV)
        model.train()
        collate = DeviceBatchCollator(build_transforms(cfg, True, DEV), cfg.DATALOADER.SIZE_DIVISIBILITY)
        g = torch.Generator().manual_seed(2)
        samples = []
        for i in range(2):
            raw = torch.randint(0, 256, (160, 320, 3), generator=g, dtype=torch.uint8)
            x1 = torch.rand(4, generator=g) * 200
            y1 = torch.rand(4, generator=g) * 80
            t = BoxList(torch.stack([x1, y1, x1 + 40 + torch.rand(4, generator=g) * 60,
                                     y1 + 30 + torch.rand(4, generator=g) * 40], 1), (320, 160), mode="xyxy")
            t.add_field("labels", torch.randint(1, 9, (4,), generator=g))
            t.add_field("is_source", torch.full((4,), i == 0, dtype=torch.bool))
            samples.append((raw, t, i))
        random.seed(1)
        images, targets, _ = collate(samples)
        assert tuple(images.tensors.shape) == (2, 3) + padded
        assert [tuple(s) for s in images.image_sizes] == [(min_size, max_size)] * 2
        losses = model(images, [t.to(DEV) for t in targets])
        total = sum(losses.values())
        total.backward()
        assert torch.isfinite(total)
    finally:
        ops.set_default_impl(ops.IMPL_SIMT)


def test_triplet_collator_equals_three_planned_batches():
    """DeviceBatchCollatorTriplet returns the reference's 9-tuple; each of its three batches equals the oracle's
    transform of the same raw images under the same (size, flip) decisions."""
    from dadetect_b200.data import DeviceBatchCollatorTriplet, build_transforms
    cfg = cfg_for(["INPUT.MIN_SIZE_TRAIN", (40, 52), "INPUT.MAX_SIZE_TRAIN", 90, "DATALOADER.SIZE_DIVISIBILITY", 32])
    coll = DeviceBatchCollatorTriplet(build_transforms(cfg, True, DEV), 32)
    g = torch.Generator().manual_seed(12)
    batch = []
    for i in range(2):
        imgs = [torch.randint(0, 256, (70, 140, 3), generator=g, dtype=torch.uint8) for _ in range(3)]
        batch.append((imgs[0], None, imgs[1], None, imgs[2], None, "a%d" % i, "b%d" % i, "c%d" % i))
    random.seed(3)
    plans = coll.plan_batch(batch)
    random.seed(3)
    out = coll(batch)
    assert len(out) == 9 and out[6] == ("a0", "a1") and out[8] == ("c0", "c1")
    for v in range(3):
        tensors = [pr.transform_image(batch[i][2 * v].numpy(), plans[i][v][0], plans[i][v][1], cfg.INPUT.PIXEL_MEAN,
                                      cfg.INPUT.PIXEL_STD, cfg.INPUT.TO_BGR255) for i in range(2)]
        want, sizes = pr.collate(tensors, 32)
        assert np.array_equal(out[2 * v].tensors.cpu().numpy(), want)
        assert [tuple(s) for s in out[2 * v].image_sizes] == sizes
    assert coll.h2d_bytes == 6 * 70 * 140 * 3
