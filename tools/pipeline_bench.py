"""Developer tool (GPU): the input-pipeline kernel (dd_preprocess_image) at the DA YAMLs' Cityscapes sizes.

  python tools/pipeline_bench.py            # kernel time, algorithmic GB/s, collator time from pinned host memory

Algorithmic bytes per image = source uint8 bytes read once + float32 bytes of the padded batch slot written once
(DESIGN §5).  The reference's CPU chain (PIL resize + ToTensor + Normalize + to_image_list) is timed beside it through
the oracle's dependencies (Pillow / torch), on this box's host cores.
"""
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dadetect_b200.data import DeviceBatchCollator, DeviceTransform

MEAN, STD = (102.9801, 115.9465, 122.7717), (1.0, 1.0, 1.0)
dev = torch.device("cuda")


def time_kernel(h, w, oh, ow, reps=50):
    tf = DeviceTransform((oh,), None, 0.0, MEAN, STD, True, dev)
    img = torch.randint(0, 256, (h, w, 3), dtype=torch.uint8, device=dev)
    hp, wp = -(-oh // 32) * 32, -(-ow // 32) * 32
    dst = torch.empty((3, hp, wp), device=dev)
    flush = torch.empty(64 * 1024 * 1024, device=dev)          # 256 MB > L2
    for _ in range(3):
        tf.run(img, (oh, ow), False, dst)
    total = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tf.run(img, (oh, ow), True, dst)
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    us = total / reps * 1e3
    nbytes = h * w * 3 + 3 * hp * wp * 4
    print("preprocess %dx%d -> %dx%d (pad %dx%d): %.1f us, %.0f GB/s algorithmic (%.2f MB)" % (
        h, w, oh, ow, hp, wp, us, nbytes / us / 1e3, nbytes / 1e6))


time_kernel(1024, 2048, 600, 1200)
time_kernel(1024, 2048, 1024, 2048)
time_kernel(1024, 2048, 800, 1600)

# collator: 2 decoded Cityscapes-sized images from host memory -> padded device batch (H2D inside)
collate = DeviceBatchCollator(DeviceTransform((600,), 1200, 0.5, MEAN, STD, True, dev), 32)
raws = [torch.randint(0, 256, (1024, 2048, 3), dtype=torch.uint8) for _ in range(2)]
random.seed(0)
for _ in range(3):
    collate([(r, None, i) for i, r in enumerate(raws)])
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    out = collate([(r, None, i) for i, r in enumerate(raws)])
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / 20 * 1e3
print("collator, 2 images 1024x2048 u8 host -> [2,3,608,1216] device: %.2f ms/batch (H2D %d bytes/batch)" % (
    ms, collate.h2d_bytes))

# the reference's CPU chain on the same images (Pillow + torch), one thread per image as a DataLoader worker would
from PIL import Image
import torchvision.transforms.functional as TF
torch.set_num_threads(1)
t0 = time.perf_counter()
for _ in range(3):
    ts = []
    for r in raws:
        im = Image.fromarray(r.numpy(), mode="RGB").resize((1200, 600), Image.BILINEAR)
        t = TF.to_tensor(im)[[2, 1, 0]] * 255
        ts.append(TF.normalize(t, MEAN, STD))
    batch = torch.zeros((2, 3, 608, 1216))
    for b, t in zip(batch, ts):
        b[:, :600, :1200].copy_(t)
ms_cpu = (time.perf_counter() - t0) / 3 * 1e3
print("reference CPU chain (PIL resize + ToTensor + Normalize + pad), same 2 images, 1 host thread: %.1f ms/batch" % ms_cpu)
