"""Host -> device input staging that overlaps with compute.

The reference feeds the GPU from a DataLoader and `images.to(device)` on the compute stream
(engine/trainer.py:213-224), so every step starts with a blocking 50 MB copy.  `DevicePrefetcher` keeps two
device-side batch slots and moves batch s+1 from pinned host memory on a dedicated copy stream while step s runs;
events order slot reuse against the consumer.  Every step still performs exactly one H2D copy of its inputs.
"""
import torch

from ..structures import BoxList


class DevicePrefetcher(object):
    def __init__(self, device, image_size, slots=2):
        self.device = torch.device(device)
        self.size = tuple(image_size)            # (W, H) of the BoxLists
        self.stream = torch.cuda.Stream(self.device)
        self.n = slots
        self.dev = [None] * slots                # (images, [(boxes, labels, is_source)])
        self.ready = [None] * slots              # copy finished (recorded on the copy stream)
        self.free = [None] * slots               # consumer finished (recorded on the compute stream)
        self.tag = [None] * slots

    def _alloc(self, slot, images, targets):
        cur = self.dev[slot]
        ok = cur is not None and cur[0].shape == images.shape and len(cur[1]) == len(targets) and all(
            c[0].shape == t["boxes"].shape for c, t in zip(cur[1], targets))
        if not ok:
            self.dev[slot] = (torch.empty(images.shape, dtype=images.dtype, device=self.device),
                              [(torch.empty(t["boxes"].shape, dtype=t["boxes"].dtype, device=self.device),
                                torch.empty(t["labels"].shape, dtype=t["labels"].dtype, device=self.device)) for t in targets])
        return self.dev[slot]

    def put(self, tag, images, targets):
        """Start the asynchronous copy of a pinned host batch (images [N,3,H,W], targets list of dicts with
        boxes / labels / is_source) into slot tag % slots."""
        slot = tag % self.n
        d_img, d_tg = self._alloc(slot, images, targets)
        with torch.cuda.stream(self.stream):
            if self.free[slot] is not None:
                self.stream.wait_event(self.free[slot])
            d_img.copy_(images, non_blocking=True)
            for (db, dl), t in zip(d_tg, targets):
                db.copy_(t["boxes"], non_blocking=True)
                dl.copy_(t["labels"], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.ready[slot] = ev
        self.tag[slot] = (tag, [bool(t["is_source"]) for t in targets])

    def get(self, tag):
        """Device batch of `tag` (must have been `put`); the current stream waits for its copy."""
        slot = tag % self.n
        assert self.tag[slot] is not None and self.tag[slot][0] == tag, "batch {} was not staged".format(tag)
        torch.cuda.current_stream().wait_event(self.ready[slot])
        d_img, d_tg = self.dev[slot]
        out = []
        for (db, dl), src in zip(d_tg, self.tag[slot][1]):
            b = BoxList(db, self.size, mode="xyxy")
            b.add_field("labels", dl)
            b._is_source_image = src              # known on the host: no device read to find the domain
            out.append(b)
        return d_img, out

    def release(self, tag):
        """Call after the consumer has enqueued its use of the batch; the slot may then be overwritten."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.free[tag % self.n] = ev
