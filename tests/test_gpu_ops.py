"""GPU parity tests: every C-ABI kernel against the CPU oracle (oracle/da_frcnn_ref.py, pinned to the
reference by tests/test_oracle_pins.py) or a plain torch fp32 CPU reference for the dense ops.
Bit-exact for anchors / indices / match labels / NMS keep lists; stated tolerances for fp32 math."""
import os

import pytest
import torch
import torch.nn.functional as F

import da_frcnn_ref as orc

pytestmark = pytest.mark.gpu

DEV = "cuda"


def ops():
    from dadetect_b200 import ops as o
    return o


def rand_boxes(g, n, w, h, min_size=4.0):
    x1 = torch.rand(n, generator=g) * w * 0.8
    y1 = torch.rand(n, generator=g) * h * 0.8
    bw = min_size + torch.rand(n, generator=g) * w * 0.4
    bh = min_size + torch.rand(n, generator=g) * h * 0.4
    return torch.stack([x1, y1, (x1 + bw).clamp(max=w - 1), (y1 + bh).clamp(max=h - 1)], 1)


# ------------------------------------------------------------------------------ ROIAlign
@pytest.mark.parametrize("sampling_ratio", [0, 2])
@pytest.mark.parametrize("channels", [8, 6])
def test_roi_align_forward_backward(sampling_ratio, channels):
    g = torch.Generator().manual_seed(1)
    feat = torch.randn(2, channels, 12, 20, generator=g)
    rois = torch.cat([torch.randint(0, 2, (40, 1), generator=g).float(), rand_boxes(g, 40, 320, 192)], 1)
    rois[0, 1:] = torch.tensor([-40.0, -30.0, 500.0, 400.0])
    rois[1, 1:] = torch.tensor([50.0, 60.0, 50.2, 60.1])
    go = torch.randn(40, channels, 14, 14, generator=g)
    fr = feat.clone().requires_grad_(True)
    want = orc.roi_align(fr, rois, 1 / 16, 14, 14, sampling_ratio)
    (gwant,) = torch.autograd.grad(want, fr, go, retain_graph=True)

    fd = feat.permute(0, 2, 3, 1).contiguous().to(DEV).requires_grad_(True)
    got = ops().roi_align(fd, rois.to(DEV), 1 / 16, 14, sampling_ratio, 1)
    (ggot,) = torch.autograd.grad(got, fd, go.permute(0, 2, 3, 1).contiguous().to(DEV), retain_graph=True)
    torch.testing.assert_close(got.permute(0, 3, 1, 2).cpu(), want, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(ggot.permute(0, 3, 1, 2).cpu(), gwant, atol=1e-5, rtol=1e-4)

    # even-bin variant == the even bins of the full op (what res5's stride-2 1x1 convs read)
    got2 = ops().roi_align(fd, rois.to(DEV), 1 / 16, 14, sampling_ratio, 2)
    assert torch.equal(got2, got[:, ::2, ::2, :])
    go2 = torch.zeros_like(go)
    go2[:, :, ::2, ::2] = go[:, :, ::2, ::2]
    (gwant2,) = torch.autograd.grad(want, fr, go2)
    (ggot2,) = torch.autograd.grad(got2, fd, go[:, :, ::2, ::2].permute(0, 2, 3, 1).contiguous().to(DEV))
    torch.testing.assert_close(ggot2.permute(0, 3, 1, 2).cpu(), gwant2, atol=1e-5, rtol=1e-4)


def test_roi_align_nchw_c_abi_matches_reference_layout(golden_dir):
    """The `_C.roi_align_forward/backward` boundary: NCHW in, [K,C,PH,PW] out (csrc/ROIAlign.h:11-45),
    checked against outputs of the reference's own compiled CPU kernel (tests/golden/ref_ops.pt)."""
    from dadetect_b200 import _C
    o = torch.load(os.path.join(golden_dir, "ref_ops.pt"), weights_only=False)
    feat, rois = o["ra_feat"].to(DEV), o["ra_rois"].to(DEV)
    torch.testing.assert_close(_C.roi_align_forward(feat, rois, 1 / 16, 14, 14, 0).cpu(), o["ra_out_s0"],
                               atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(_C.roi_align_forward(feat, rois, 1 / 16, 7, 7, 2).cpu(), o["ra_out_s2"],
                               atol=1e-5, rtol=1e-5)
    g = torch.randn(rois.shape[0], feat.shape[1], 7, 7, generator=torch.Generator().manual_seed(2))
    fr = o["ra_feat"].clone().requires_grad_(True)
    (want,) = torch.autograd.grad(orc.roi_align(fr, o["ra_rois"], 1 / 16, 7, 7, 2), fr, g)
    got = _C.roi_align_backward(g.to(DEV), rois, 1 / 16, 7, 7, feat.shape[0], feat.shape[1], feat.shape[2],
                                feat.shape[3], 2)
    torch.testing.assert_close(got.cpu(), want, atol=1e-5, rtol=1e-4)


# ------------------------------------------------------------------------------ NMS
def test_nms_reference_kats(golden_dir):
    k = torch.load(os.path.join(golden_dir, "ref_kats.pt"), weights_only=False)
    for thr, want in zip(k["nms5_thresh"], k["nms5_keep"]):
        got = ops().nms(k["nms5_boxes"].to(DEV), k["nms5_scores"].to(DEV), thr)
        assert got.cpu().tolist() == sorted(want)
    got = ops().nms(k["nms53_boxes"].to(DEV), k["nms53_scores"].to(DEV), 0.5)
    assert got.cpu().tolist() == k["nms53_keep"].tolist()


@pytest.mark.parametrize("n,thr", [(1, 0.5), (63, 0.3), (64, 0.5), (65, 0.7), (3000, 0.7), (12000, 0.7), (2048, 0.5),
                                   (16384, 0.7)])
def test_nms_matches_oracle_bit_exact(n, thr):
    g = torch.Generator().manual_seed(n)
    boxes = rand_boxes(g, n, 2048, 1024, 8.0)
    if n > 100:      # cluster boxes so that a realistic fraction is suppressed
        centers = rand_boxes(g, 40, 2048, 1024, 30.0)
        boxes = centers[torch.randint(0, 40, (n,), generator=g)] + torch.randn(n, 4, generator=g) * 6
        boxes[:, 2:] = torch.maximum(boxes[:, 2:], boxes[:, :2] + 1)
    scores = torch.rand(n, generator=g)
    want = orc.nms(boxes, scores, thr, strict=True)
    got = ops().nms(boxes.to(DEV), scores.to(DEV), thr)
    assert torch.equal(got.cpu(), want)
    # sorted-input variant with early exit == boxlist_nms(max_proposals) (boxlist_ops.py:30-33)
    order = torch.sort(scores, descending=True, stable=True)[1]
    keep, cnt = ops().nms_sorted(boxes[order].contiguous().to(DEV), thr, 50)
    kept = order[keep[: int(cnt.item())].cpu()]
    want_first = order[torch.isin(order, want)][:50]
    assert torch.equal(kept, want_first)


def test_nms_batched_ragged_counts_match_single_image():
    """One batched call (device-side per-image counts, incl. an empty image) == per-image calls == oracle."""
    g = torch.Generator().manual_seed(5)
    cap, counts, thr, post = 3000, [3000, 1234, 0, 65], 0.7, 300
    centers = rand_boxes(g, 40, 2048, 1024, 30.0)
    batch = torch.zeros(len(counts), cap, 4)
    want = []
    for i, n in enumerate(counts):
        b = centers[torch.randint(0, 40, (cap,), generator=g)] + torch.randn(cap, 4, generator=g) * 6
        b[:, 2:] = torch.maximum(b[:, 2:], b[:, :2] + 1)
        batch[i] = b                                    # rows >= n are garbage the kernel must ignore
        scores = torch.arange(n, 0, -1, dtype=torch.float32)          # already in descending-score order
        want.append(orc.nms(b[:n], scores, thr, strict=True)[:post] if n else torch.zeros(0, dtype=torch.int64))
    keep, cnt = ops().nms_sorted_batched(batch.to(DEV), torch.tensor(counts, dtype=torch.int32, device=DEV), thr, post)
    cnt = cnt.cpu().tolist()
    for i in range(len(counts)):
        assert cnt[i] == len(want[i])
        assert torch.equal(keep[i, : cnt[i]].cpu(), want[i])


@pytest.mark.parametrize("n_cap,n,batch,max_pos,p_pos,p_neg,int_keys", [
    (5000, 4321, 256, 64, 0.1, 0.6, False), (122880, 122880, 256, 128, 0.0005, 0.9, False),
    (600, 600, 256, 128, 0.02, 0.1, False), (300, 150, 256, 64, 0.3, 0.3, True), (64, 64, 16, 4, 0.0, 0.5, True), (301, 301, 64, 16, 0.2, 0.5, True),
    # the cluster kernel (n_cap > 8192): massive ties across its eight slices, ragged device-side counts, the largest
    # supported set, fewer candidates than wanted
    (122880, 100001, 256, 128, 0.002, 0.5, True), (20000, 20000, 512, 128, 0.5, 0.4, True),
    (9001, 8999, 256, 128, 0.0, 0.9, False), (131072, 131072, 256, 128, 0.3, 0.3, True),
    (16384, 9, 256, 128, 0.5, 0.5, True), (40000, 40000, 256, 128, 0.0001, 0.0005, False)])
def test_balanced_sample_matches_reference_semantics(n_cap, n, batch, max_pos, p_pos, p_neg, int_keys):
    """Device sampler == BalancedPositiveNegativeSampler with `keys` standing for the random permutation:
    min(#pos, max_pos) positives and min(#neg, batch - num_pos) negatives with the smallest keys (ties -> lower
    index), returned as ascending indices; candidates beyond the device-side count are ignored."""
    g = torch.Generator().manual_seed(n_cap + batch)
    u = torch.rand(2, n_cap, generator=g)
    labels = torch.where(u < p_pos, 3, torch.where(u < p_pos + p_neg, 0, -1)).to(torch.int32)
    keys = torch.randint(0, 40, (2, n_cap), generator=g).float() if int_keys else torch.rand(2, n_cap, generator=g)
    counts = torch.tensor([n, max(n - 7, 1)], dtype=torch.int32)
    sel, cnt = ops().balanced_sample(labels.to(DEV), counts.to(DEV), keys.to(DEV), batch, max_pos)
    sel, cnt = sel.cpu(), cnt.cpu()
    for i in range(2):
        lab, key = labels[i, : counts[i]], keys[i, : counts[i]]
        pos, neg = torch.nonzero(lab >= 1).squeeze(1), torch.nonzero(lab == 0).squeeze(1)
        num_pos = min(pos.numel(), max_pos)
        num_neg = min(neg.numel(), batch - num_pos)
        pick = lambda idx, k: idx[torch.sort(key[idx], stable=True)[1][:k]]
        want = torch.sort(torch.cat([pick(pos, num_pos), pick(neg, num_neg)]))[0]
        assert cnt[i].tolist() == [num_pos, num_pos + num_neg]
        assert torch.equal(sel[i, : want.numel()], want)
        assert int(sel[i, want.numel():].abs().sum()) == 0


def test_balanced_sample_replayed_permutation_at_rpn_size():
    """Keys = rank in a recorded permutation (how the parity tests replay the oracle's draws), non-candidates 3e7:
    the cluster kernel picks exactly positive[perm[:k]] / negative[perm[:k]] (balanced_positive_negative_sampler.py:57-63)."""
    g = torch.Generator().manual_seed(5)
    n = 122880
    u = torch.rand(n, generator=g)
    labels = torch.where(u < 0.003, 1, torch.where(u < 0.7, 0, -1)).to(torch.int32)
    keys = torch.full((n,), 3.0e7)
    want = []
    for cond, k in ((labels >= 1, 128), (labels == 0, None)):
        idx = torch.nonzero(cond).squeeze(1)
        perm = torch.randperm(idx.numel(), generator=g)
        keys[idx[perm]] = torch.arange(idx.numel(), dtype=torch.float32)
        k = min(idx.numel(), 128) if k else 256 - min(int((labels >= 1).sum()), 128)
        want.append(idx[perm[:k]])
    want = torch.sort(torch.cat(want))[0]
    sel, cnt = ops().balanced_sample(labels.view(1, -1).to(DEV), None, keys.view(1, -1).to(DEV), 256, 128)
    assert int(cnt[0, 1]) == want.numel() == 256
    assert torch.equal(sel[0].cpu(), want)


def test_rpn_labels_and_fused_sampled_losses():
    """dd_rpn_anchor_labels + dd_rpn_sampled_losses (the RPN loss tail: BCE + smooth-L1 + both gradients in one launch)
    against the torch restatement of rpn/loss.py:118-141 used by the CPU suite; two source images + one target image,
    padded GT capacity."""
    import cpu_ops_emulation as emu
    g = torch.Generator().manual_seed(21)
    A, S, n_img, B, cap = 3000, 2, 3, 256, 16
    xy = torch.rand(A, 2, generator=g) * 300
    anchors = torch.cat([xy, xy + 20 + torch.rand(A, 2, generator=g) * 80], dim=1)
    gt_cat = torch.zeros(n_img * cap, 4)
    gt_off = torch.arange(0, (n_img + 1) * cap, cap, dtype=torch.int32)
    live = [5, 9, 3]
    for i in range(n_img):
        p = torch.rand(live[i], 2, generator=g) * 280
        gt_cat[i * cap: i * cap + live[i]] = torch.cat([p, p + 30 + torch.rand(live[i], 2, generator=g) * 60], dim=1)
    vis = (torch.rand(A, generator=g) < 0.8).to(torch.uint8)
    o = ops()
    labs, ms = [], []
    for s in range(S):
        m, _ = o.match(gt_cat[s * cap:(s + 1) * cap].to(DEV), anchors.to(DEV), 0.5, 0.3, True,
                       m_dev=torch.tensor([live[s]], dtype=torch.int32, device=DEV))
        lab = o.rpn_anchor_labels(m, vis.to(DEV))
        assert torch.equal(lab.cpu(), emu.rpn_anchor_labels(m.cpu(), vis))
        labs.append(lab)
        ms.append(m)
    lab, m = torch.stack(labs), torch.stack(ms)
    assert int((lab == 1).sum()) > 10 and int((lab == -1).sum()) > 10
    keys = torch.rand(S, A, generator=g).to(DEV)
    sel, cnt = o.balanced_sample(lab, None, keys, B, 128)
    logits = torch.randn(n_img, 10, 100, 3, generator=g)
    deltas = torch.randn(n_img, 10, 100, 12, generator=g) * 0.5
    src = torch.tensor([0, 1], dtype=torch.int32)
    a = [logits.clone().requires_grad_(True), deltas.clone().requires_grad_(True)]
    b = [logits.to(DEV).requires_grad_(True), deltas.to(DEV).requires_grad_(True)]
    want = emu.rpn_sampled_losses(a[0], a[1], anchors, sel.cpu(), cnt.cpu(), lab.cpu(), m.cpu(), gt_cat, gt_off, src, 1 / 9)
    got = o.rpn_sampled_losses(b[0], b[1], anchors.to(DEV), sel, cnt, lab, m, gt_cat.to(DEV), gt_off.to(DEV),
                               src.to(DEV), 1 / 9)
    for w, g_ in zip(want, got):
        assert abs(float(w) - float(g_)) <= 1e-5 * max(1.0, abs(float(w))), (float(w), float(g_))
    (want[0] * 0.7 + want[1] * 1.3).backward()
    (got[0] * 0.7 + got[1] * 1.3).backward()
    for x, y in zip(a, b):
        assert float((x.grad - y.grad.cpu()).abs().max()) <= 1e-6 + 1e-5 * float(x.grad.abs().max())
        assert float(x.grad.abs().sum()) > 0


def test_nms_sparse_suppression_stops_at_max_keep():
    """The benchmark's regime for the cluster kernel: 12 000 boxes spread over the image (few suppressions, 64 kept
    per block), the scan stops at POST_NMS_TOP_N = 2000 kept (boxlist_ops.py:30-33); two images, ragged counts."""
    g = torch.Generator().manual_seed(77)
    cap, counts, thr, post = 12000, [12000, 7001], 0.7, 2000
    batch = torch.zeros(2, cap, 4)
    want = []
    for i, n in enumerate(counts):
        xy = torch.rand(cap, 2, generator=g) * torch.tensor([1900.0, 900.0])
        wh = 20 + torch.rand(cap, 2, generator=g) * 300
        batch[i] = torch.cat([xy, xy + wh], dim=1)
        scores = torch.arange(n, 0, -1, dtype=torch.float32)
        want.append(orc.nms(batch[i, :n], scores, thr, strict=True)[:post])
    keep, cnt = ops().nms_sorted_batched(batch.to(DEV), torch.tensor(counts, dtype=torch.int32, device=DEV), thr, post)
    for i in range(2):
        assert int(cnt[i]) == len(want[i]) == post
        assert torch.equal(keep[i, :post].cpu(), want[i])


def test_roi_labels_and_gather_sampled_match_restatement():
    """dd_roi_labels + dd_roi_gather_sampled (the box head's sampling bookkeeping in two launches) against the torch
    restatement of box_head/loss.py:55-130 used by the CPU suite: source + target image, padded GT capacity."""
    import cpu_ops_emulation as emu
    g = torch.Generator().manual_seed(31)
    n_img, cap, B, gcap = 2, 500, 64, 8
    live = [5, 3]
    nprop = torch.tensor([480, 333], dtype=torch.int32)
    xy = torch.rand(n_img, cap, 2, generator=g) * 300
    boxes = torch.cat([xy, xy + 20 + torch.rand(n_img, cap, 2, generator=g) * 80], dim=2)
    obj = torch.rand(n_img, cap, generator=g)
    gt_cat = torch.zeros(n_img * gcap, 4)
    for i in range(n_img):
        gt_cat[i * gcap: i * gcap + live[i]] = boxes[i, 10: 10 + live[i]] + 3.0
    gt_off = torch.tensor([0, gcap, 2 * gcap], dtype=torch.int32)
    gt_counts = torch.tensor(live, dtype=torch.int32)
    gt_labels = torch.randint(1, 9, (n_img, gcap), generator=g)
    src = torch.tensor([1, 0], dtype=torch.uint8)
    o = ops()
    labs, ms = [], []
    for i in range(n_img):
        m, _ = o.match(gt_cat[i * gcap:(i + 1) * gcap].to(DEV), boxes[i].to(DEV), 0.5, 0.3, False,
                       m_dev=gt_counts[i:i + 1].to(DEV))
        lab = o.roi_labels(m, gt_labels[i].to(DEV), bool(src[i]), nprop[i:i + 1].to(DEV))
        assert torch.equal(lab.cpu(), emu.roi_labels(m.cpu(), gt_labels[i], bool(src[i]), nprop[i:i + 1]))
        labs.append(lab)
        ms.append(m)
    lab, m = torch.stack(labs), torch.stack(ms)
    assert int((lab[0] > 0).sum()) >= 3 and int((lab[0] == -1).sum()) >= cap - 480
    keys = torch.rand(n_img, cap, generator=g).to(DEV)
    sel, cnt = o.balanced_sample(lab, nprop.to(DEV), keys, B, 16)
    w = (10.0, 10.0, 5.0, 5.0)
    got = o.roi_gather_sampled(boxes.to(DEV), obj.to(DEV), sel, cnt, lab, m, gt_cat.to(DEV), gt_off.to(DEV),
                               gt_counts.to(DEV), src.to(DEV), w)
    want = emu.roi_gather_sampled(boxes, obj, sel.cpu(), cnt.cpu(), lab.cpu(), m.cpu(), gt_cat, gt_off, gt_counts, src, w)
    for k in want:
        a, b = got[k].cpu(), want[k]
        if a.dtype.is_floating_point:
            torch.testing.assert_close(a, b, atol=1e-5, rtol=1e-5)
        else:
            assert torch.equal(a, b), k


def test_proposals_gather_appends_gt_and_pads():
    g = torch.Generator().manual_seed(9)
    n, k, post, cap = 3, 50, 10, 14
    boxes, scores = torch.rand(n, k, 4, generator=g), torch.rand(n, k, generator=g)
    keep = torch.stack([torch.randperm(k, generator=g)[:post] for _ in range(n)])
    cnt = torch.tensor([10, 4, 0], dtype=torch.int32)
    gt = torch.rand(7, 4, generator=g)
    offs = torch.tensor([0, 3, 5, 7], dtype=torch.int32)
    app = torch.tensor([1, 0, 1], dtype=torch.uint8)
    ob, os_, oc = ops().proposals_gather(boxes.to(DEV), scores.to(DEV), keep.to(DEV), cnt.to(DEV), gt.to(DEV),
                                         offs.to(DEV), app.to(DEV), cap)
    ob, os_, oc = ob.cpu(), os_.cpu(), oc.cpu()
    assert oc.tolist() == [13, 4, 2]
    for i in range(n):
        want_b = boxes[i][keep[i, : cnt[i]]]
        want_s = scores[i][keep[i, : cnt[i]]]
        if app[i]:
            want_b = torch.cat([want_b, gt[offs[i]: offs[i + 1]]])
            want_s = torch.cat([want_s, torch.ones(offs[i + 1] - offs[i])])
        c = oc[i]
        assert torch.equal(ob[i, :c], want_b) and torch.equal(os_[i, :c], want_s)
        assert float(ob[i, c:].abs().sum()) == 0.0 and float(os_[i, c:].abs().sum()) == 0.0


def test_nms_empty():
    got = ops().nms(torch.zeros(0, 4, device=DEV), torch.zeros(0, device=DEV), 0.5)
    assert got.numel() == 0 and got.dtype == torch.int64


# ------------------------------------------------------------------------------ anchors / proposals
def test_anchor_grid_exact():
    cell = orc.cell_anchors(16, (32, 64, 128, 256, 512), (0.5, 1.0, 2.0))
    for fh, fw, iw, ih in ((12, 20, 320, 192), (64, 128, 2048, 1024)):
        want = orc.grid_anchors(fh, fw, 16, cell)
        vis = orc.anchor_visibility(want, iw, ih, 0)
        a, v = ops().anchor_grid(cell.to(DEV), fh, fw, 16, iw, ih, 0)
        assert torch.equal(a.cpu(), want)
        assert torch.equal(v.cpu().bool(), vis)


@pytest.mark.parametrize("fh,fw,k", [(12, 20, 2000), (64, 128, 12000), (5, 7, 525)])
def test_rpn_topk_decode(fh, fw, k):
    g = torch.Generator().manual_seed(fh * fw)
    a = 15
    n = 2
    # distinct values => no top-k ties (tie order is unspecified in the reference)
    total = n * a * fh * fw
    logits = ((torch.randperm(total, generator=g).float() / total - 0.5) * 8).view(n, a, fh, fw)
    deltas = torch.randn(n, 4 * a, fh, fw, generator=g) * 0.5
    deltas[0, 2] = 6.0          # exercises the log(1000/16) clamp on dw
    cell = orc.cell_anchors(16, (32, 64, 128, 256, 512), (0.5, 1.0, 2.0))
    anchors = orc.grid_anchors(fh, fw, 16, cell)
    iw, ih = fw * 16, fh * 16
    obj = orc.permute_and_flatten(logits, n, 1, fh, fw).view(n, -1)
    reg = orc.permute_and_flatten(deltas, n, 4, fh, fw)
    k = min(k, a * fh * fw)
    sc_want, idx_want = obj.sigmoid().topk(k, dim=1, sorted=True)
    boxes, scores, idx, valid = ops().rpn_topk_decode(logits.permute(0, 2, 3, 1).contiguous().to(DEV),
                                                      deltas.permute(0, 2, 3, 1).contiguous().to(DEV),
                                                      anchors.to(DEV), k, iw, ih, 0)
    assert valid.cpu().tolist() == [k, k]
    assert torch.equal(idx.cpu().long(), idx_want)            # no ties in the random logits
    torch.testing.assert_close(scores.cpu(), sc_want, atol=1e-6, rtol=1e-6)
    for i in range(n):
        want = orc.clip_boxes(orc.box_decode(reg[i][idx_want[i]], anchors[idx_want[i]], (1.0, 1.0, 1.0, 1.0)), iw, ih)
        torch.testing.assert_close(boxes[i].cpu(), want, atol=2e-3, rtol=1e-5)


@pytest.mark.parametrize("fh,fw,a,k,case", [(64, 128, 15, 12000, "ties"), (64, 128, 15, 12000, "one_region"),
                                            (41, 37, 15, 6000, "ragged"), (64, 128, 16, 16384, "largest"),
                                            (64, 128, 15, 12000, "min_size")])
def test_rpn_topk_decode_cluster_kernel_cases(fh, fw, a, k, case):
    """The 8-CTA cluster kernel (more than 16 384 anchors per image): heavily tied logits (the lowest indices of the
    ties at the cut win, and the output is ordered by (score desc, index asc) — what a stable descending sort gives),
    all high scores inside ONE CTA's share, an anchor count that is not a multiple of 4, the largest supported
    set, and a min-size filter that removes candidates from the middle of the ranking."""
    g = torch.Generator().manual_seed(fh + fw + k)
    n, A = 2, a * fh * fw
    if case == "ties":
        logits = torch.randint(-6, 7, (n, A), generator=g).float() * 0.25
    elif case == "one_region":
        logits = torch.randn(n, A, generator=g) - 10.0
        logits[:, 4096:4096 + 3000] += 20.0                       # one 4096-logit chunk holds the whole head of the ranking
    else:
        logits = torch.randn(n, A, generator=g) * 3
    deltas = torch.randn(n, A, 4, generator=g) * 0.5
    sizes = (32, 64, 128, 256, 512) if a == 15 else (32, 64, 128, 256, 512, 640, 700, 800)
    ratios = (0.5, 1.0, 2.0) if a == 15 else (0.5, 2.0)
    anchors = orc.grid_anchors(fh, fw, 16, orc.cell_anchors(16, sizes, ratios))
    assert anchors.shape[0] == A
    iw, ih = fw * 16, fh * 16
    min_size = 40.0 if case == "min_size" else 0.0
    boxes, scores, idx, valid = ops().rpn_topk_decode(logits.view(n, fh, fw, a).to(DEV),
                                                      deltas.view(n, fh, fw, 4 * a).to(DEV), anchors.to(DEV), k, iw, ih,
                                                      min_size)
    boxes, scores, idx, valid = boxes.cpu(), scores.cpu(), idx.cpu().long(), valid.cpu().tolist()
    for i in range(n):
        order = torch.sort(logits[i], descending=True, stable=True)[1][:k]          # ties: lower index first
        b = orc.clip_boxes(orc.box_decode(deltas[i][order], anchors[order], (1.0, 1.0, 1.0, 1.0)), iw, ih)
        # the kernel's own boxes decide its filter; the expectation uses them where the torch decode is within rounding
        keep = ((b[:, 2] - b[:, 0] + 1) >= min_size) & ((b[:, 3] - b[:, 1] + 1) >= min_size)
        want = order[keep]
        if case == "min_size":
            assert 0 < want.numel() < k
            assert abs(valid[i] - want.numel()) <= 3                # boxes within rounding of the threshold
            got = set(idx[i, : valid[i]].tolist())
            assert len(got ^ set(want.tolist())) <= 6
            pos = {v: j for j, v in enumerate(idx[i, : valid[i]].tolist())}
            common = [v for v in want.tolist() if v in pos]
            assert all(pos[x] < pos[y] for x, y in zip(common, common[1:]))          # order preserved
        else:
            assert valid[i] == k
            assert torch.equal(idx[i], want)
            torch.testing.assert_close(scores[i], logits[i][want].sigmoid(), atol=1e-6, rtol=1e-6)
            torch.testing.assert_close(boxes[i], b, atol=2e-3, rtol=1e-5)


def test_rpn_topk_ties_take_lowest_index_and_min_size_filter():
    fh, fw, a = 4, 4, 3
    logits = torch.zeros(1, fh, fw, a)
    logits[0, 1, 1, 1] = 3.0
    deltas = torch.zeros(1, fh, fw, 4 * a)
    anchors = orc.grid_anchors(fh, fw, 16, orc.cell_anchors(16, (32, 64, 128), (1.0,)))
    boxes, scores, idx, valid = ops().rpn_topk_decode(logits.to(DEV), deltas.to(DEV), anchors.to(DEV), 10, 64, 64, 0)
    assert idx[0].cpu().tolist() == [(1 * fw + 1) * a + 1] + list(range(9))
    _, _, _, valid2 = ops().rpn_topk_decode(logits.to(DEV), deltas.to(DEV), anchors.to(DEV), 10, 64, 64, 1000.0)
    assert valid2.cpu().tolist() == [0]


# ------------------------------------------------------------------------------ matcher / coder
@pytest.mark.parametrize("m,n", [(1, 5), (7, 300), (20, 122880)])
def test_match_bit_exact(m, n):
    g = torch.Generator().manual_seed(m * 1000 + n)
    gt = rand_boxes(g, m, 2048, 1024, 16.0).floor()
    if n == 122880:
        pred = orc.grid_anchors(64, 128, 16, orc.cell_anchors(16, (32, 64, 128, 256, 512), (0.5, 1.0, 2.0)))
    else:
        pred = rand_boxes(g, n, 2048, 1024)
        pred[: min(m, n)] = gt[: min(m, n)] + 0.25
    iou = orc.box_iou(gt, pred)
    for hi, lo, lq in ((0.7, 0.3, True), (0.5, 0.5, False)):
        want = orc.matcher(iou.clone(), hi, lo, lq)
        got, vals = ops().match(gt.to(DEV), pred.to(DEV), hi, lo, lq)
        assert torch.equal(got.cpu(), want)
        assert torch.equal(vals.cpu(), iou.max(dim=0)[0])


def test_match_raises_on_empty():
    with pytest.raises(ValueError):
        ops().match(torch.zeros(0, 4, device=DEV), torch.zeros(3, 4, device=DEV), 0.5, 0.5, False)
    with pytest.raises(ValueError):
        ops().match(torch.zeros(3, 4, device=DEV), torch.zeros(0, 4, device=DEV), 0.5, 0.5, False)


def test_box_encode_decode():
    g = torch.Generator().manual_seed(9)
    gt, pred = rand_boxes(g, 7, 320, 192), rand_boxes(g, 300, 320, 192)
    matches = torch.randint(-2, 7, (300,), generator=g)
    for wts in ((1.0, 1.0, 1.0, 1.0), (10.0, 10.0, 5.0, 5.0)):
        want = orc.box_encode(gt[matches.clamp(min=0)], pred, wts)
        got = ops().box_encode(gt.to(DEV), pred.to(DEV), matches.to(DEV), wts)
        torch.testing.assert_close(got.cpu(), want, atol=1e-5, rtol=1e-5)
        want_wrap = orc.box_encode(gt[matches], pred, wts)
        got_wrap = ops().box_encode(gt.to(DEV), pred.to(DEV), matches.to(DEV), wts, wrap_negative=True)
        torch.testing.assert_close(got_wrap.cpu(), want_wrap, atol=1e-5, rtol=1e-5)
        codes = torch.randn(300, 36, generator=g)
        want_d = orc.box_decode(codes, pred, wts)
        got_d = ops().box_decode(codes.to(DEV), pred.to(DEV), wts)
        torch.testing.assert_close(got_d.cpu(), want_d, atol=1e-3, rtol=1e-5)


def test_box_decode_reference_kat(golden_dir):
    k = torch.load(os.path.join(golden_dir, "ref_kats.pt"), weights_only=False)
    got = ops().box_decode(k["coder_deltas"].to(DEV), k["coder_boxes"].to(DEV), (1.0, 1.0, 1.0, 1.0))
    torch.testing.assert_close(got.cpu(), k["coder_decoded"], atol=1e-4, rtol=0)


# ------------------------------------------------------------------------------ dense tier
CONV_CASES = [
    # n, h, w, cin, cout, k, stride, pad
    (2, 16, 24, 64, 64, 1, 1, 0),
    (2, 16, 24, 64, 128, 1, 2, 0),
    (1, 15, 23, 32, 48, 3, 1, 1),
    (2, 32, 40, 3, 64, 7, 2, 3),
    (3, 7, 7, 128, 256, 3, 1, 1),
    (2, 8, 12, 256, 15, 1, 1, 0),
    (2, 8, 12, 64, 1, 1, 1, 0),
    (1, 9, 11, 16, 20, 3, 2, 1),
]


def _to_nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_forward_dgrad_wgrad_simt(case):
    n, h, w, cin, cout, k, stride, pad = case
    o = ops()
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    scale = 0.5 + torch.rand(cout, generator=g)
    bias = torch.randn(cout, generator=g) * 0.1
    xr, wr = x.clone().requires_grad_(True), wt.clone().requires_grad_(True)
    y0 = F.conv2d(xr, wr, stride=stride, padding=pad) * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
    res = torch.randn(y0.shape, generator=g)
    want = F.relu(y0 + res)
    go = torch.randn(want.shape, generator=g)
    gx_want, gw_want = torch.autograd.grad(want, (xr, wr), go)

    xd = _to_nhwc(x).to(DEV)
    wd = wt.permute(0, 2, 3, 1).contiguous().to(DEV)
    sd, bd, rd = scale.to(DEV), bias.to(DEV), _to_nhwc(res).to(DEV)
    got = o.conv2d_forward_raw(xd, wd, sd, bd, rd, k, k, stride, pad, True, impl=o.IMPL_SIMT)
    torch.testing.assert_close(got.permute(0, 3, 1, 2).cpu(), want, atol=2e-5, rtol=1e-4)
    gpre = o.relu_backward_raw(_to_nhwc(go).to(DEV), got)
    gx = o.conv2d_dgrad_raw(gpre, wd, sd, tuple(xd.shape), k, k, stride, pad, impl=o.IMPL_SIMT)
    torch.testing.assert_close(gx.permute(0, 3, 1, 2).cpu(), gx_want, atol=2e-5, rtol=1e-4)
    if cin % 4 == 0:
        gw = o.conv2d_wgrad_raw(gpre, xd, sd, cout, k, k, stride, pad, impl=o.IMPL_SIMT)
        torch.testing.assert_close(gw.permute(0, 3, 1, 2).cpu(), gw_want, atol=1e-4, rtol=1e-4)
    # fused dgrad epilogue: (acc + addend) * (mask_act > 0)
    addend = torch.randn(x.shape, generator=g)
    act = torch.randn(x.shape, generator=g)
    gx2 = o.conv2d_dgrad_raw(gpre, wd, sd, tuple(xd.shape), k, k, stride, pad, addend=_to_nhwc(addend).to(DEV),
                             mask_act=_to_nhwc(act).to(DEV), impl=o.IMPL_SIMT)
    torch.testing.assert_close(gx2.permute(0, 3, 1, 2).cpu(), (gx_want + addend) * (act > 0), atol=2e-5, rtol=1e-4)


def test_conv_autograd_function_and_linear():
    o = ops()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 32, 10, 12, generator=g)
    wt = torch.randn(48, 32, 3, 3, generator=g) * 0.05
    b = torch.randn(48, generator=g) * 0.1
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, wt, b))
    want = F.relu(F.conv2d(xr, wr, br, padding=1))
    go = torch.randn(want.shape, generator=g)
    gw = torch.autograd.grad(want, (xr, wr, br), go)
    xd = _to_nhwc(x).to(DEV).requires_grad_(True)
    wd = wt.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    bd = b.to(DEV).requires_grad_(True)
    got = o.conv_bn_act(xd, wd, None, bd, pad=1, relu=True)
    gg = torch.autograd.grad(got, (xd, wd, bd), _to_nhwc(go).to(DEV))
    torch.testing.assert_close(got.permute(0, 3, 1, 2).cpu(), want, atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(gg[0].permute(0, 3, 1, 2).cpu(), gw[0], atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(gg[1].cpu(), gw[1], atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(gg[2].cpu(), gw[2], atol=1e-4, rtol=1e-4)

    xl = torch.randn(37, 64, generator=g)
    wl = torch.randn(9, 64, generator=g) * 0.1
    bl = torch.randn(9, generator=g)
    xlr, wlr, blr = (t.clone().requires_grad_(True) for t in (xl, wl, bl))
    want = F.linear(xlr, wlr, blr)
    go = torch.randn(want.shape, generator=g)
    gw = torch.autograd.grad(want, (xlr, wlr, blr), go)
    xld, wld, bld = (t.to(DEV).requires_grad_(True) for t in (xl, wl, bl))
    got = o.linear(xld, wld, bld)
    gg = torch.autograd.grad(got, (xld, wld, bld), go.to(DEV))
    torch.testing.assert_close(got.cpu(), want, atol=2e-5, rtol=1e-4)
    for a, b_ in zip(gg, gw):
        torch.testing.assert_close(a.cpu(), b_, atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("impl", ["simt", "tcgen05"])
@pytest.mark.parametrize("first_stride,in_relu,premasked,pool", [(2, True, True, False), (1, False, False, False),
                                                                 (1, False, False, True)])
def test_fused_stage_matches_per_layer_autograd(impl, first_stride, in_relu, premasked, pool):
    """ops.bottleneck_stage (explicit backward, masks and fan-in adds in dgrad epilogues) against the same
    blocks run layer by layer through conv_bn_act + autograd."""
    from dadetect_b200.modeling.backbone import make_stage
    o = ops()
    o.set_default_impl(o.IMPL_TCGEN05 if impl == "tcgen05" else o.IMPL_SIMT)
    try:
        torch.manual_seed(3)
        stage = make_stage(32, 32, 64, 3, first_stride).to(DEV)
        for m in stage.modules():
            if hasattr(m, "running_var"):
                m.weight.uniform_(0.5, 1.5)
                m.bias.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
        x0 = torch.randn(2, 12, 20, 32, device=DEV)
        if in_relu:
            x0 = x0.relu()
        go = torch.randn((2, 64) if pool else (2, 12 // first_stride, 20 // first_stride, 64), device=DEV)
        res = []
        for fused in (False, True):
            stage.fused = fused
            x = x0.clone().requires_grad_(True)
            y = stage(x, input_is_relu=in_relu, grad_premasked=premasked, pool_output=pool)
            g = go * (y > 0) if premasked else go          # a pre-masked upstream gradient
            grads = torch.autograd.grad(y, [x] + list(stage.parameters()), g)
            gx = grads[0] * (x0 > 0) if (in_relu and not fused) else grads[0]
            res.append((y.detach(), gx, grads[1:]))
        tol = dict(atol=1e-5, rtol=1e-4) if impl == "simt" else dict(atol=2e-2, rtol=2e-2)
        torch.testing.assert_close(res[1][0], res[0][0], **tol)
        torch.testing.assert_close(res[1][1], res[0][1], **tol)
        for a, b in zip(res[1][2], res[0][2]):
            torch.testing.assert_close(a, b, atol=tol["atol"] * 10, rtol=tol["rtol"])
    finally:
        o.set_default_impl(o.IMPL_SIMT)


@pytest.mark.parametrize("shape", [(2, 64, 96), (1, 34, 50)])
def test_stem_tensor_core_path(shape):
    """7x7/2 stem as an overlapping-window implicit GEMM on tcgen05 (TF32) against F.conv2d fp32."""
    n, h, w = shape
    o = ops()
    g = torch.Generator().manual_seed(h)
    x = torch.rand(n, 3, h, w, generator=g) * 255 - 110
    wt = torch.randn(64, 3, 7, 7, generator=g) / 12.0
    scale = 0.5 + torch.rand(64, generator=g)
    bias = torch.randn(64, generator=g) * 0.1
    want = F.relu(F.conv2d(x, wt, stride=2, padding=3) * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1))
    got = o.stem_conv7x7s2(x.to(DEV), wt.to(DEV), scale.to(DEV), bias.to(DEV))
    err = got.permute(0, 3, 1, 2).cpu() - want
    rms = float(want.pow(2).mean().sqrt())
    assert float(err.pow(2).mean().sqrt()) <= 2e-3 * rms
    assert float(err.abs().max()) <= 6e-2 * rms


def test_pooling_and_layout():
    o = ops()
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 64, 17, 22, generator=g)
    torch.testing.assert_close(o.nchw_to_nhwc(x.to(DEV)).cpu(), _to_nhwc(x))
    torch.testing.assert_close(o.nhwc_to_nchw(_to_nhwc(x).to(DEV)).cpu(), x)
    want = F.max_pool2d(x, 3, 2, 1)
    assert torch.equal(o.maxpool3x3s2(_to_nhwc(x).to(DEV)).permute(0, 3, 1, 2).cpu(), want)
    r = torch.randn(5, 32, 7, 7, generator=g, requires_grad=True)
    want = F.avg_pool2d(r, 7).view(5, -1)
    go = torch.randn(5, 32, generator=g)
    (gw,) = torch.autograd.grad(want, r, go)
    rd = _to_nhwc(r.detach()).to(DEV).requires_grad_(True)
    got = o.avgpool_hw(rd)
    (gg,) = torch.autograd.grad(got, rd, go.to(DEV))
    torch.testing.assert_close(got.cpu(), want, atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(gg.permute(0, 3, 1, 2).cpu(), gw, atol=1e-7, rtol=1e-5)


# ------------------------------------------------------------------------------ losses
def test_bce_losses():
    o = ops()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 12, 20, 1, generator=g) * 3
    xr = x.clone().requires_grad_(True)
    want = orc.da_img_loss(xr.permute(0, 3, 1, 2), [True, False])
    (gw,) = torch.autograd.grad(want * 0.7, xr)
    xd = x.to(DEV).requires_grad_(True)
    got = o.bce_with_logits_mean(xd, None, torch.tensor([1, 0], dtype=torch.uint8, device=DEV), 240)
    (gg,) = torch.autograd.grad(got * 0.7, xd)
    assert abs(float(got) - float(want)) <= 1e-4 * abs(float(want))     # fp32-sum tier: 1e-4 rel
    torch.testing.assert_close(gg.cpu(), gw, atol=1e-7, rtol=1e-4)
    t = (torch.rand(77, generator=g) > 0.5).float()
    z = torch.randn(77, generator=g)
    zr = z.clone().requires_grad_(True)
    want = F.binary_cross_entropy_with_logits(zr, t)
    (gw,) = torch.autograd.grad(want, zr)
    zd = z.to(DEV).requires_grad_(True)
    got = o.bce_with_logits_mean(zd, t.to(DEV))
    (gg,) = torch.autograd.grad(got, zd)
    assert abs(float(got) - float(want)) <= 1e-4 * abs(float(want))
    torch.testing.assert_close(gg.cpu(), gw, atol=1e-7, rtol=1e-4)


def test_fastrcnn_losses():
    o = ops()
    g = torch.Generator().manual_seed(6)
    rows, c = 300, 9
    logits = torch.randn(rows, c, generator=g)
    reg = torch.randn(rows, 4 * c, generator=g)
    labels = torch.randint(0, c, (rows,), generator=g)
    regt = torch.randn(rows, 4, generator=g) * 0.5
    dom = torch.cat([torch.ones(180, dtype=torch.bool), torch.zeros(120, dtype=torch.bool)])
    samples = [dict(labels=labels, regression_targets=regt, domain_labels=dom)]
    lr, rr = logits.clone().requires_grad_(True), reg.clone().requires_grad_(True)
    wc, wb, _ = orc.fastrcnn_loss(lr, rr, samples)
    gwc, gwb = torch.autograd.grad(wc + wb, (lr, rr))
    ld, rd = logits.to(DEV).requires_grad_(True), reg.to(DEV).requires_grad_(True)
    gc = o.softmax_ce_mean(ld, labels.to(DEV), dom.to(torch.uint8).to(DEV))
    gb = o.box_reg_loss(rd, regt.to(DEV), labels.to(DEV), dom.to(torch.uint8).to(DEV))
    ggc, ggb = torch.autograd.grad(gc + gb, (ld, rd))
    assert abs(float(gc) - float(wc)) <= 1e-4 * abs(float(wc))
    assert abs(float(gb) - float(wb)) <= 1e-4 * abs(float(wb))
    torch.testing.assert_close(ggc.cpu(), gwc, atol=1e-7, rtol=1e-4)
    torch.testing.assert_close(ggb.cpu(), gwb, atol=1e-7, rtol=1e-4)


def test_smooth_l1_consistency_triplet():
    o = ops()
    g = torch.Generator().manual_seed(7)
    x, t = torch.randn(50, 4, generator=g), torch.randn(50, 4, generator=g) * 0.2
    xr = x.clone().requires_grad_(True)
    want = orc.smooth_l1(xr, t, 1 / 9, False) / 256
    (gw,) = torch.autograd.grad(want, xr)
    xd = x.to(DEV).requires_grad_(True)
    got = o.smooth_l1_sum(xd, t.to(DEV), 1 / 9, 256.0)
    (gg,) = torch.autograd.grad(got, xd)
    assert abs(float(got) - float(want)) <= 1e-4 * abs(float(want))
    torch.testing.assert_close(gg.cpu(), gw, atol=1e-7, rtol=1e-4)
    assert float(o.smooth_l1_sum(torch.zeros(0, 4, device=DEV), torch.zeros(0, 4, device=DEV), 1 / 9, 256.0)) == 0.0

    img = torch.randn(2, 1, 12, 20, generator=g)
    ins = torch.randn(37, 1, generator=g)
    dom = torch.cat([torch.ones(21, dtype=torch.bool), torch.zeros(16, dtype=torch.bool)])
    ir, sr = img.clone().requires_grad_(True), ins.clone().requires_grad_(True)
    want = orc.consistency_loss(ir.sigmoid(), sr.sigmoid(), dom)
    gwi, gws = torch.autograd.grad(want, (ir, sr))
    idv, sdv = img.reshape(2, -1).to(DEV).requires_grad_(True), ins.reshape(-1).to(DEV).requires_grad_(True)
    got = o.consistency_loss(idv, sdv, 21)
    ggi, ggs = torch.autograd.grad(got, (idv, sdv))
    assert abs(float(got) - float(want)) <= 1e-4 * abs(float(want))
    torch.testing.assert_close(ggi.cpu().view_as(gwi), gwi, atol=1e-8, rtol=1e-3)
    torch.testing.assert_close(ggs.cpu().view_as(gws), gws, atol=1e-8, rtol=1e-3)

    # image-level triplet: NCHW [1,C,H,W], distance over W
    a, p, n = (torch.randn(1, 8, 5, 16, generator=g) for _ in range(3))
    ar, pr, nr = (v.clone().requires_grad_(True) for v in (a, p, n))
    want = orc.triplet_margin_loss(ar, pr, nr, 1.0)
    gw = torch.autograd.grad(want, (ar, pr, nr))
    ad, pd, nd = (_to_nhwc(v).to(DEV).requires_grad_(True) for v in (a, p, n))
    got = o.triplet_margin_loss(ad, pd, nd, 1.0, 5 * 8, 16, 8)
    gg = torch.autograd.grad(got, (ad, pd, nd))
    assert abs(float(got) - float(want)) <= 1e-4 * abs(float(want))
    for x1, x2 in zip(gg, gw):
        torch.testing.assert_close(x1.permute(0, 3, 1, 2).cpu(), x2, atol=1e-7, rtol=1e-3)
    # instance-level triplet: [R, D]
    a, p, n = (torch.randn(9, 64, generator=g) for _ in range(3))
    ar, pr, nr = (v.clone().requires_grad_(True) for v in (a, p, n))
    want = orc.triplet_margin_loss(ar, pr, nr, 0.7)
    gw = torch.autograd.grad(want, (ar, pr, nr))
    ad, pd, nd = (v.to(DEV).requires_grad_(True) for v in (a, p, n))
    got = o.triplet_margin_loss(ad, pd, nd, 0.7, 9, 64, 1)
    gg = torch.autograd.grad(got, (ad, pd, nd))
    assert abs(float(got) - float(want)) <= 1e-4 * abs(float(want))
    for x1, x2 in zip(gg, gw):
        torch.testing.assert_close(x1.cpu(), x2, atol=1e-7, rtol=1e-3)


def test_grl_dropout_advgrl_sgd(golden_dir):
    o = ops()
    g = torch.Generator().manual_seed(8)
    x = torch.randn(33, 7, generator=g)
    xd = x.to(DEV).requires_grad_(True)
    y = o.gradient_scalar(xd, -0.1)
    assert torch.equal(y, xd)
    (gx,) = torch.autograd.grad(y, xd, torch.ones_like(y))
    torch.testing.assert_close(gx.cpu(), torch.full_like(x, -0.1))
    keep = (torch.rand(33, 7, generator=g) > 0.5).float()
    yd = o.dropout_with_mask(xd, keep.to(DEV))
    torch.testing.assert_close(yd.cpu(), x * keep * 2)
    ref = torch.load(os.path.join(golden_dir, "ref_ops.pt"), weights_only=False)
    for L, w in ref["adv_grl"]:                       # weights produced by the reference's Adv_GRL
        wd = o.adv_grl_weight(torch.tensor([L], device=DEV), ref["adv_bce"], 0.1, 0.1, 30)
        assert abs(float(wd) - w) <= 1e-6 * max(1.0, abs(w))
    wdev = torch.tensor([-2.5], device=DEV)
    y = o.gradient_scalar_dev(xd, wdev)
    (gx,) = torch.autograd.grad(y, xd, torch.ones_like(y))
    torch.testing.assert_close(gx.cpu(), torch.full_like(x, -2.5))

    p = torch.randn(1000, generator=g)
    grads = [torch.randn(1000, generator=g) for _ in range(3)]
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.SGD([pr], lr=0.01, momentum=0.9, weight_decay=5e-4)
    pd, buf = p.to(DEV), torch.zeros(1000, device=DEV)
    for i, gr in enumerate(grads):
        pr.grad = gr.clone()
        opt.step()
        o.sgd_momentum_(pd, gr.to(DEV), buf, 0.01, 0.9, 5e-4, 1.0, i == 0)
    torch.testing.assert_close(pd.cpu(), pr.detach(), atol=1e-6, rtol=1e-5)


# ------------------------------------------------------------------------------ dense tier, tcgen05 arm
TC_CASES = [
    (2, 16, 24, 64, 64, 1, 1, 0),
    (2, 16, 24, 64, 128, 1, 2, 0),
    (1, 15, 23, 32, 48, 3, 1, 1),
    (3, 7, 7, 128, 256, 3, 1, 1),
    (2, 8, 12, 256, 15, 1, 1, 0),
    (37, 1, 1, 2048, 1024, 1, 1, 0),
    (5, 7, 7, 512, 512, 3, 1, 1),
    (1, 32, 64, 256, 256, 3, 1, 1),
    (2, 128, 128, 64, 64, 1, 1, 0),       # 256 flat tiles: the persistent loop wraps (> 148 CTAs' worth)
    (2, 32, 48, 256, 128, 1, 2, 0),       # strided TMA views (forward / wgrad), scattered TMA store (dgrad)
    (2, 16, 24, 128, 60, 1, 1, 0),        # Cout % 32 != 0: the last staged chunk is clipped by the TMA store
    (1, 24, 40, 64, 512, 3, 1, 1),        # several column tiles per pixel tile
    (300, 7, 7, 64, 96, 1, 1, 0),         # 1x1 on ROI maps: flat pixel axis, ragged last tile
    (41, 7, 7, 64, 128, 3, 1, 1),         # 3x3 on ROI maps (3xTF32: dense pixel axis, taps as row offsets), ragged last tile
    (3, 5, 9, 64, 64, 3, 1, 1),           # the same on non-square maps
    (1, 7, 7, 32, 64, 3, 1, 1),           # one map: a single partial tile
]


@pytest.mark.timeout(300)
@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tcgen05_arm(case):
    """tcgen05 (TF32 operands, fp32 accumulate) against the fp32 CPU reference.  TF32 keeps 10 mantissa bits,
    so the tolerance is statistical: rms error <= 2e-3 of the rms magnitude, max error <= 2e-2 of it."""
    n, h, w, cin, cout, k, stride, pad = case
    o = ops()
    assert o.tcgen05_available()
    g = torch.Generator().manual_seed(sum(case) + 1)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    scale = 0.5 + torch.rand(cout, generator=g)
    bias = torch.randn(cout, generator=g) * 0.1
    xr, wr = x.clone().requires_grad_(True), wt.clone().requires_grad_(True)
    y0 = F.conv2d(xr, wr, stride=stride, padding=pad) * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
    res = torch.randn(y0.shape, generator=g)
    want = F.relu(y0 + res)
    go = torch.randn(want.shape, generator=g)
    gx_want, gw_want = torch.autograd.grad(want, (xr, wr), go)

    def close(got, ref, what):
        err = got - ref
        rms = float(ref.pow(2).mean().sqrt())
        assert float(err.pow(2).mean().sqrt()) <= 2e-3 * rms, (what, float(err.pow(2).mean().sqrt()), rms)
        assert float(err.abs().max()) <= 2e-2 * max(rms, 1e-6) * 3, (what, float(err.abs().max()), rms)

    xd = _to_nhwc(x).to(DEV)
    wd = wt.permute(0, 2, 3, 1).contiguous().to(DEV)
    sd, bd, rd = scale.to(DEV), bias.to(DEV), _to_nhwc(res).to(DEV)
    got = o.conv2d_forward_raw(xd, wd, sd, bd, rd, k, k, stride, pad, True, impl=o.IMPL_TCGEN05)
    close(got.permute(0, 3, 1, 2).cpu(), want.detach(), "forward")
    exact = o.conv2d_forward_raw(xd, wd, sd, bd, rd, k, k, stride, pad, True, impl=o.IMPL_SIMT)
    gpre = o.relu_backward_raw(_to_nhwc(go).to(DEV), exact)
    addend = torch.randn(x.shape, generator=g)
    act = torch.randn(x.shape, generator=g)
    gx = o.conv2d_dgrad_raw(gpre, wd, sd, tuple(xd.shape), k, k, stride, pad, addend=_to_nhwc(addend).to(DEV),
                            mask_act=_to_nhwc(act).to(DEV), impl=o.IMPL_TCGEN05)
    close(gx.permute(0, 3, 1, 2).cpu(), (gx_want + addend) * (act > 0), "dgrad")
    gw = o.conv2d_wgrad_raw(gpre, xd, sd, cout, k, k, stride, pad, impl=o.IMPL_TCGEN05)
    close(gw.permute(0, 3, 1, 2).cpu(), gw_want, "wgrad")


@pytest.mark.timeout(300)
@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tcgen05_x3_arm_is_fp32_grade(case):
    """3xTF32 on the tensor cores (operands split into TF32-exact hi / lo parts in the kernel): forward and data
    gradient agree with the fp64 reference to 5e-6 of the rms magnitude — 100x tighter than TF32 (5e-4..1e-3).
    The tensor core's accumulator rounds toward zero on every accumulation step (a bias that grows with the chain
    length: 1e-5 at K = 1152 with a single accumulator), so the kernel spreads a tile over 8 partial accumulators
    in TMEM and adds them with round-to-nearest in the epilogue."""
    n, h, w, cin, cout, k, stride, pad = case
    o = ops()
    g = torch.Generator().manual_seed(sum(case) + 2)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    scale = 0.5 + torch.rand(cout, generator=g)
    bias = torch.randn(cout, generator=g) * 0.1
    xr = x.clone().double().requires_grad_(True)
    y0 = F.conv2d(xr, wt.double(), stride=stride, padding=pad) * scale.double().view(1, -1, 1, 1) + bias.double().view(1, -1, 1, 1)
    res = torch.randn(y0.shape, generator=g)
    want = F.relu(y0 + res.double())
    go = torch.randn(want.shape, generator=g)
    (gx_want,) = torch.autograd.grad(want, (xr,), go.double())

    def close(got, ref, what):
        err = got.double() - ref
        rms = float(ref.pow(2).mean().sqrt())
        assert float(err.pow(2).mean().sqrt()) <= 5e-6 * rms, (what, float(err.pow(2).mean().sqrt()), rms)
        assert float(err.abs().max()) <= 1e-4 * max(rms, 1e-6), (what, float(err.abs().max()), rms)

    xd = _to_nhwc(x).to(DEV)
    wd = wt.permute(0, 2, 3, 1).contiguous().to(DEV)
    sd, bd, rd = scale.to(DEV), bias.to(DEV), _to_nhwc(res).to(DEV)
    got = o.conv2d_forward_raw(xd, wd, sd, bd, rd, k, k, stride, pad, True, impl=o.IMPL_TCGEN05_X3)
    close(got.permute(0, 3, 1, 2).cpu(), want.detach(), "forward")
    gpre = _to_nhwc((go.double() * (want > 0)).float()).to(DEV)
    addend = torch.randn(x.shape, generator=g)
    act = torch.randn(x.shape, generator=g)
    gx = o.conv2d_dgrad_raw(gpre, wd, sd, tuple(xd.shape), k, k, stride, pad, addend=_to_nhwc(addend).to(DEV),
                            mask_act=_to_nhwc(act).to(DEV), impl=o.IMPL_TCGEN05_X3)
    close(gx.permute(0, 3, 1, 2).cpu(), (gx_want.detach() + addend.double()) * (act > 0), "dgrad")
    wr = wt.clone().double().requires_grad_(True)
    y1 = F.conv2d(x.double(), wr, stride=stride, padding=pad) * scale.double().view(1, -1, 1, 1)
    (gw_want,) = torch.autograd.grad(y1, (wr,), (go.double() * (want > 0)))
    gw = o.conv2d_wgrad_raw(gpre, xd, sd, cout, k, k, stride, pad, impl=o.IMPL_TCGEN05_X3)
    close(gw.permute(0, 3, 1, 2).cpu(), gw_want, "wgrad")


# The shapes bench.py actually times (BASELINE configs[1..3] at 2 x 1024 x 2048, 512 ROIs): the dominant GEMM
# (RPN 3x3 1024 -> 1024 on the 64 x 128 map, K = 9216), res5 on 7x7 ROI maps, the memory-bound res2 expansion, the
# fused 96-column RPN predictor.  Reference: torch fp64 convolution on the same GPU (a stock op used as the checker).
TC_BENCH_CASES = [
    (2, 64, 128, 1024, 1024, 3, 1, 1),
    (512, 7, 7, 512, 2048, 1, 1, 0),
    (2, 256, 512, 64, 256, 1, 1, 0),
    (512, 7, 7, 512, 512, 3, 1, 1),
    (2, 64, 128, 1024, 96, 1, 1, 0),
    (512, 7, 7, 1024, 512, 1, 1, 0),
]


@pytest.mark.timeout(600)
@pytest.mark.parametrize("arm", ["tcgen05", "tcgen05x3"])
@pytest.mark.parametrize("case", TC_BENCH_CASES)
def test_conv_tcgen05_at_bench_shapes(case, arm):
    """Forward (+BN scale/bias, residual, ReLU), data gradient (+addend, ReLU mask) and weight gradient of both
    tensor-core arms at the full-size shapes of the benchmarked step: TF32 within 2e-3 rms of the fp64 result,
    3xTF32 within 5e-6 rms (fp32 grade; 3e-5 on the weight gradient, a sum over up to 262 144 pixels whose fp32
    partial sums meet through red.add in L2 — measured 1.2e-5 .. 1.5e-5)."""
    n, h, w, cin, cout, k, stride, pad = case
    o = ops()
    impl = o.IMPL_TCGEN05 if arm == "tcgen05" else o.IMPL_TCGEN05_X3
    g = torch.Generator(device=DEV).manual_seed(sum(case) + 3)
    x = torch.randn(n, h, w, cin, generator=g, device=DEV)
    wt = torch.randn(cout, k, k, cin, generator=g, device=DEV) / (cin * k * k) ** 0.5
    scale = 0.5 + torch.rand(cout, generator=g, device=DEV)
    bias = torch.randn(cout, generator=g, device=DEV) * 0.1
    xr = x.permute(0, 3, 1, 2).double().requires_grad_(True)
    wr = wt.permute(0, 3, 1, 2).double().requires_grad_(True)
    y0 = F.conv2d(xr, wr, stride=stride, padding=pad) * scale.double().view(1, -1, 1, 1) + bias.double().view(1, -1, 1, 1)
    res = torch.randn(n, y0.shape[2], y0.shape[3], cout, generator=g, device=DEV)
    want = F.relu(y0 + res.permute(0, 3, 1, 2).double())
    go = torch.randn(res.shape, generator=g, device=DEV)
    gpre = (go * (want.permute(0, 2, 3, 1) > 0)).contiguous()                        # NHWC fp32, already ReLU-masked
    gx_want, gw_want = torch.autograd.grad(want, (xr, wr), go.permute(0, 3, 1, 2).double())
    tol_rms = 2e-3 if arm == "tcgen05" else 5e-6

    def close(got_nhwc, ref_nchw, what, rms_tol):
        err = got_nhwc.permute(0, 3, 1, 2).double() - ref_nchw
        rms = float(ref_nchw.pow(2).mean().sqrt())
        assert float(err.pow(2).mean().sqrt()) <= rms_tol * rms, (what, float(err.pow(2).mean().sqrt()), rms)
        assert float(err.abs().max()) <= 30 * rms_tol * max(rms, 1e-6), (what, float(err.abs().max()), rms)

    got = o.conv2d_forward_raw(x, wt, scale, bias, res, k, k, stride, pad, True, impl=impl)
    close(got, want.detach(), "forward", tol_rms)
    del got, y0
    addend = torch.randn(x.shape, generator=g, device=DEV)
    act = torch.randn(x.shape, generator=g, device=DEV)
    gx = o.conv2d_dgrad_raw(gpre, wt, scale, tuple(x.shape), k, k, stride, pad, addend=addend, mask_act=act, impl=impl)
    close(gx, (gx_want + addend.permute(0, 3, 1, 2).double()) * (act.permute(0, 3, 1, 2) > 0), "dgrad", tol_rms)
    del gx, gx_want
    gw = o.conv2d_wgrad_raw(gpre, x, scale, cout, k, k, stride, pad, impl=impl)
    # d(want)/d(wr) already carries the BN scale (it multiplies the conv output)
    close(gw, gw_want, "wgrad", tol_rms if arm == "tcgen05" else 3e-5)


def test_match_and_encode_with_padded_gt_and_device_count():
    """GT boxes padded to a fixed capacity with the live count on the device (signature-free step graph): the
    Matcher (incl. the low-quality restore, which would otherwise fire on EVERY anchor for an all-zero padding
    row), the negative-index wrap of the encoder and the GT append of the proposals see only the live rows."""
    o = ops()
    g = torch.Generator().manual_seed(77)
    cap = 32
    pred = orc.grid_anchors(16, 24, 16, orc.cell_anchors(16, (32, 64, 128, 256, 512), (0.5, 1.0, 2.0)))
    for m in (1, 5, 32):
        gt = rand_boxes(g, m, 384, 256, 16.0).floor()
        padded = torch.zeros(cap, 4)
        padded[:m] = gt
        padded[m:] = torch.rand(cap - m, 4, generator=g) * 100          # stale rows of an earlier batch
        cnt = torch.tensor([m], dtype=torch.int32, device=DEV)
        iou = orc.box_iou(gt, pred)
        for hi, lo, lq in ((0.7, 0.3, True), (0.5, 0.5, False)):
            want = orc.matcher(iou.clone(), hi, lo, lq)
            got, vals = o.match(padded.to(DEV), pred.to(DEV), hi, lo, lq, m_dev=cnt)
            assert torch.equal(got.cpu(), want)
            assert torch.equal(vals.cpu(), iou.max(dim=0)[0])
        matches = torch.randint(-2 if m >= 2 else -1, m, (pred.shape[0],), generator=g)
        want_wrap = orc.box_encode(gt[matches], pred, (10.0, 10.0, 5.0, 5.0))
        got_wrap = o.box_encode(padded.to(DEV), pred.to(DEV), matches.to(DEV), (10.0, 10.0, 5.0, 5.0),
                                wrap_negative=True, m_dev=cnt)
        torch.testing.assert_close(got_wrap.cpu(), want_wrap, atol=1e-5, rtol=1e-5)
    # proposals_gather with per-image live counts inside fixed-capacity GT rows
    n, k, post = 2, 40, 8
    boxes, scores = torch.rand(n, k, 4, generator=g), torch.rand(n, k, generator=g)
    keep = torch.stack([torch.randperm(k, generator=g)[:post] for _ in range(n)])
    cnt = torch.tensor([8, 3], dtype=torch.int32)
    gt = torch.rand(2 * cap, 4, generator=g)
    offs = torch.tensor([0, cap, 2 * cap], dtype=torch.int32)
    live = torch.tensor([5, 9], dtype=torch.int32)
    app = torch.tensor([1, 1], dtype=torch.uint8)
    ob, os_, oc = o.proposals_gather(boxes.to(DEV), scores.to(DEV), keep.to(DEV), cnt.to(DEV), gt.to(DEV), offs.to(DEV),
                                     app.to(DEV), post + cap, gt_counts=live.to(DEV))
    assert oc.cpu().tolist() == [13, 12]
    for i in range(n):
        want_b = torch.cat([boxes[i][keep[i, : cnt[i]]], gt[i * cap: i * cap + live[i]]])
        assert torch.equal(ob[i, : oc[i]].cpu(), want_b)
        assert float(ob[i, oc[i]:].abs().sum()) == 0.0


def test_adaptive_margin_update_follows_reference_rule():
    """dd_adaptive_margin_update against the reference's host rule (da_heads/loss.py:182-200, restated in
    oracle._adaptive_margin) over a sequence of previous losses, including the stop at int(margin) == int(max)."""
    o = ops()
    for margin_cfg, max_margin in ((0.5, 3.0), (0.998, 1.5), (0.0, 3.0), (1.0, 1.0)):
        state = torch.zeros(1, dtype=torch.float64, device=DEV)
        cur = 0.0
        prevs = [1.0, 0.0, 0.0, 0.3, 0.0, 0.0, 0.0, 2.0, 0.0]
        for prev in prevs:
            cur = orc._adaptive_margin(cur, prev, True, 0.001, max_margin, margin_cfg)
            out = o.adaptive_margin_update(state, torch.tensor([prev], dtype=torch.float32, device=DEV), margin_cfg, 0.001,
                                           max_margin)
            assert float(state) == cur, (margin_cfg, max_margin, float(state), cur)
            assert float(out) == float(torch.tensor(cur, dtype=torch.float32))
    # and the loss kernel reads the margin from the device
    g = torch.Generator().manual_seed(3)
    a, p, n = (torch.randn(64, 48, generator=g).to(DEV) for _ in range(3))
    m = torch.tensor([0.75], dtype=torch.float32, device=DEV)
    l_dev = o.triplet_margin_loss(a, p, n, m, 64, 48, 1)
    l_host = o.triplet_margin_loss(a, p, n, 0.75, 64, 48, 1)
    assert abs(float(l_dev) - float(l_host)) <= 1e-6 * float(l_host)      # (atomic partial sums: order may differ)


@pytest.mark.parametrize("impl", ["tcgen05", "tcgen05x3"])
def test_dgrad_batched_weight_preparation_is_bit_identical(impl):
    """dd_conv2d_dgrad_prepare_batch (one launch for a whole stage's layers) followed by prepared dgrad calls gives
    exactly what the per-call preparation gives — 18 layers, i.e. two launches, scale present and absent."""
    o = ops()
    code = o.IMPL_TCGEN05 if impl == "tcgen05" else o.IMPL_TCGEN05_X3
    g = torch.Generator().manual_seed(5)
    shapes = [(64, 32, 1), (32, 64, 3), (128, 32, 1), (40, 24, 3), (256, 64, 1), (32, 96, 1)] * 3
    layers, gys = [], []
    for i, (cout, cin, k) in enumerate(shapes):
        w = (torch.randn(cout, k, k, cin, generator=g) / (cin * k * k) ** 0.5).to(DEV)
        s = (0.5 + torch.rand(cout, generator=g)).to(DEV) if i % 2 == 0 else None
        layers.append((w, s))
        gys.append(torch.randn(2, 9, 13, cout, generator=g).to(DEV))
    prepared = o.dgrad_prepare_batch(layers, torch.device(DEV), impl=code)
    assert len(prepared) == len(layers) and all(p is not None for p in prepared)
    for (w, s), gy, ws, (cout, cin, k) in zip(layers, gys, prepared, shapes):
        a = o.conv2d_dgrad_raw(gy, w, s, (2, 9, 13, cin), k, k, 1, k // 2, impl=code)
        b = o.conv2d_dgrad_raw(gy, w, s, (2, 9, 13, cin), k, k, 1, k // 2, impl=code, prepared_ws=ws)
        assert torch.equal(a, b), (cout, cin, k)
    assert o.dgrad_prepare_batch(layers[:2], torch.device(DEV), impl=o.IMPL_SIMT) == [None, None]


@pytest.mark.timeout(900)
@pytest.mark.parametrize("mode", ["0", "2", "3"])
def test_conv_tcgen05_single_cta_and_forced_pair_modes(mode):
    """conv_tc_kernel runs as single CTAs, as CTA pairs with cta_group::2 MMAs (256-pixel x BN tiles) or as multicast
    pairs (two single-CTA-MMA CTAs sharing the B tile through TMA multicast); the library picks per layer.
    DD_TC_CTA2=0 forces single CTAs, =2 cta_group::2 pairs and =3 multicast pairs for EVERY shape: the whole
    dense-tier suite (13 + 13 shapes, both arms, fused stages) must pass unchanged in each."""
    import subprocess
    import sys
    if os.environ.get("DD_TC_CTA2"):
        pytest.skip("already inside a forced-mode run")
    env = dict(os.environ, DD_TC_CTA2=mode)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_ops.py"), "-m", "gpu", "-q",
                        "-x", "-p", "no:cacheprovider", "-k",
                        "test_conv_tcgen05_arm or test_conv_tcgen05_x3_arm_is_fp32_grade or fused_stage or dgrad_batched"],
                       env=env, cwd=root, capture_output=True, text=True, timeout=800)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
