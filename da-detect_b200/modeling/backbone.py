"""ResNet-50/101 C4 trunk and res5 head over the fused NHWC conv kernels.

Module/parameter names mirror the reference so its checkpoints load unchanged
(maskrcnn_benchmark/modeling/backbone/resnet.py:80-336, backbone.py:68-73):
``backbone.body.stem.conv1.weight``, ``backbone.body.layer{1,2,3}.{i}.conv{1,2,3}.weight``,
``....bn{1,2,3}.{weight,bias,running_mean,running_var}``, ``....downsample.{0.weight,1.*}``.
Each Conv2d + FrozenBatchNorm2d (+ residual add) (+ ReLU) group of the reference is ONE kernel here.
"""
from collections import OrderedDict

import torch
from torch import nn

from .. import ops

STAGE_BLOCKS = {"R-50-C4": (3, 4, 6), "R-101-C4": (3, 4, 23),
                "R-50-FPN": (3, 4, 6, 3), "R-101-FPN": (3, 4, 23, 3), "R-152-FPN": (3, 8, 36, 3)}      # resnet.py:41-77


class FrozenBatchNorm2d(nn.Module):
    """Buffers only (layers/batch_norm.py:6-24); ``affine()`` folds them into the per-channel
    (scale, bias) pair the conv epilogue applies.  Cached until a buffer is modified."""

    def __init__(self, n):
        super().__init__()
        self.register_buffer("weight", torch.ones(n))
        self.register_buffer("bias", torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))
        self._cache = None

    def affine(self):
        key = (self.weight._version, self.bias._version, self.running_mean._version, self.running_var._version,
               self.weight.data_ptr())
        if self._cache is None or self._cache[0] != key:
            with torch.no_grad():
                scale = self.weight * self.running_var.rsqrt()
                bias = self.bias - self.running_mean * scale
            self._cache = (key, scale.contiguous(), bias.contiguous())
        return self._cache[1], self._cache[2]


class Conv2dParams(nn.Module):
    """Holds a conv weight (reference-shaped, stored channels_last = physical OHWI) and optional bias."""

    def __init__(self, cin, cout, k, stride=1, padding=0, bias=False):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k).contiguous(memory_format=torch.channels_last))
        self.bias = nn.Parameter(torch.zeros(cout)) if bias else None
        self.stride, self.padding = stride, padding
        nn.init.kaiming_uniform_(self.weight, a=1)

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if self.weight.dim() == 4 and not self.weight.is_contiguous(memory_format=torch.channels_last):
            self.weight.data = self.weight.data.contiguous(memory_format=torch.channels_last)
        return out


class Bottleneck(nn.Module):
    """BottleneckWithFixedBatchNorm with STRIDE_IN_1X1 (resnet.py:227-314)."""

    def __init__(self, cin, mid, cout, stride):
        super().__init__()
        self.downsample = None
        if cin != cout:
            self.downsample = nn.Sequential(Conv2dParams(cin, cout, 1, stride), FrozenBatchNorm2d(cout))
        self.conv1 = Conv2dParams(cin, mid, 1, stride)
        self.bn1 = FrozenBatchNorm2d(mid)
        self.conv2 = Conv2dParams(mid, mid, 3, 1, 1)
        self.bn2 = FrozenBatchNorm2d(mid)
        self.conv3 = Conv2dParams(mid, cout, 1)
        self.bn3 = FrozenBatchNorm2d(cout)
        self.stride = stride

    def forward(self, x, stride_override=None):
        stride = self.stride if stride_override is None else stride_override
        s1, b1 = self.bn1.affine()
        s2, b2 = self.bn2.affine()
        s3, b3 = self.bn3.affine()
        out = ops.conv_bn_act(x, self.conv1.weight, s1, b1, stride=stride, relu=True)
        out = ops.conv_bn_act(out, self.conv2.weight, s2, b2, pad=1, relu=True)
        identity = x
        if self.downsample is not None:
            sd, bd = self.downsample[1].affine()
            identity = ops.conv_bn_act(x, self.downsample[0].weight, sd, bd, stride=stride)
        return ops.conv_bn_act(out, self.conv3.weight, s3, b3, residual=identity, relu=True)


class Stage(nn.Sequential):
    """make_stage (resnet.py:197-224): the blocks keep their reference names (``layerN.i.*``); the forward of
    the whole stage is one fused autograd node (ops.bottleneck_stage) unless `fused` is switched off."""

    fused = True

    def forward(self, x, first_stride_override=None, input_is_relu=False, grad_premasked=False, pool_output=False):
        if not self.fused:
            for i, blk in enumerate(self):
                x = blk(x, stride_override=first_stride_override if i == 0 else None)
            return ops.avgpool_hw(x) if pool_output else x
        blocks, strides = [], []
        for i, blk in enumerate(self):
            d = {}
            for j, (conv, bn) in enumerate(((blk.conv1, blk.bn1), (blk.conv2, blk.bn2), (blk.conv3, blk.bn3)), 1):
                sc, bi = bn.affine()
                d["w%d" % j], d["s%d" % j], d["b%d" % j] = conv.weight, sc, bi
            if blk.downsample is not None:
                sc, bi = blk.downsample[1].affine()
                d["wd"], d["sd"], d["bd"] = blk.downsample[0].weight, sc, bi
            blocks.append(d)
            strides.append(first_stride_override if (i == 0 and first_stride_override is not None) else blk.stride)
        return ops.bottleneck_stage(x, blocks, strides, input_is_relu, grad_premasked, pool_output)


def make_stage(cin, mid, cout, blocks, first_stride):
    layers = []
    for i in range(blocks):
        layers.append(Bottleneck(cin, mid, cout, first_stride if i == 0 else 1))
        cin = cout
    return Stage(*layers)


class Stem(nn.Module):
    """7x7/2 conv + FrozenBN + ReLU (one kernel) + 3x3/2 max-pool (resnet.py:317-336)."""

    def __init__(self, cout=64):
        super().__init__()
        self.conv1 = Conv2dParams(3, cout, 7, 2, 3)
        self.bn1 = FrozenBatchNorm2d(cout)

    def forward(self, x_nchw):
        """x_nchw: the reference's input layout [N,3,H,W]; everything downstream is NHWC."""
        s, b = self.bn1.affine()
        if ops.stem_tc_supported(x_nchw, self.conv1.weight):
            x = ops.stem_conv7x7s2(x_nchw, self.conv1.weight, s, b, relu=True)
        else:
            x = ops.conv_bn_act(ops.nchw_to_nhwc(x_nchw), self.conv1.weight, s, b, stride=2, pad=3, relu=True)
        return ops.maxpool3x3s2(x)


class ResNetC4(nn.Module):
    """ResNet.forward (resnet.py:138-145): NCHW image in, NHWC maps out.  *-C4 bodies return [res4]; *-FPN bodies
    return [res2, res3, res4, res5] (every stage has return_features, resnet.py:60-66)."""

    def __init__(self, cfg):
        super().__init__()
        body = cfg.MODEL.BACKBONE.CONV_BODY
        if body not in STAGE_BLOCKS:
            raise NotImplementedError("backbone {} is outside the accelerated path (SURVEY §8f)".format(body))
        R = cfg.MODEL.RESNETS
        if not R.STRIDE_IN_1X1 or R.NUM_GROUPS != 1 or R.TRANS_FUNC != "BottleneckWithFixedBatchNorm":
            raise NotImplementedError("only the MSRA ResNet layout (STRIDE_IN_1X1, FrozenBN, 1 group) is built")
        self.stem = Stem(R.STEM_OUT_CHANNELS)
        cin = R.STEM_OUT_CHANNELS
        self.stages = []
        for li, nb in enumerate(STAGE_BLOCKS[body]):
            mid = R.NUM_GROUPS * R.WIDTH_PER_GROUP * 2 ** li
            cout = R.RES2_OUT_CHANNELS * 2 ** li
            name = "layer{}".format(li + 1)
            self.add_module(name, make_stage(cin, mid, cout, nb, 1 if li == 0 else 2))
            self.stages.append(name)
            cin = cout
        self.out_channels = cin
        self.return_all = body.endswith("-FPN")
        self._freeze(cfg.MODEL.BACKBONE.FREEZE_CONV_BODY_AT)

    def _freeze(self, freeze_at):
        """_freeze_backbone (resnet.py:127-136)."""
        for idx in range(max(freeze_at, 0)):
            m = self.stem if idx == 0 else getattr(self, "layer{}".format(idx))
            for p in m.parameters():
                p.requires_grad = False

    def forward(self, x):
        x = self.stem(x)
        if self.return_all:
            # every stage output also feeds an FPN lateral conv, so no ReLU mask may be deferred across a boundary
            outs = []
            for name in self.stages:
                x = getattr(self, name)(x)
                outs.append(x)
            return outs
        last = len(self.stages) - 1
        for i, name in enumerate(self.stages):
            # every stage output is a ReLU output consumed only by the next stage, so the mask (x > 0) of the
            # gradient crossing a stage boundary is applied by the consumer's dgrad epilogue
            x = ops.grad_milestone(x, "in:" + name)          # backward: this stage's weight gradients are complete
            x = getattr(self, name)(x, input_is_relu=i > 0, grad_premasked=i < last)
        return [ops.grad_milestone(x, "out:body")]           # backward: every head's weight gradients are complete


class ResNetHead(nn.Module):
    """res5 head (resnet.py:148-194 as built at roi_box_feature_extractors.py:27-37).  When its input
    is the even-bin ROIAlign output, the first block's stride-2 1x1 convs run with stride 1 on the
    7x7 grid — identical values, a quarter of the pooling work."""

    def __init__(self, cfg):
        super().__init__()
        R = cfg.MODEL.RESNETS
        cout = R.RES2_OUT_CHANNELS * 8
        mid = R.NUM_GROUPS * R.WIDTH_PER_GROUP * 8
        self.layer4 = make_stage(cout // 2, mid, cout, 3, 2)
        self.out_channels = cout

    def forward(self, x, input_is_even_bins, pooled=False):
        """pooled: return the 7x7 average [K,C] (what the predictors and the DA instance head consume) — its
        backward is then fused with the ReLU mask of the res5 output."""
        return self.layer4(x, first_stride_override=1 if input_is_even_bins else None, pool_output=pooled)


class FPN(nn.Module):
    """Feature pyramid (modeling/backbone/fpn.py:8-85) over [C2..C5]: 1x1 laterals ``fpn_inner{i}``, 3x3 outputs
    ``fpn_layer{i}`` (conv_with_kaiming_uniform: bias, no norm, no ReLU — make_layers.py:97-122), nearest 2x
    top-down path, LastLevelMaxPool for P6.  The lateral add rides in the lateral conv's epilogue (the up-sampled
    coarser map is its residual input)."""

    def __init__(self, in_channels_list, out_channels):
        super().__init__()
        self.inner_blocks, self.layer_blocks = [], []
        for idx, cin in enumerate(in_channels_list, 1):
            inner, layer = "fpn_inner{}".format(idx), "fpn_layer{}".format(idx)
            self.add_module(inner, Conv2dParams(cin, out_channels, 1, bias=True))
            self.add_module(layer, Conv2dParams(out_channels, out_channels, 3, 1, 1, bias=True))
            self.inner_blocks.append(inner)
            self.layer_blocks.append(layer)

    def forward(self, feats):
        def conv(name, x, residual=None, pad=0):
            m = getattr(self, name)
            return ops.conv_bn_act(x, m.weight, None, m.bias, residual=residual, pad=pad)

        last_inner = conv(self.inner_blocks[-1], feats[-1])
        results = [conv(self.layer_blocks[-1], last_inner, pad=1)]
        for feat, inner, layer in zip(feats[:-1][::-1], self.inner_blocks[:-1][::-1], self.layer_blocks[:-1][::-1]):
            n, h, w, _ = feat.shape
            if (h, w) != (2 * last_inner.shape[1], 2 * last_inner.shape[2]):
                raise RuntimeError("FPN: a {}x{} map cannot take the 2x up-sampled {}x{} map — pad the images to a "
                                   "multiple of 32 (DATALOADER.SIZE_DIVISIBILITY)".format(
                                       h, w, last_inner.shape[1], last_inner.shape[2]))
            last_inner = conv(inner, feat, residual=ops.upsample2x(last_inner))     # fpn.py:62-67
            results.insert(0, conv(layer, last_inner, pad=1))
        results.append(ops.subsample2(results[-1]))                                  # LastLevelMaxPool
        return results


class _BodyFPN(nn.Sequential):
    def forward(self, x):
        return self.fpn(self.body(x))


def build_backbone(cfg):
    """build_backbone (backbone.py:11-18,21-43,68-73): nn.Sequential(OrderedDict([('body', ResNet)(, ('fpn', FPN))]))."""
    body = ResNetC4(cfg)
    if cfg.MODEL.BACKBONE.CONV_BODY.endswith("-FPN"):
        if cfg.MODEL.FPN.USE_GN or cfg.MODEL.FPN.USE_RELU:
            raise NotImplementedError("FPN.USE_GN / FPN.USE_RELU are not used by any DA config")
        c2 = cfg.MODEL.RESNETS.RES2_OUT_CHANNELS
        fpn = FPN([c2, c2 * 2, c2 * 4, c2 * 8], cfg.MODEL.BACKBONE.OUT_CHANNELS)
        return _BodyFPN(OrderedDict([("body", body), ("fpn", fpn)]))
    return nn.Sequential(OrderedDict([("body", body)]))
