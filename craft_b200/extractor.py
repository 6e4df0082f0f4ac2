"""Feature / context encoders (core/extractor.py) -- NOT on the named hot path:
they stay stock PyTorch (cuDNN) and exist here only so that `CRAFT(args)` is self-contained and the
reference checkpoints load key for key (conv1, norm1, layer{1,2,3}.{0,1}.{conv1,conv2,norm1,norm2,
norm3,downsample.{0,1}}, conv2)."""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F


def _norm(kind, ch):
    if kind == "batch":
        return nn.BatchNorm2d(ch)
    if kind == "instance":
        return nn.InstanceNorm2d(ch)
    if kind == "group":
        return nn.GroupNorm(num_groups=ch // 8, num_channels=ch)
    if kind == "none":
        return nn.Sequential()
    raise ValueError(kind)


class ResidualBlock(nn.Module):
    def __init__(self, in_planes, planes, norm_fn="group", stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_planes, planes, kernel_size=3, padding=1, stride=stride)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, padding=1)
        self.relu = nn.ReLU(inplace=True)
        self.norm1 = _norm(norm_fn, planes)
        self.norm2 = _norm(norm_fn, planes)
        self.downsample = None
        if stride != 1:
            self.norm3 = _norm(norm_fn, planes)
            self.downsample = nn.Sequential(nn.Conv2d(in_planes, planes, kernel_size=1, stride=stride), self.norm3)

    def forward(self, x):
        y = self.relu(self.norm1(self.conv1(x)))
        y = self.relu(self.norm2(self.conv2(y)))
        if self.downsample is not None:
            x = self.downsample(x)
        return self.relu(x + y)

    def can_conv64(self, fz):
        c1, c2 = self.conv1, self.conv2
        return (fz.own_conv64 and self.downsample is None
                and all(c.in_channels == 64 and c.out_channels == 64 and c.kernel_size == (3, 3) and c.stride == (1, 1)
                        and c.padding == (1, 1) and c.dilation == (1, 1) and c.groups == 1 for c in (c1, c2)))

    def forward_pad(self, x, fz, out_pad=True):
        """The same block on a padded-flat activation (ops.PadAct) with both convolutions on the persistent tcgen05
        kernel (csrc/conv_enc.cuh); the InstanceNorm statistics come out of the convolution's epilogue."""
        ops = fz.ops
        if fz.fold_bn:
            w1, b1 = fz.folded64(self.conv1, self.norm1)
            w2, b2 = fz.folded64(self.conv2, self.norm2)
            y = ops.conv3x3_c64(x, w1, bias=b1, relu=True)
            y2 = ops.conv3x3_c64(y, w2, bias=b2, relu=True)
            return ops.nhwc_affine_pad(y2, None, res=x, relu_out=True, out_pad=out_pad)
        t, ab1 = ops.conv3x3_c64(x, fz.w64(self.conv1), stats_eps=self.norm1.eps)
        y = ops.nhwc_affine_pad(t, ab1, relu_in=True)
        t2, ab2 = ops.conv3x3_c64(y, fz.w64(self.conv2), stats_eps=self.norm2.eps)
        return ops.nhwc_affine_pad(t2, ab2, res=x, relu_in=True, relu_out=True, out_pad=out_pad)

    def forward_fused(self, x, fz):
        """Inference on channels-last activations: cuDNN convs + craft_b200 norm/relu/residual kernels."""
        if fz.fold_bn:      # eval BatchNorm folded into the convolutions, bias + ReLU in cuDNN's epilogue
            y = fz.conv_bn(self.conv1, self.norm1, x, relu=True)
            y2 = fz.conv_bn(self.conv2, self.norm2, y, relu=True)
            if self.downsample is not None:
                x = fz.conv_bn(self.downsample[0], self.norm3, x, relu=False)
            return fz.affine(y2, None, res=x, relu_out=True)
        y = fz.conv(self.conv1, x)
        y = fz.norm_act(self.norm1, y, self.conv1, relu=True)
        y2 = fz.conv(self.conv2, y)
        if fz.one_launch_in:      # instance norm: statistics + transform + residual in one cooperative launch each
            if self.downsample is not None:
                x = fz.instnorm(fz.conv(self.downsample[0], x), self.norm3)
            return fz.instnorm(y2, self.norm2, res=x, relu_in=True, relu_out=True)
        ab2 = fz.scale_shift(self.norm2, y2, self.conv2)
        if self.downsample is not None:
            xd = fz.conv(self.downsample[0], x)
            return fz.affine(y2, ab2, res=xd, rab=fz.scale_shift(self.norm3, xd, self.downsample[0]),
                             relu_in=True, relu_out=True)
        return fz.affine(y2, ab2, res=x, relu_in=True, relu_out=True)


class _Fused:
    """Helpers of the fused inference path: convolutions stay cuDNN on
    channels-last tensors (no NCHW<->NHWC conversion kernels); InstanceNorm / eval BatchNorm + ReLU +
    residual add run as craft_b200 kernels (csrc/encoder.cuh)."""

    def __init__(self, kind, half=False):
        from . import hotpath, ops
        self.kind, self.ops, self.half = kind, ops, half
        self.dtype = torch.float16 if half else torch.float32     # activation / conv operand type (stats stay fp32)
        self.cache = hotpath.PackedWeights()
        # CRAFT_B200_FUSED_IN=1: InstanceNorm2d as ONE cooperative kernel per tensor (ops.instnorm_apply) instead of
        # stats + finalize + affine.  Default off: 28 launches fewer and -0.2 ms of summed kernel time, but the same
        # 4.19-4.21 ms per pair -- with the context encoder on the second stream the encoder phase is bound by SM
        # time, not by its launch chain (profiles/r02_instnorm_one_launch.txt)
        self.one_launch_in = kind == "instance" and os.environ.get("CRAFT_B200_FUSED_IN", "0") == "1"
        # eval-mode BatchNorm (the context encoder) folded into the convolution weights, its shift and the ReLU applied
        # by cuDNN's fused conv-bias-activation: no norm pass at all behind those convolutions.  CRAFT_B200_FOLD_BN=0:
        # the separate scale/shift pass of nhwc_affine
        self.fold_bn = kind == "batch" and os.environ.get("CRAFT_B200_FOLD_BN", "1") != "0"
        # the 64 -> 64 3x3 convolutions of layer1 on the persistent tcgen05 kernel of csrc/conv_enc.cuh (fp16
        # activations in the padded-flat layout) instead of cuDNN; CRAFT_B200_CONV64=0: cuDNN
        # Measured (profiles/r02_conv64.txt): 19.8 us against cuDNN's 23.2 us for two 224x512 images, 26.8 us with the
        # InstanceNorm statistics in its epilogue against 33.4 us for cuDNN + the statistics kernels.  For the
        # context encoder (one image, folded BatchNorm, no statistics) the gain is smaller than the layout-change
        # pass it needs, so only CRAFT_B200_CONV64=2 turns it on there.
        c64 = os.environ.get("CRAFT_B200_CONV64", "1")
        self.own_conv64 = half and ((kind == "instance" and c64 != "0") or (self.fold_bn and c64 == "2"))

    def conv(self, m, x, bias=False):
        """cuDNN convolution WITHOUT its bias: a conv bias in front of a normalisation is either a no-op
        (instance norm subtracts the per-channel mean) or folds into the norm's shift (batch norm), which
        saves one elementwise pass per convolution."""
        w = self.cache.get(("w", id(m)), [m.weight],
                           lambda: m.weight.detach().to(self.dtype).contiguous(memory_format=torch.channels_last))
        b = None
        if bias:
            b = self.cache.get(("b", id(m)), [m.bias], lambda: m.bias.detach().to(self.dtype))
        return F.conv2d(x, w, b, m.stride, m.padding)

    def _folded(self, m, norm, weight_of=None, key="wbn"):
        """(w', b') with w' = w * gamma/sqrt(var+eps) per output channel and b' = beta + (bias - mean) * that scale."""
        def build():
            a = norm.weight.detach().float() * torch.rsqrt(norm.running_var.detach().float() + norm.eps)
            b = norm.bias.detach().float() + (m.bias.detach().float() - norm.running_mean.detach().float()) * a
            w = (weight_of() if weight_of is not None else m.weight.detach().float()) * a.view(-1, 1, 1, 1)
            return w.to(self.dtype).contiguous(memory_format=torch.channels_last), b.to(self.dtype).contiguous()
        return self.cache.get((key, id(m)), [m.weight, m.bias, norm.weight, norm.bias, norm.running_mean, norm.running_var], build)

    def w64(self, m):
        return self.cache.get(("w64", id(m)), [m.weight], lambda: self.ops.pack_conv64_weight(m.weight))

    def folded64(self, m, norm):
        def build():
            a = norm.weight.detach().float() * torch.rsqrt(norm.running_var.detach().float() + norm.eps)
            b = norm.bias.detach().float() + (m.bias.detach().float() - norm.running_mean.detach().float()) * a
            return self.ops.pack_conv64_weight(m.weight, scale=a), b.contiguous()
        return self.cache.get(("wbn64", id(m)), [m.weight, m.bias, norm.weight, norm.bias, norm.running_mean, norm.running_var], build)

    def conv_bn(self, m, norm, x, relu):
        w, b = self._folded(m, norm)
        if relu:
            return torch.cudnn_convolution_relu(x, w, b, m.stride, m.padding, (1, 1), 1)
        return F.conv2d(x, w, b, m.stride, m.padding)

    def scale_shift(self, norm, y, conv):
        if self.kind == "instance":
            return self.ops.instnorm_stats(y.permute(0, 2, 3, 1), eps=norm.eps)

        def build():
            a = norm.weight.detach().float() * torch.rsqrt(norm.running_var.detach().float() + norm.eps)
            b = norm.bias.detach().float() + (conv.bias.detach().float() - norm.running_mean.detach().float()) * a
            return torch.stack([a, b], dim=1)[None].contiguous()
        return self.cache.get(("bn", id(norm)), [norm.weight, norm.bias, norm.running_mean, norm.running_var, conv.bias], build)

    def affine(self, v, ab, res=None, rab=None, relu_in=False, relu_out=False):
        out = torch.empty_like(v)     # keeps the channels-last strides
        self.ops.nhwc_affine(v.permute(0, 2, 3, 1), ab, res.permute(0, 2, 3, 1) if res is not None else None, rab,
                             relu_in, relu_out, out=out.permute(0, 2, 3, 1))
        return out

    def norm_act(self, norm, y, conv, relu=True):
        if self.one_launch_in:
            return self.instnorm(y, norm, relu_in=relu)
        return self.affine(y, self.scale_shift(norm, y, conv), relu_in=relu)

    def instnorm(self, y, norm, res=None, relu_in=False, relu_out=False):
        out = torch.empty_like(y)     # keeps the channels-last strides
        self.ops.instnorm_apply(y.permute(0, 2, 3, 1), res.permute(0, 2, 3, 1) if res is not None else None, None,
                                relu_in, relu_out, eps=norm.eps, out=out.permute(0, 2, 3, 1))
        return out

    def conv1_s2d(self, m, s, norm=None):
        """The 7x7 stride-2 first convolution (core/extractor.py:129) on the space-to-depth input of
        ops.image_s2d: a 4x4 stride-1 convolution over 16 channels, no padding.  Row 2y + ky - 3 of the image is
        row y + a of the s2d grid with parity py, ky = 2a + py + 3, a in [-2, 1]; the same along x."""
        def build32():
            w = m.weight.detach().float()                       # [O, 3, 7, 7]
            O = w.shape[0]
            w4 = torch.zeros((O, 16, 4, 4), dtype=torch.float32, device=w.device)
            for a in range(-2, 2):
                for py in range(2):
                    ky = 2 * a + py + 3
                    if not 0 <= ky <= 6:
                        continue
                    for b in range(-2, 2):
                        for px in range(2):
                            kx = 2 * b + px + 3
                            if 0 <= kx <= 6:
                                c0 = (py * 2 + px) * 3
                                w4[:, c0:c0 + 3, a + 2, b + 2] = w[:, :, ky, kx]
            return w4

        def build():
            return build32().to(self.dtype).contiguous(memory_format=torch.channels_last)
        if norm is not None:        # folded eval BatchNorm + ReLU (see conv_bn)
            w4, b4 = self._folded(m, norm, weight_of=build32, key="wbn_s2d")
            return torch.cudnn_convolution_relu(s.permute(0, 3, 1, 2), w4, b4, (1, 1), (0, 0), (1, 1), 1)
        w4 = self.cache.get(("w_s2d", id(m)), [m.weight], build)
        return F.conv2d(s.permute(0, 3, 1, 2), w4)              # NCHW view of the channels-last buffer


class BasicEncoder(nn.Module):
    def __init__(self, output_dim=128, norm_fn="batch", dropout=0.0):
        super().__init__()
        self.norm_fn = norm_fn
        self.norm1 = _norm(norm_fn, 64) if norm_fn != "group" else nn.GroupNorm(num_groups=8, num_channels=64)
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3)
        self.relu1 = nn.ReLU(inplace=True)
        self.in_planes = 64
        self.layer1 = self._make_layer(64, stride=1)
        self.layer2 = self._make_layer(96, stride=2)
        self.layer3 = self._make_layer(128, stride=2)
        self.conv2 = nn.Conv2d(128, output_dim, kernel_size=1)
        self.dropout = nn.Dropout2d(p=dropout) if dropout > 0 else None
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, (nn.BatchNorm2d, nn.InstanceNorm2d, nn.GroupNorm)):
                if m.weight is not None:
                    nn.init.constant_(m.weight, 1)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def _make_layer(self, dim, stride=1):
        layers = (ResidualBlock(self.in_planes, dim, self.norm_fn, stride=stride),
                  ResidualBlock(dim, dim, self.norm_fn, stride=1))
        self.in_planes = dim
        return nn.Sequential(*layers)

    def _can_fuse(self, x):
        return (x.is_cuda and not self.training and not torch.is_grad_enabled() and x.dtype == torch.float32
                and self.norm_fn in ("instance", "batch") and not torch.is_autocast_enabled()
                and getattr(self, "use_fused", True))

    def _forward_fused(self, x, s2d=None):
        """[N,3,H,W] fp32 -> conv2 output as a channels-last tensor (NCHW shape, NHWC memory), f16 or f32."""
        half = bool(getattr(self, "fused_half", False))
        if getattr(self, "_fz", None) is None or self._fz.kind != self.norm_fn or self._fz.half != half:
            self._fz = _Fused(self.norm_fn, half)
        fz = self._fz
        own = all(blk.can_conv64(fz) for blk in self.layer1)      # layer1 on the persistent tcgen05 convolution
        if fz.fold_bn:
            if s2d is not None:
                x = fz.conv1_s2d(self.conv1, s2d, norm=self.norm1)
            else:
                x = fz.conv_bn(self.conv1, self.norm1, x.to(fz.dtype).contiguous(memory_format=torch.channels_last), relu=True)
            if own:
                x = fz.ops.nhwc_affine_pad(x.permute(0, 2, 3, 1), None)          # layout change only
        else:
            if s2d is not None:
                x = fz.conv1_s2d(self.conv1, s2d)
            else:
                x = fz.conv(self.conv1, x.to(fz.dtype).contiguous(memory_format=torch.channels_last))
            if own:     # norm1 + ReLU, written straight into the padded-flat layout
                x = fz.ops.nhwc_affine_pad(x.permute(0, 2, 3, 1), fz.scale_shift(self.norm1, x, self.conv1), relu_in=True)
            else:
                x = fz.norm_act(self.norm1, x, self.conv1, relu=True)
        layers = (self.layer1, self.layer2, self.layer3)
        if own:
            for i, blk in enumerate(self.layer1):
                x = blk.forward_pad(x, fz, out_pad=(i + 1 < len(self.layer1)))
            x = x.permute(0, 3, 1, 2)               # dense [N,H,W,64] -> the NCHW view of channels-last memory cuDNN takes
            layers = layers[1:]
        for layer in layers:
            for blk in layer:
                x = blk.forward_fused(x, fz)
        return fz.conv(self.conv2, x, bias=True)

    def forward_nhwc(self, x=None, s2d=None):
        """Inference fast path: [N,3,H,W] normalised frames -- or `s2d`, the output of ops.image_s2d on the RAW
        frames (normalisation, space-to-depth and padding of the first convolution in one kernel) -- to
        [N, H/8, W/8, C] contiguous (channels-last), in the encoder's activation type.  Callers check _can_fuse."""
        y = self._forward_fused(x, s2d).permute(0, 2, 3, 1)
        return y if y.is_contiguous() else y.contiguous()

    def fused_dtype(self):
        return torch.float16 if bool(getattr(self, "fused_half", False)) else torch.float32

    def forward(self, x):
        is_list = isinstance(x, (tuple, list))
        if is_list:
            batch_dim = x[0].shape[0]
            x = torch.cat(x, dim=0)
        if self._can_fuse(x):
            x = self._forward_fused(x).float().contiguous()      # NCHW fp32, the reference's output format
            if is_list:
                x = torch.split(x, [batch_dim, batch_dim], dim=0)
            return x
        x = self.relu1(self.norm1(self.conv1(x)))
        x = self.conv2(self.layer3(self.layer2(self.layer1(x))))
        if self.training and self.dropout is not None:
            x = self.dropout(x)
        if is_list:
            x = torch.split(x, [batch_dim, batch_dim], dim=0)
        return x
