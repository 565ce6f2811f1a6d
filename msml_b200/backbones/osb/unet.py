"""Occlusion-Segmentation Branch — drop-in for ref backbones/osb/unet.py (Unet :94-240, unet() :243-279).

iResNet-style encoder, five Global-Conv modules, five transposed-conv decoder stages.  Emits four
DETACHED 18-channel maps (7^2, 14^2, 28^2, 56^2) for the FM operators plus the final 2-channel
segmentation.  The DAP head (PixelShuffle(3)+AvgPool(3), ref :158-161,223) is the fused CUDA
kernel ops.dap (csrc/dap.cu): a mean over 9-channel groups that never builds the (B,2,336,336)
intermediate; ``forward_with_mask`` also returns the argmax occlusion mask from the same kernel.
"""
import torch
from torch import nn

from ... import ops
from .._blocks import make_stage

__all__ = ["unet", "Unet"]


class _GlobalConvModule(nn.Module):
    """Large-kernel separable context: (k x 1 -> 1 x k) + (1 x k -> k x 1)   (ref :16-38)."""

    def __init__(self, in_dim, out_dim, kernel_size):
        super().__init__()
        kh, kw = kernel_size
        ph, pw = (kh - 1) // 2, (kw - 1) // 2
        self.conv_l1 = nn.Conv2d(in_dim, out_dim, (kh, 1), padding=(ph, 0))
        self.conv_l2 = nn.Conv2d(out_dim, out_dim, (1, kw), padding=(0, pw))
        self.conv_r1 = nn.Conv2d(in_dim, out_dim, (1, kw), padding=(0, pw))
        self.conv_r2 = nn.Conv2d(out_dim, out_dim, (kh, 1), padding=(ph, 0))

    def forward(self, x):
        return (ops.conv2d(ops.conv2d(x, self.conv_l1), self.conv_l2) +
                ops.conv2d(ops.conv2d(x, self.conv_r1), self.conv_r2))

    def forward_padded(self, x, po):
        """Same module with ``po`` zero output channels appended (they stay exactly zero): 18-channel maps become
        24-channel ones so that cuDNN's tensor-core kernels apply (C % 8 == 0)."""
        F = torch.nn.functional

        def conv(t, m, pad_in):
            w, b = ops.padded_params(m, pad_in, po)
            return F.conv2d(t, w, b, m.stride, m.padding)
        return conv(conv(x, self.conv_l1, 0), self.conv_l2, po) + conv(conv(x, self.conv_r1, 0), self.conv_r2, po)


class _DAP(nn.Module):
    def __init__(self, k):
        super().__init__()
        self.k = k

    def forward(self, x):
        return ops.dap(x, self.k)


class Unet(nn.Module):
    pad_channels = True      # 16-bit CUDA path: run the decoder on channel counts padded to multiples of 8

    def __init__(self, block, layers, groups=1, num_classes=2, kernel_size=7, dap_k=3, gray=True, input_size=128):
        super().__init__()
        del block, groups
        self.conv1 = nn.Conv2d(1 if gray else 3, 64, kernel_size=3, stride=2, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(64, eps=1e-05)
        self.prelu = nn.PReLU(64)
        self.layer1 = make_stage(64, 64, layers[0], 2, feeds_bn=True)
        self.layer2 = make_stage(64, 128, layers[1], 2, feeds_bn=True)
        self.layer3 = make_stage(128, 256, layers[2], 2, feeds_bn=True)
        self.layer4 = make_stage(256, 512, layers[3], 2, feeds_bn=True)
        self.bn2 = nn.BatchNorm2d(512, eps=1e-05)

        seg = num_classes * dap_k ** 2
        ks = (kernel_size, kernel_size)
        self.gcm1 = _GlobalConvModule(512, num_classes * 4, ks)
        self.gcm2 = _GlobalConvModule(256, seg, ks)
        self.gcm3 = _GlobalConvModule(128, seg, ks)
        self.gcm4 = _GlobalConvModule(64, seg, ks)
        self.gcm5 = _GlobalConvModule(64, seg, ks)
        if input_size == 128:
            self.deconv1 = nn.ConvTranspose2d(num_classes * 4, seg, kernel_size=4, stride=2, padding=1, bias=False)
        elif input_size == 112:
            self.deconv1 = nn.ConvTranspose2d(num_classes * 4, seg, kernel_size=3, stride=2, padding=1, bias=False)
        else:
            raise ValueError('Error in input_size.')
        for i in (2, 3, 4, 5):
            setattr(self, 'deconv%d' % i, nn.ConvTranspose2d(2 * seg, seg, kernel_size=4, stride=2, padding=1, bias=False))
        self.DAP = _DAP(dap_k)

    def _decode_padded(self, x4, x3, x2, x1, x0):
        """The decoder on channel counts padded to multiples of 8 (bf16 autocast path).  Padding channels carry zero
        weights and zero biases, so the first ``seg`` channels are bit-identical in exact arithmetic; the results are
        sliced back to the reference's shapes."""
        F = torch.nn.functional
        seg = self.deconv2.weight.shape[1]
        po = (-seg) % 8
        w, b = ops.padded_params(self.deconv1, (-self.deconv1.weight.shape[0]) % 8, po, transposed=True)
        g1 = self.gcm1.forward_padded(x4, (-self.deconv1.weight.shape[0]) % 8)
        d = self.deconv1
        s = F.conv_transpose2d(g1, w, b, d.stride, d.padding, d.output_padding)
        outs = [s]
        for i, xi in zip((2, 3, 4, 5), (x3, x2, x1, x0)):
            d = getattr(self, 'deconv%d' % i)
            g = getattr(self, 'gcm%d' % i).forward_padded(xi, po)
            wd = ops._weight_of(d)                                       # (2*seg, seg, 4, 4): rows [prev seg | gcm]
            z = wd.new_zeros((po,) + tuple(wd.shape[1:])) if po else None
            wp = torch.cat((wd[:seg], z, wd[seg:], z), 0) if po else wd
            wp = F.pad(wp, (0, 0, 0, 0, 0, po)) if po else wp
            s = F.conv_transpose2d(torch.cat((s, g), 1), wp, None, d.stride, d.padding, d.output_padding)
            outs.append(s)
        return [o[:, :seg] for o in outs]

    def _decode(self, x):
        x0 = ops.bn_act(ops.conv2d_padded_in(x, self.conv1, x.shape[1] - self.conv1.in_channels), self.bn1, self.prelu,
                       emit_next_stats=True)      # layer1's first bn1 reads this tensor next
        x1 = self.layer1(x0)
        x2 = self.layer2(x1)
        x3 = self.layer3(x2)
        x4 = ops.bn_act(self.layer4(x3), self.bn2)
        if self.pad_channels and x4.is_cuda and x4.dtype != torch.float32:
            seg0, seg1, seg2, seg3, seg5_ = self._decode_padded(x4, x3, x2, x1, x0)
            return [seg0.detach(), seg1.detach(), seg2.detach(), seg3.detach()], seg5_
        seg0 = ops.conv_transpose2d(self.gcm1(x4), self.deconv1)
        seg1 = ops.conv_transpose2d(torch.cat((seg0, self.gcm2(x3)), 1), self.deconv2)
        seg2 = ops.conv_transpose2d(torch.cat((seg1, self.gcm3(x2)), 1), self.deconv3)
        seg3 = ops.conv_transpose2d(torch.cat((seg2, self.gcm4(x1)), 1), self.deconv4)
        seg5_ = ops.conv_transpose2d(torch.cat((seg3, self.gcm5(x0)), 1), self.deconv5)
        return [seg0.detach(), seg1.detach(), seg2.detach(), seg3.detach()], seg5_   # detach link, ref :227-230

    def forward(self, x):
        feats, seg5_ = self._decode(x)
        return feats + [self.DAP(seg5_)]

    def forward_with_mask(self, x):
        """-> ([seg0..seg3, seg5], argmax mask (B,H,W) int64): DAP and the 2-way argmax in one kernel."""
        feats, seg5_ = self._decode(x)
        seg5, mask = ops.dap_with_mask(seg5_.float(), self.DAP.k)
        return feats + [seg5], mask


_LAYERS = {'r18': [2, 2, 2, 2], 'r34': [3, 4, 6, 3], 'r50': [3, 4, 14, 3], 'r100': [3, 13, 30, 3], 'r200': [6, 26, 60, 6]}


def unet(pre_trained=False, backbone='r18', gray=True, input_size=128, **kwargs):
    if pre_trained:
        print('No pretrained model for mskfuse29_light_y_seg')
    for tag, layers in _LAYERS.items():
        if tag in backbone:
            return Unet(None, layers, num_classes=2, gray=gray, input_size=input_size, **kwargs)
    raise ValueError('Error backbone type in OSB.')
