"""Feature-Masking operator — drop-in for ref backbones/fm/fmoperator.py (FMCnn :84-311,
FMNone :314-325).

    x = cat(yf, yo) -> same_conv -> res_block (N bottlenecks) = z
    out = arith(yf, act(z)) [+ f_out] + yf                           (ref :285-310)

The conv stack stays on cuDNN; the whole elementwise tail — activation,
arithmetic strategy, peer add, skip — and its backward are ONE fused CUDA kernel each
(ops.fm_gate -> msml_fm_gate_fwd / msml_fm_gate_bwd, csrc/fm_gate.cu).  Parameter names
(same_conv, res_block.{i}.conv1 ... prelu3, conv1, conv2, conv_m) match the reference.
"""
import numpy as np
import torch
from torch import nn

from ... import ops
from .._blocks import conv1x1, conv3x3

__all__ = ["FMCnn", "FMNone"]


class resblock_bottle(nn.Module):
    """1x1 -> 3x3 -> 1x1 bottleneck with BN/PReLU (ref :35-68); bottle width C/2 if C <= 128 else 128."""

    def __init__(self, in_channels, out_channels, bottle_channels=128):
        super().__init__()
        if in_channels <= 128:
            bottle_channels = in_channels // 2
        self.conv1 = conv1x1(in_channels, bottle_channels)
        self.bn1 = nn.BatchNorm2d(bottle_channels, eps=1e-05)
        self.prelu1 = nn.PReLU(bottle_channels)
        self.conv2 = conv3x3(bottle_channels, bottle_channels)
        self.bn2 = nn.BatchNorm2d(bottle_channels, eps=1e-05)
        self.prelu2 = nn.PReLU(bottle_channels)
        self.conv3 = conv1x1(bottle_channels, out_channels)
        self.bn3 = nn.BatchNorm2d(out_channels, eps=1e-05)
        self.prelu3 = nn.PReLU(out_channels)

    def forward(self, x):
        y = ops.bn_act(ops.conv2d(x, self.conv1), self.bn1, self.prelu1)
        y = ops.bn_act(ops.conv2d(y, self.conv2), self.bn2, self.prelu2)
        return ops.bn_act(ops.conv2d(y, self.conv3), self.bn3, self.prelu3, x)     # prelu3(bn3(.) + identity)


def _conv_bn_prelu_x2(c):
    return nn.Sequential(nn.Conv2d(c, c, 3, 1, 1), nn.BatchNorm2d(c, eps=1e-05), nn.PReLU(c),
                         nn.Conv2d(c, c, 3, 1, 1), nn.BatchNorm2d(c, eps=1e-05), nn.PReLU(c))


def _run_conv_bn_prelu(seq, x):
    """An nn.Sequential of (Conv2d, BatchNorm2d, PReLU) triples (or the empty one of use_conv False, ref :138-158) through the
    fused BN + PReLU kernels; module and parameter names stay those of the reference's Sequential."""
    mods = list(seq)
    for i in range(0, len(mods), 3):
        x = ops.bn_act(ops.conv2d(x, mods[i]), mods[i + 1], mods[i + 2])
    return x


class _Invert(nn.Module):
    def forward(self, x):
        return 1 - x


class FMCnn(nn.Module):
    _acts = {"tanh": torch.tanh, "sigmoid": torch.sigmoid}
    _ariths = ("add", "sub", "div", "mul")

    def __init__(self, height, width, channel_f, kernel_size=3, resblocks=2, activation='tanh',
                 arith_strategy='add', peer_params: dict = None):
        super().__init__()
        self.height, self.width, self.channel_f = height, width, channel_f
        self.same_conv = (conv1x1 if kernel_size == 1 else conv3x3)(18 + channel_f, channel_f)
        self.res_block = nn.Sequential(*[resblock_bottle(channel_f, channel_f) for _ in range(resblocks)])
        self.activation = activation
        self.mask_norm = self._acts[activation]          # KeyError on a bad name, as the reference
        if arith_strategy not in self._ariths:
            raise KeyError(arith_strategy)
        self.arith_strategy = arith_strategy

        # peer-guided distillation branch (ref :129-166); inactive when use_ori is False
        self.use_ori = peer_params.get('use_ori')
        en_conv = peer_params.get('use_conv')
        self.conv1 = _conv_bn_prelu_x2(channel_f) if self.use_ori and en_conv else nn.Sequential()
        self.conv2 = _conv_bn_prelu_x2(channel_f) if self.use_ori and en_conv else nn.Sequential()
        mask_trans = peer_params.get('mask_trans')
        if not self.use_ori:
            self.conv_m = nn.Sequential()
        elif mask_trans == 'conv':
            self.conv_m = nn.Sequential(nn.Conv2d(channel_f, channel_f, 3, 1, 1), nn.BatchNorm2d(channel_f, eps=1e-05))
        elif mask_trans == 'invert':
            self.conv_m = _Invert()
        else:
            raise ValueError('mask_trans type error')
        self.en_save = False

    # -- optional feature dumps used by the reference's eval scripts (ref :177-200) ----------
    def _save_intermediate_features(self, feat_type, feat_tensor):
        if not self.en_save:
            return
        arr = feat_tensor.detach().float().flatten().cpu().numpy()
        if feat_type == 'contaminated':
            self.contaminated_feat = arr
        elif feat_type == 'mask':
            self.mask_feat = arr
        elif feat_type == 'purified':
            self.purified_feat = arr
        else:
            raise ValueError('Intermediate feature type error!')

    def plot_intermediate_features(self, gt_occ_msk, save_folder="."):
        """Scatter plots of Y_f vs M and Y_f vs Z_f coloured by the ground-truth occlusion mask
        (ref :202-275).  Needs matplotlib and PIL, which the training path never imports."""
        import os
        import matplotlib.pyplot as plt
        from PIL import Image
        occ = (gt_occ_msk.numpy().astype(np.uint8) * 255)
        small = np.stack([np.array(Image.fromarray(o, mode='L').resize((self.width, self.height))) // 255 for o in occ])
        small = np.repeat(small[:, None], self.channel_f, axis=1).reshape(-1)
        assert small.size == self.mask_feat.size
        colors = np.where(small == 1, 0.7, 0.3)
        for tag, ys, label in (("cm", self.mask_feat, "Mask Generated by FM Operators"),
                               ("cp", self.purified_feat, "Face Feature Purified by FM Operators")):
            name = 'fm_%s_%d_%s.jpg' % (tag, self.height, self.arith_strategy)
            plt.figure(dpi=300)
            plt.title(name)
            plt.xlabel('Contaminated Face Feature')
            plt.ylabel(label)
            plt.scatter(x=self.contaminated_feat, y=ys, s=1, c=colors, alpha=0.4)
            if tag == "cp":
                lo, hi = self.contaminated_feat.min(), self.contaminated_feat.max()
                plt.plot([lo, hi], [lo, hi], 'r--', linewidth=1)
            plt.savefig(os.path.join(save_folder, name))
            plt.clf()

    def forward(self, yf, yo, yt=None):
        """yf (B,C,H,W) face features, yo (B,18,H,W) occlusion maps, yt peer features (train only)
        -> (Z_f with the shape of yf, l2 distillation loss or None)."""
        x, pad, yf = ops.fm_cat(yf, yo)        # C+18 channels zero-padded to a multiple of 8 (K-C); yf for the tail
        z = self.res_block(ops.conv2d_padded_in(x, self.same_conv, pad))
        f_out, l2 = None, None
        if self.en_save:
            self._save_intermediate_features('contaminated', yf)
            self._save_intermediate_features('mask', self.mask_norm(z))
        if self.use_ori:
            # peer-guided branch (ref :293-302): both masked products in one pass (K-P); with mask_trans 'invert' the kernel
            # forms 1 - act(z) itself, with 'conv' the gate has to exist for conv_m
            if isinstance(self.conv_m, _Invert):
                prod = ops.fm_peer_mul(z, yf, yt, "invert", self.activation)
            else:
                m_bar = ops.bn_act(ops.conv2d(self.mask_norm(z), self.conv_m[0]), self.conv_m[1])
                prod = ops.fm_peer_mul(m_bar, yf, yt)
            pf, pt = prod if yt is not None else (prod, None)
            f_out = _run_conv_bn_prelu(self.conv1, pf)
            if yt is not None:
                l2 = ops.mse_loss(_run_conv_bn_prelu(self.conv2, pt), f_out)
        if self.en_save:
            self._save_intermediate_features('purified', ops.fm_gate(yf, z, self.activation, self.arith_strategy) - yf)
        out = ops.fm_gate(yf, z, self.activation, self.arith_strategy, f_out)   # fused tail (K-A)
        return out, l2


class FMNone(nn.Module):
    def forward(self, yf, yo, yt=None):
        return yf, None
