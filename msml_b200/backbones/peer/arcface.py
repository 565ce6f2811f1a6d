"""Peer (teacher) network — drop-in for ref backbones/peer/arcface.py (IResNet :71-194, arcface18/34/50/100 and
cosface50_casia :215-237): a frozen vanilla iResNet that returns its embedding and the four DETACHED stage outputs
``[ft0 (B,64,56,56), ft1 (B,128,28,28), ft2 (B,256,14,14), ft3 (B,512,7,7)]`` which the FM operators distil from
(ref backbones/fm/fmoperator.py:293-302).

Same constructors and module / parameter names as the reference, so its checkpoints (``r18-backbone.pth`` …, the
insightface iResNet weights the reference's README points to) load unchanged.  As in the reference, ``pretrained=True`` is
the default and raises FileNotFoundError when the file is missing; ``pretrained=False`` gives a randomly initialised
peer (ref :119-129 initialisation), which is what the parity tests use — the reference ships no weights.
The BatchNorm / PReLU / residual chains run the fused NHWC kernels (ops.bn_act, eval mode), convolutions cuDNN.
"""
import os

import torch
from torch import nn

from ... import ops
from .._blocks import IBasicBlock, make_stage

__all__ = ['IResNet', 'arcface18', 'arcface34', 'arcface50', 'arcface100', 'cosface50_casia']

model_dir = {
    'arcface18': './backbones/pretrained/r18-backbone.pth',
    'arcface34': './backbones/pretrained/r34-backbone.pth',
    'arcface50': './backbones/pretrained/r50-backbone.pth',
    'arcface100': './backbones/pretrained/r100-backbone.pth',
    'cosface50_casia': './backbones/pretrained/cos50_no_occ_2.pth',
}


class IResNet(nn.Module):
    fc_scale = 7 * 7

    def __init__(self, block, layers, dim_feature=512, dropout=0, zero_init_residual=False, groups=1, width_per_group=64,
                 replace_stride_with_dilation=None, fp16=False):
        super().__init__()
        del block
        if groups != 1 or width_per_group != 64:
            raise ValueError('BasicBlock only supports groups=1 and base_width=64')
        if replace_stride_with_dilation is not None and any(replace_stride_with_dilation):
            raise NotImplementedError("Dilation > 1 not supported in BasicBlock")
        self.fp16 = fp16
        self.conv1 = nn.Conv2d(3, 64, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(64, eps=1e-05)
        self.prelu = nn.PReLU(64)
        self.layer1 = make_stage(64, 64, layers[0], 2, feeds_bn=True)
        self.layer2 = make_stage(64, 128, layers[1], 2, feeds_bn=True)
        self.layer3 = make_stage(128, 256, layers[2], 2, feeds_bn=True)
        self.layer4 = make_stage(256, 512, layers[3], 2, feeds_bn=True)
        self.bn2 = nn.BatchNorm2d(512, eps=1e-05)
        self.dropout = nn.Dropout(p=dropout, inplace=True)
        self.fc = nn.Linear(512 * self.fc_scale, dim_feature)
        self.features = nn.BatchNorm1d(dim_feature, eps=1e-05)
        nn.init.constant_(self.features.weight, 1.0)
        self.features.weight.requires_grad = False
        for m in self.modules():     # ref :119-129
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, 0, 0.1)
            elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        if zero_init_residual:
            for m in self.modules():
                if isinstance(m, IBasicBlock):
                    nn.init.constant_(m.bn2.weight, 0)

    def forward(self, x):
        """img (B, 3, 112, 112) -> (feature (B, dim_feature), [ft0, ft1, ft2, ft3] detached)   (ref :159-194)"""
        inter = []
        if x.is_cuda:
            x = x.contiguous(memory_format=torch.channels_last)
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=bool(self.fp16) and x.is_cuda):
            x = ops.bn_act(ops.conv2d(x, self.conv1), self.bn1, self.prelu, emit_next_stats=True)
            for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
                x = layer(x)
                inter.append(x.detach())
            x = ops.bn_act(x, self.bn2)
            x = self.dropout(torch.flatten(x, 1))
        x = self.features(self.fc(x.float() if self.fp16 else x))
        return x, inter


def _iresnet_v(arch, block, layers, pretrained, progress, **kwargs):
    del progress
    model = IResNet(block, layers, **kwargs)
    if pretrained:
        if os.path.isfile(model_dir[arch]):
            model.load_state_dict(torch.load(model_dir[arch], map_location=torch.device('cpu')))
        else:
            raise FileNotFoundError('Make sure the file {' + model_dir[arch] + '} exists!')
    return model.eval()


def arcface18(pretrained=True, progress=True, **kwargs):
    return _iresnet_v('arcface18', IBasicBlock, [2, 2, 2, 2], pretrained, progress, **kwargs)


def arcface34(pretrained=True, progress=True, **kwargs):
    return _iresnet_v('arcface34', IBasicBlock, [3, 4, 6, 3], pretrained, progress, **kwargs)


def arcface50(pretrained=True, progress=True, **kwargs):
    return _iresnet_v('arcface50', IBasicBlock, [3, 4, 14, 3], pretrained, progress, **kwargs)


def arcface100(pretrained=True, progress=True, **kwargs):
    return _iresnet_v('arcface100', IBasicBlock, [3, 13, 30, 3], pretrained, progress, **kwargs)


def cosface50_casia(pretrained=True, progress=True, **kwargs):
    return _iresnet_v('cosface50_casia', IBasicBlock, [3, 4, 14, 3], pretrained, progress, **kwargs)
