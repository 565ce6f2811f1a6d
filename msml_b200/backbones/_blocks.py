"""Building blocks shared by the FRB trunk and the OSB encoder (both use the same improved-residual
unit in the reference: ref backbones/frb/iresnet.py:38-67 and backbones/osb/unet.py:62-91).
Module / parameter names are kept so reference checkpoints load unchanged; the BatchNorm / PReLU /
residual-add chains between the convolutions run as the fused NHWC kernels of csrc/bn_act.cu
(ops.bn_act: statistics + normalise + add + activation in two passes; a unit whose output feeds the
next unit's bn1 also hands over that tensor's batch statistics, msml_bn_fwd_ex), the convolutions
on cuDNN."""
from torch import nn

from .. import ops


def conv3x3(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False)


def conv1x1(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, 1, stride=stride, bias=False)


class IBasicBlock(nn.Module):
    """BN -> conv3x3 -> BN -> PReLU -> conv3x3(stride) -> BN, plus (projected) identity."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.feeds_bn = False       # the unit's output goes straight into another BatchNorm (make_stage / the trunk set it)
        self.bn1 = nn.BatchNorm2d(inplanes, eps=1e-05)
        self.conv1 = conv3x3(inplanes, planes)
        self.bn2 = nn.BatchNorm2d(planes, eps=1e-05)
        self.prelu = nn.PReLU(planes)
        self.conv2 = conv3x3(planes, planes, stride)
        self.bn3 = nn.BatchNorm2d(planes, eps=1e-05)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        y, x = ops.bn_act_fork(x, self.bn1)                    # x feeds bn1 and the skip: one fused gradient sum in backward
        out = ops.conv2d(y, self.conv1)
        out = ops.conv2d(ops.bn_act(out, self.bn2, self.prelu), self.conv2)
        skip = x if self.downsample is None else ops.bn_act(ops.conv2d(x, self.downsample[0]), self.downsample[1])
        # bn3(out) + identity; when the next unit's bn1 (or the trunk's bn2) reads it, its batch statistics come along
        return ops.bn_act(out, self.bn3, None, skip, emit_next_stats=self.feeds_bn)


def make_stage(inplanes, planes, blocks, stride, feeds_bn=False):
    """One resolution stage: the first block strides / projects, the rest keep the shape.  ``feeds_bn``: the stage's
    output is read by a BatchNorm next (the following stage's bn1 or the trunk's bn2, with nothing in between)."""
    down = None
    if stride != 1 or inplanes != planes:
        down = nn.Sequential(conv1x1(inplanes, planes, stride), nn.BatchNorm2d(planes, eps=1e-05))
    layers = [IBasicBlock(inplanes, planes, stride, down)]
    layers += [IBasicBlock(planes, planes) for _ in range(1, blocks)]
    for unit in layers[:-1]:
        unit.feeds_bn = True
    layers[-1].feeds_bn = bool(feeds_bn)
    return nn.Sequential(*layers)
