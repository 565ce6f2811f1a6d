"""Face-Recognition Branch — drop-in for ref backbones/frb/iresnet.py (IResNet :70-236,
iresnet18/34/50 :444-481): iResNet trunk with one Feature-Masking operator after each of the
four stages (ref :213-223).  The FM operators carry the fused CUDA tail (see fm/fmoperator.py).

Peer-guided training (ref :131-146,203-206): ``peer_params['use_ori']`` builds the frozen teacher of
backbones/peer/arcface.py exactly as the reference does (arcface18/34/50 for an Arc head, cosface50_casia for a Cos head
on iresnet50) and feeds its four detached stage outputs to the FM operators when ``ori`` is given.  The reference loads the
teacher from ``./backbones/pretrained/*.pth``, which it does not ship (FileNotFoundError, as there);
``peer_params['peer_pretrained'] = False`` (an addition) builds the same teacher with random weights, and ``set_peer``
injects any module with the ``(feature, [ft0..ft3])`` contract.  The image decoder (``use_decoder``) stays out of scope:
its loss is dropped by the reference itself (tuple bug, ref :228).
"""
import torch
from torch import nn

from ... import ops
from .._blocks import IBasicBlock, make_stage

__all__ = ['IResNet', 'iresnet18', 'iresnet34', 'iresnet50']


class IResNet(nn.Module):
    fc_scale = 7 * 7

    def __init__(self, block, layers, fm_ops, dim_feature=512, dropout=0, zero_init_residual=False,
                 fp16=False, peer_params: dict = None):
        super().__init__()
        del block
        self.fp16 = fp16
        self.conv1 = nn.Conv2d(3, 64, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(64, eps=1e-05)
        self.prelu = nn.PReLU(64)
        self.layer1 = make_stage(64, 64, layers[0], 2)
        self.layer2 = make_stage(64, 128, layers[1], 2)
        self.layer3 = make_stage(128, 256, layers[2], 2)
        self.layer4 = make_stage(256, 512, layers[3], 2)
        self.bn2 = nn.BatchNorm2d(512, eps=1e-05)
        self.dropout = nn.Dropout(p=dropout, inplace=True)
        self.fc = nn.Linear(512 * self.fc_scale, dim_feature)
        self.features = nn.BatchNorm1d(dim_feature, eps=1e-05)
        nn.init.constant_(self.features.weight, 1.0)
        self.features.weight.requires_grad = False

        assert len(fm_ops) == 4
        self.fm_ops = nn.ModuleList(fm_ops)

        peer_params = peer_params or {}
        self.peer = None      # peer type is consistent with msml.header_type (ref :131-146)
        self.header_type = str(peer_params.get('header_type', '')).lower()
        if peer_params.get('use_ori'):
            from ..peer import arcface18, arcface34, arcface50, cosface50_casia
            pre = bool(peer_params.get('peer_pretrained', True))
            layers = list(layers)
            if 'arc' in self.header_type:
                ctor = {(2, 2, 2, 2): arcface18, (3, 4, 6, 3): arcface34, (3, 4, 14, 3): arcface50}.get(tuple(layers))
                if ctor is not None:
                    self.peer = ctor(pretrained=pre).requires_grad_(False)
            elif 'cos' in self.header_type:
                if layers == [3, 4, 14, 3]:
                    self.peer = cosface50_casia(pretrained=pre).requires_grad_(False)
            else:
                raise ValueError('Error type of iresnet, cannot decide peer network.')
        self.use_decoder = bool(peer_params.get('use_decoder'))
        # NOTE (parity with the code as written): like the reference, the teacher is registered BEFORE the initialisation
        # loop below, so its convolutions are re-drawn from N(0, 0.1) and its BatchNorm affine parameters reset to (1, 0)
        # even when pretrained weights were just loaded (ref :131-146 then :152-157) — only PReLU slopes, fc and the BN
        # running statistics of the checkpoint survive.  Load the teacher's state_dict after construction to avoid that.

        for m in self.modules():     # ref :157-162
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, 0, 0.1)
            elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        if zero_init_residual:
            for m in self.modules():
                if isinstance(m, IBasicBlock):
                    nn.init.constant_(m.bn2.weight, 0)

    def set_peer(self, peer: nn.Module):
        """Inject a frozen teacher returning (feature, [ft0..ft3]) (ref backbones/peer/arcface.py)."""
        self.peer = peer.requires_grad_(False)

    def forward(self, x, segs, ori, segs_ready=None):
        ft = (None, None, None, None)
        if ori is not None:
            if self.peer is None:
                raise RuntimeError("`ori` given but this FRB has no peer network: build MSML with peer_params['use_ori'] = True "
                                   "(and an Arc / Cos head), or inject one with IResNet.set_peer()")
            _, ft = self.peer(ori)
        x = ops.bn_act(ops.conv2d_padded_in(x, self.conv1, x.shape[1] - self.conv1.in_channels), self.bn1, self.prelu,
                       emit_next_stats=True)      # layer1's first bn1 reads this tensor next
        kd_terms = []
        for i, layer in enumerate((self.layer1, self.layer2, self.layer3, self.layer4)):
            x = ops.grad_marker(x, i)       # backward: everything from stage i on has its parameter gradients complete
            x = layer(x)
            if segs_ready is not None:          # the OSB ran on a side stream: join it before its maps are first read
                segs_ready()
                segs_ready = None
            x, l = self.fm_ops[i](x, segs[i], ft[i])
            kd_terms.append(l)
        if segs_ready is not None:
            segs_ready()
        x = ops.bn_act(x, self.bn2)
        x = self.dropout(torch.flatten(x, 1))
        x = self.features(ops.linear(x.float(), self.fc))
        kd = sum(kd_terms) if ori is not None and all(t is not None for t in kd_terms) else 0.
        return x, kd * 1.0


def _iresnet(layers, fm_ops, pretrained, **kwargs):
    if pretrained:
        raise FileNotFoundError('pretrained FRB weights are not shipped; load a state_dict explicitly')
    return IResNet(IBasicBlock, layers, fm_ops, **kwargs)


def iresnet18(fm_ops, pretrained=False, dim_feature=512, dropout=0., peer_params=None):
    return _iresnet([2, 2, 2, 2], fm_ops, pretrained, dim_feature=dim_feature, dropout=dropout, peer_params=peer_params)


def iresnet34(fm_ops, pretrained=False, dim_feature=512, dropout=0., peer_params=None):
    return _iresnet([3, 4, 6, 3], fm_ops, pretrained, dim_feature=dim_feature, dropout=dropout, peer_params=peer_params)


def iresnet50(fm_ops, pretrained=False, dim_feature=512, dropout=0., peer_params=None):
    return _iresnet([3, 4, 14, 3], fm_ops, pretrained, dim_feature=dim_feature, dropout=dropout, peer_params=peer_params)
