"""MSML top-level model — drop-in for ref backbones/msml.py (:14-174).

    OSB (U-Net) -> 4 detached multi-scale occlusion maps + final segmentation
    FRB (iResNet) with an FM operator after each stage, gated by those maps     <- fused CUDA tail
    head: in-model Softmax / AMCosFace / AMArcFace, or none (PartialFC outside)

forward(x, label=None, ori=None):
    eval                      -> (feature (B,512), final_seg (B,2,112,112))           ref :173-174
    train, in-model head      -> (final_cls, final_seg, kd)                            ref :170-172
    train, header_type=None   -> (feature, final_seg)   for PartialFC.forward_backward (north_star;
                                 the reference's commented PartialFC step, train.py:282-318, needs it)

B200 specifics: the whole model runs channels-last (NHWC) so the fused FM kernels see 128-bit
coalesced channel vectors; ``fp16=True`` selects bf16 autocast for OSB + FRB (fp32 master weights;
bf16 needs no loss scaling), the 512-d feature is returned in fp32 as in the reference (ref :169).
"""
import torch

from .. import ops
import torch.nn as nn

from .fm import FMCnn, FMNone
from .frb import iresnet18, iresnet34, iresnet50
from .osb import unet
from ..headers.margin_losses import AMArcFace, AMCosFace, Softmax

__all__ = ['MSML']


class MSML(nn.Module):
    frb_type_list = ('lightcnn', 'iresnet18', 'iresnet34', 'iresnet50',)
    osb_type_list = ('unet',)
    head_type_list = ('Softmax', 'AMArcFace', 'AMCosFace')

    def __init__(self, frb_type: str, osb_type: str, fm_layers: tuple, num_classes: int, fp16: bool = False,
                 frb_pretrained: bool = False, fm_params: tuple = (3, 2, 'tanh', 'add'),
                 header_type: str = 'Softmax', header_params: tuple = (64.0, 0.5, 0.0, 0.0),
                 dropout: float = 0., use_osb: bool = True, peer_params: dict = None):
        super().__init__()
        assert len(fm_layers) == 4
        peer_params = dict(peer_params or {'use_ori': False, 'use_conv': False, 'mask_trans': 'conv', 'use_decoder': False})
        self._prepare_shapes(frb_type, osb_type)
        self._prepare_fm(fm_layers, fm_params, peer_params)
        self._prepare_frb(frb_type, dropout, peer_params, header_type, pretrained=frb_pretrained)
        self._prepare_osb(osb_type)
        self.num_classes = num_classes
        self._prepare_header(header_type, header_params)
        self.fp16 = fp16
        self.use_osb = use_osb
        self.to(memory_format=torch.channels_last)

    def _prepare_shapes(self, frb_type, osb_type):
        if 'lightcnn' in frb_type:
            raise ValueError('FRB type error: the LightCNN trunk is outside the B200 hot path (SURVEY.md section 2)')
        elif 'iresnet' in frb_type:
            self.input_size, self.gray = 112, False
            self.heights, self.f_channels, self.dim_feature = (56, 28, 14, 7), (64, 128, 256, 512), 512
        else:
            raise ValueError('FRB type error')
        if 'unet' in osb_type:
            self.s_channels = (18, 18, 18, 18)
        else:
            raise ValueError('OSB type error')

    def _prepare_fm(self, fm_layers, fm_params, peer_params):
        fm_ops = []
        for i, fm_type in enumerate(fm_layers):
            if fm_type == 0:
                fm_ops.append(FMNone())
            elif fm_type == 1:
                kernel_size, num_res, act, arith = fm_params
                fm_ops.append(FMCnn(self.heights[i], self.heights[i], self.f_channels[i], kernel_size=kernel_size,
                                    resblocks=num_res, activation=act, arith_strategy=arith, peer_params=peer_params))
            else:
                raise ValueError('FM Operators type error')
        self.fm_ops = fm_ops

    def _prepare_frb(self, frb_type, dropout=0., peer_params: dict = None, header_type: str = "", pretrained=False):
        peer_params["header_type"] = header_type or ""
        ctor = {'18': iresnet18, '34': iresnet34, '50': iresnet50}
        for tag, fn in ctor.items():
            if tag in frb_type:
                self.frb = fn(self.fm_ops, pretrained=pretrained, dropout=dropout, peer_params=peer_params)
                return
        raise ValueError('IResNet type {} not found'.format(frb_type))

    def _prepare_osb(self, osb_type):
        if 'unet' in osb_type:
            self.osb = unet(backbone='r18', gray=self.gray, input_size=self.input_size)

    def _prepare_header(self, head_type, header_params):
        if head_type is None or str(head_type).lower() in ('none', 'partialfc'):
            self.classification = None          # head lives outside the model (PartialFC)
            return
        assert head_type in self.head_type_list
        s, m, a, k = header_params
        if 'Softmax' in head_type:
            self.classification = Softmax(self.dim_feature, self.num_classes, device_id=None)
        elif 'AMCosFace' in head_type:
            self.classification = AMCosFace(self.dim_feature, self.num_classes, device_id=None, s=s, m=m, a=a, k=k)
        elif 'AMArcFace' in head_type:
            self.classification = AMArcFace(self.dim_feature, self.num_classes, device_id=None, s=s, m=m, a=a, k=k)
        else:
            raise ValueError('Header type error!')

    def forward(self, x, label=None, ori=None):
        x = x.contiguous(memory_format=torch.channels_last)
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=bool(self.fp16)):
            if self.fp16 and x.is_cuda:
                # one bf16 copy of the image with the channels zero-padded 3 -> 8: both stem convolutions (OSB, FRB) then run
                # cuDNN's tensor-core kernels instead of the 3-channel generic ones (their weights get zero input channels)
                x, _ = ops.cat_channels_padded((x.to(torch.bfloat16),))
            segs_ready = None
            if self.use_osb:
                if ops.side_stream_enabled() and x.is_cuda and self.training:
                    # the FRB needs the segmentation maps only at its first FM operator: the OSB runs on the side stream
                    # next to the FRB stem and stage 1 (both are chains of latency-bound kernels that leave SMs idle)
                    seg_list, segs_ready = ops.run_on_side_stream(self.osb, x)
                else:
                    seg_list = self.osb(x)          # [seg0, seg1, seg2, seg3, seg5] small to big
                final_seg = seg_list[4]
                segs = seg_list[3::-1]              # [seg3, seg2, seg1, seg0] big to small
            else:
                segs, final_seg = (None, None, None, None), None
            feature, kd = self.frb(x, segs, ori, segs_ready)
        feature = feature.float()
        if self.training and self.classification is not None:
            final_cls = self.classification(feature, label) + kd
            return final_cls, final_seg, kd
        return feature, final_seg
