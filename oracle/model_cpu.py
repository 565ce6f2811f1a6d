"""Oracle: plain torch-CPU fp32 restatement of the MSML backbone + in-model head, written as pure
functions of a state_dict.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by the tests
as the end-to-end checker for model sizes that have no committed golden (ires50) and by
bench.py's cpu_baseline / --impl reference legs as the timed CPU port of the reference's step.

Follows (op for op, fp32, NCHW, no fusion):
  IBasicBlock        ref backbones/frb/iresnet.py:38-67 (same unit in backbones/osb/unet.py:62-91)
  Unet.forward       ref backbones/osb/unet.py:189-240, DAP :158-161
  resblock_bottle    ref backbones/fm/fmoperator.py:35-68
  FMCnn.forward      ref backbones/fm/fmoperator.py:277-311   (peer branch off: use_ori False)
  IResNet.forward    ref backbones/frb/iresnet.py:190-236
  MSML.forward       ref backbones/msml.py:150-174
  AMArcFace/AMCosFace ref headers/margin_losses.py:241-305,356-418
Pinned by tests/test_oracle_golden.py::test_model_cpu_matches_reference against
tests/golden/model_iresnet18.npz (outputs of the reference model itself).
"""
import torch
import torch.nn.functional as F

LAYERS = {"iresnet18": [2, 2, 2, 2], "iresnet34": [3, 4, 6, 3], "iresnet50": [3, 4, 14, 3]}


class _P:
    """state_dict view with a key prefix; BN running stats are updated in place in train mode."""

    def __init__(self, sd, prefix="", training=False):
        self.sd, self.prefix, self.training = sd, prefix, training

    def sub(self, name):
        return _P(self.sd, self.prefix + name + ".", self.training)

    def __getitem__(self, name):
        return self.sd[self.prefix + name]

    def has(self, name):
        return (self.prefix + name) in self.sd


def _bn(p, x):
    return F.batch_norm(x, p["running_mean"], p["running_var"], p["weight"], p["bias"], p.training, 0.1, 1e-5)


def _prelu(p, x):
    return F.prelu(x, p["weight"])


def _basic_block(p, x, stride):
    out = _bn(p.sub("bn1"), x)
    out = F.conv2d(out, p["conv1.weight"], None, 1, 1)
    out = _prelu(p.sub("prelu"), _bn(p.sub("bn2"), out))
    out = F.conv2d(out, p["conv2.weight"], None, stride, 1)
    out = _bn(p.sub("bn3"), out)
    if p.has("downsample.0.weight"):
        x = _bn(p.sub("downsample.1"), F.conv2d(x, p["downsample.0.weight"], None, stride, 0))
    return out + x


def _stage(p, x, blocks):
    for i in range(blocks):
        x = _basic_block(p.sub(str(i)), x, 2 if i == 0 else 1)
    return x


def _gcm(p, x):
    def c(name, t, pad):
        return F.conv2d(t, p[name + ".weight"], p[name + ".bias"], 1, pad)
    left = c("conv_l2", c("conv_l1", x, (3, 0)), (0, 3))
    right = c("conv_r2", c("conv_r1", x, (0, 3)), (3, 0))
    return left + right


def dap(x, k=3):
    B, CK, H, W = x.shape
    return x.view(B, CK // (k * k), k * k, H, W).mean(2)


def unet_forward(p, x):
    x0 = _prelu(p.sub("prelu"), _bn(p.sub("bn1"), F.conv2d(x, p["conv1.weight"], None, 2, 1)))
    x1 = _stage(p.sub("layer1"), x0, 2)
    x2 = _stage(p.sub("layer2"), x1, 2)
    x3 = _stage(p.sub("layer3"), x2, 2)
    x4 = _bn(p.sub("bn2"), _stage(p.sub("layer4"), x3, 2))
    up = lambda name, t: F.conv_transpose2d(t, p[name + ".weight"], None, 2, 1)
    seg0 = up("deconv1", _gcm(p.sub("gcm1"), x4))
    seg1 = up("deconv2", torch.cat((seg0, _gcm(p.sub("gcm2"), x3)), 1))
    seg2 = up("deconv3", torch.cat((seg1, _gcm(p.sub("gcm3"), x2)), 1))
    seg3 = up("deconv4", torch.cat((seg2, _gcm(p.sub("gcm4"), x1)), 1))
    seg5 = dap(up("deconv5", torch.cat((seg3, _gcm(p.sub("gcm5"), x0)), 1)))
    return [seg0.detach(), seg1.detach(), seg2.detach(), seg3.detach(), seg5]


def _bottle(p, x):
    y = _prelu(p.sub("prelu1"), _bn(p.sub("bn1"), F.conv2d(x, p["conv1.weight"])))
    y = _prelu(p.sub("prelu2"), _bn(p.sub("bn2"), F.conv2d(y, p["conv2.weight"], None, 1, 1)))
    y = _bn(p.sub("bn3"), F.conv2d(y, p["conv3.weight"]))
    return _prelu(p.sub("prelu3"), y + x)


def fm_forward(p, yf, yo, act="sigmoid", arith="mul", resblocks=2):
    w = p["same_conv.weight"]
    z = F.conv2d(torch.cat((yf, yo), 1), w, None, 1, w.shape[-1] // 2)
    for i in range(resblocks):
        z = _bottle(p.sub("res_block.%d" % i), z)
    g = torch.sigmoid(z) if act == "sigmoid" else torch.tanh(z)
    x = {"add": yf + g, "sub": yf - g, "div": yf / g, "mul": yf * g}[arith]
    return x + yf


def iresnet_forward(p, x, segs, layers, fm_layers=(1, 1, 1, 1), fm_params=(3, 2, "sigmoid", "mul")):
    x = _prelu(p.sub("prelu"), _bn(p.sub("bn1"), F.conv2d(x, p["conv1.weight"], None, 1, 1)))
    for i in range(4):
        x = _stage(p.sub("layer%d" % (i + 1)), x, layers[i])
        if fm_layers[i]:
            x = fm_forward(p.sub("fm_ops.%d" % i), x, segs[i], fm_params[2], fm_params[3], fm_params[1])
    x = torch.flatten(_bn(p.sub("bn2"), x), 1)
    x = F.linear(x, p["fc.weight"], p["fc.bias"])
    f = p.sub("features")
    return F.batch_norm(x, f["running_mean"], f["running_var"], f["weight"], f["bias"], p.training, 0.1, 1e-5)


def am_head(weight, feature, label, kind, s, m, a=0.0, k=0.0):
    cos = F.linear(F.normalize(feature), F.normalize(weight))
    rows = torch.where(label != -1)[0]
    theta_t = torch.acos(cos[rows, label[rows]])
    m_eff = m - k * (theta_t - a)
    if kind == "arc":
        tgt = torch.cos(theta_t + m_eff)
    else:
        tgt = cos[rows, label[rows]] - m_eff
    out = cos.clone()
    out[rows, label[rows]] = tgt
    return out * s


def msml_forward(sd, x, frb_type="iresnet18", training=False, fm_layers=(1, 1, 1, 1),
                 fm_params=(3, 2, "sigmoid", "mul")):
    """-> (feature (B,512), final_seg (B,2,H,W)) from a state_dict with the reference's key names."""
    p = _P(sd, "", training)
    seg = unet_forward(p.sub("osb"), x)
    feature = iresnet_forward(p.sub("frb"), x, seg[3::-1], LAYERS[frb_type], fm_layers, fm_params)
    return feature, seg[4]


def trainable_state(module_or_sd):
    """Detached fp32 CPU copy of a state_dict with requires_grad on the floating parameters."""
    sd = module_or_sd if isinstance(module_or_sd, dict) else module_or_sd.state_dict()
    out = {}
    for k, v in sd.items():
        t = v.detach().float().cpu().clone() if v.is_floating_point() else v.detach().cpu().clone()
        leaf = k.rsplit(".", 1)[-1]
        if t.is_floating_point() and leaf not in ("running_mean", "running_var"):
            t.requires_grad_(True)
        out[k] = t
    return out
