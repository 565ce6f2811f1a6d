"""GPU parity of the drop-in MSML model against vectors generated from the reference model
(tests/golden/model_iresnet18.npz; weights are reproduced with oracle.detfill on both sides)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from gpu_util import assert_close, host, need_gpu
from oracle.detfill import det_labels, det_tensor, fill_state_dict_

pytestmark = pytest.mark.gpu


def _model(fp16=False, header="AMArcFace"):
    from msml_b200.backbones import MSML
    net = MSML("iresnet18", "unet", (1, 1, 1, 1), 97, fp16=fp16, header_type=header,
               header_params=(64.0, 0.5, 0.0, 0.0), fm_params=(3, 2, "sigmoid", "mul"))
    fill_state_dict_(net)
    return net.cuda()


@pytest.fixture
def no_tf32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_eval_forward_fp32_matches_reference(no_tf32):
    need_gpu()
    g = load_golden("model_iresnet18")
    net = _model().eval()
    x = det_tensor("model.x", (2, 3, 112, 112)).cuda()
    with torch.no_grad():
        feat, seg = net(x)
    assert feat.shape == (2, 512) and seg.shape == (2, 2, 112, 112)
    assert_close(host(feat), g["eval_feature"], 1e-3, atol_frac=1e-3, what="eval feature")
    assert_close(host(seg), g["eval_seg"], 1e-3, atol_frac=1e-3, what="eval seg")
    # argmax occlusion mask from the fused DAP kernel == argmax of the reference segmentation,
    # except where the reference's two logits are closer than fp32 conv noise
    with torch.no_grad():
        segs, mask = net.osb.forward_with_mask(x.contiguous(memory_format=torch.channels_last))
    ref_mask = (g["eval_seg"][:, 1] > g["eval_seg"][:, 0]).astype(np.int64)
    margin = np.abs(g["eval_seg"][:, 1] - g["eval_seg"][:, 0])
    differ = mask.cpu().numpy() != ref_mask
    assert not (differ & (margin > 1e-3 * np.abs(g["eval_seg"]).max())).any()
    assert torch.equal(mask, (segs[4][:, 1] > segs[4][:, 0]).long())     # bit-exact vs own logits


def test_train_step_fp32_matches_reference(no_tf32):
    need_gpu()
    g = load_golden("model_iresnet18")
    net = _model().train()
    x = det_tensor("model.x", (2, 3, 112, 112)).cuda()
    label = det_labels("model.l", 2, 97).cuda()
    final_cls, final_seg, kd = net(x, label)
    assert kd == 0.0
    loss = torch.nn.functional.cross_entropy(final_cls, label) + final_seg.float().mean()
    loss.backward()
    # the in-model head contracts in bf16 on tcgen05: logits carry bf16 rounding of the operands
    assert_close(host(final_cls), g["train_cls"], 2e-2, atol_frac=1e-2, what="train cls")
    assert_close(host(final_seg), g["train_seg"], 1e-3, atol_frac=1e-3, what="train seg")
    assert abs(float(loss) - float(g["loss"])) <= 2e-2 * abs(float(g["loss"]))
    named = dict(net.named_parameters())
    # frb.fc.bias feeds BatchNorm1d (batch statistics): its true gradient is zero, the stored one is
    # rounding noise in both implementations
    for key in [k[5:] for k in g if k.startswith("grad.") and k != "grad.frb.fc.bias"]:
        assert_close(host(named[key].grad), g["grad." + key], 5e-2, atol_frac=3e-2, what="grad " + key)
    checked = 0
    for key in [k[9:] for k in g if k.startswith("gradnorm.")]:
        if named[key].grad is None or key == "frb.fc.bias":
            continue
        want = float(g["gradnorm." + key])
        if want < 1e-3:          # parameters in front of a batch-statistics BN: gradient is cancellation noise
            continue
        got = float(named[key].grad.float().norm())
        assert abs(got - want) <= 5e-2 * want + 1e-6, (key, got, want)
        checked += 1
    assert checked > 200


def test_eval_forward_bf16_within_tolerance():
    need_gpu()
    g = load_golden("model_iresnet18")
    net = _model(fp16=True).eval()
    x = det_tensor("model.x", (2, 3, 112, 112)).cuda()
    with torch.no_grad():
        feat, seg = net(x)
    assert feat.dtype == torch.float32
    cos = torch.nn.functional.cosine_similarity(feat.cpu().double(), torch.from_numpy(g["eval_feature"]).double())
    assert (cos > 0.995).all(), cos
    assert_close(host(seg), g["eval_seg"], 5e-2, atol_frac=5e-2, what="bf16 seg")


def test_headless_train_returns_features_and_masks():
    need_gpu()
    net = _model(header=None).train()
    x = det_tensor("model.x", (2, 3, 112, 112)).cuda()
    feat, seg = net(x)
    assert feat.shape == (2, 512) and feat.requires_grad and seg.shape == (2, 2, 112, 112)
    feat.sum().backward()
    assert net.frb.fm_ops[0].same_conv.weight.grad is not None
    assert net.osb.conv1.weight.grad is None          # detach link: no seg loss => OSB gets no gradient


# ------------------------------------------------------------------------------- peer-guided training
PEER_ON = {"use_ori": True, "use_conv": True, "mask_trans": "conv", "use_decoder": False, "peer_pretrained": False}


def _peer_model():
    from msml_b200.backbones import MSML
    net = MSML("iresnet18", "unet", (1, 1, 1, 1), 97, fp16=False, header_type="AMArcFace", header_params=(64.0, 0.5, 0.0, 0.0),
               fm_params=(3, 2, "sigmoid", "mul"), peer_params=dict(PEER_ON))
    fill_state_dict_(net)
    return net.cuda()


def test_peer_network_and_ori_path_match_reference(no_tf32):
    """ref backbones/peer/arcface.py:159-194 (frozen teacher, four detached stage outputs) and the `ori` path through
    MSML.forward -> IResNet.forward -> FMCnn (ref msml.py:150-172, iresnet.py:203-223, fmoperator.py:293-302): the
    distillation loss kd, logits (+ kd, ref msml.py:171), and the gradients of the distillation convolutions, against
    vectors produced by the reference with a randomly initialised teacher (tests/golden/make_golden.py: gen_model_peer)."""
    need_gpu()
    g = load_golden("model_iresnet18_peer")
    net = _peer_model()
    assert not any(p.requires_grad for p in net.frb.peer.parameters())
    x = det_tensor("model.x", (2, 3, 112, 112)).cuda()
    ori = det_tensor("model.ori", (2, 3, 112, 112)).cuda()
    label = det_labels("model.l", 2, 97).cuda()
    net.eval()
    with torch.no_grad():
        feat, seg = net(x)
        pf, inter = net.frb.peer(ori)
    assert [tuple(t.shape[1:]) for t in inter] == [(64, 56, 56), (128, 28, 28), (256, 14, 14), (512, 7, 7)]
    assert not any(t.requires_grad for t in inter)
    assert_close(host(feat), g["eval_feature"], 1e-3, atol_frac=1e-3, what="eval feature")
    assert_close(host(seg), g["eval_seg"], 1e-3, atol_frac=1e-3, what="eval seg")
    assert_close(host(pf), g["peer_feature"], 1e-3, atol_frac=1e-3, what="peer feature")
    assert_close(host(inter[3]), g["peer_ft3"], 1e-3, atol_frac=1e-3, what="peer ft3")
    assert abs(float(inter[0].float().norm()) - float(g["peer_ft0_norm"])) <= 1e-3 * float(g["peer_ft0_norm"])
    net.train()
    final_cls, final_seg, kd = net(x, label, ori)
    assert abs(float(kd) - float(g["kd"])) <= 2e-3 * abs(float(g["kd"])), (float(kd), float(g["kd"]))
    loss = torch.nn.functional.cross_entropy(final_cls, label) + final_seg.float().mean()
    loss.backward()
    assert_close(host(final_cls), g["train_cls"], 2e-2, atol_frac=1e-2, what="train cls")
    assert_close(host(final_seg), g["train_seg"], 1e-3, atol_frac=1e-3, what="train seg")
    assert abs(float(loss) - float(g["loss"])) <= 2e-2 * abs(float(g["loss"]))
    named = dict(net.named_parameters())
    assert all(p.grad is None for p in net.frb.peer.parameters())
    for key in [k[5:] for k in g if k.startswith("grad.")]:
        assert_close(host(named[key].grad), g["grad." + key], 5e-2, atol_frac=3e-2, what="grad " + key)
    checked = 0
    for key in [k[9:] for k in g if k.startswith("gradnorm.")]:
        if named[key].grad is None or key == "frb.fc.bias":
            continue
        want = float(g["gradnorm." + key])
        if want < 1e-3:
            continue
        got = float(named[key].grad.float().norm())
        assert abs(got - want) <= 5e-2 * want + 1e-6, (key, got, want)
        checked += 1
    assert checked > 250


def test_peer_constructor_contract():
    """As the reference: pretrained peers by default (FileNotFoundError without the unshipped files, ref arcface.py:204-206),
    `ori` without a peer is an error, a Softmax head cannot choose a peer (ref iresnet.py:145-146)."""
    need_gpu()
    from msml_b200.backbones import MSML
    from msml_b200.backbones.peer import arcface18, arcface50
    with pytest.raises(FileNotFoundError):
        arcface18()
    p = arcface50(pretrained=False)
    assert not p.training and sum(1 for _ in p.layer3) == 14
    with pytest.raises(FileNotFoundError):
        MSML("iresnet18", "unet", (1, 1, 1, 1), 97, header_type="AMArcFace", peer_params={k: v for k, v in PEER_ON.items() if k != "peer_pretrained"})
    with pytest.raises(ValueError):
        MSML("iresnet18", "unet", (1, 1, 1, 1), 97, header_type="Softmax", peer_params=dict(PEER_ON))
    net = _model().train()
    with pytest.raises(RuntimeError):
        net(det_tensor("model.x", (2, 3, 112, 112)).cuda(), det_labels("model.l", 2, 97).cuda(), det_tensor("model.ori", (2, 3, 112, 112)).cuda())
