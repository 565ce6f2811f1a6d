"""GPU parity of the in-model margin heads against vectors generated from the reference
(tests/golden/margins.npz <- ref headers/margin_losses.py: Softmax :41-68, AMCosFace :241-305, AMArcFace :356-418; the
label vector [-1, 4, -1, 5, 3, -1] is the reference's own fixture, :432-439).

Two levels: (1) the margin kernels alone (msml_margin_fwd / msml_margin_bwd) on an fp32 cosine matrix, at fp32 tolerance —
this is where k != 0, the label -1 rows and the AM derivative are pinned; (2) the drop-in modules end to end (tcgen05
contraction with bf16 operands + the margin kernels) at the bf16 tolerance north_star states (rtol 2e-2, plus an absolute
floor because rtol alone is meaningless for near-zero logits)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden
from gpu_util import assert_close, dev, host, need_gpu

pytestmark = pytest.mark.gpu
TAGS = ["arc_p", "arc_f", "arc_am_p", "arc_am_f", "cos_p", "cos_f", "cos_am_p", "cos_am_f"]


def _case(g, tag):
    kind = "arc" if tag.startswith("arc") else "cos"
    s, m, a, k = (float(v) for v in g[tag + ".smak"])
    return kind, s, m, a, k


@pytest.mark.parametrize("tag", TAGS)
def test_margin_kernels_fp32_vs_reference(tag):
    """msml_margin_fwd / msml_margin_bwd on the fp32 cosine torch computes from the golden inputs: logits and, through
    autograd around the two normalisations and the matmul, the reference's embedding / weight gradients."""
    need_gpu()
    from msml_b200 import ops
    g = load_golden("margins")
    kind, s, m, a, k = _case(g, tag)
    e = dev(g[tag + ".e"]).requires_grad_(True)
    w = dev(g[tag + ".w"]).requires_grad_(True)
    label = dev(g[tag + ".label"])
    cos = F.normalize(e) @ F.normalize(w).t()
    logits = ops.margin_logits(cos, label, kind, s, m, a, k)
    assert_close(host(logits), g[tag + ".logits"], 1e-4, atol=2e-5 * s, what=tag + " logits")
    logits.backward(dev(g[tag + ".dl"]))
    assert_close(host(e.grad), g[tag + ".de"], 1e-3, atol=1e-5 * s, what=tag + " de")
    assert_close(host(w.grad), g[tag + ".dw"], 1e-3, atol=1e-5 * s, what=tag + " dw")
    rows = np.nonzero(g[tag + ".label"] == -1)[0]
    if rows.size:                                       # label -1: scale only, bit for bit (ref :395-397 skips those rows)
        assert torch.equal(logits[rows], (cos * s)[rows])


@pytest.mark.parametrize("tag", TAGS)
def test_am_head_modules_match_reference(tag):
    """AMArcFace / AMCosFace drop-ins: same constructor, same forward(embedding, label), tcgen05 contraction inside."""
    need_gpu()
    from msml_b200.headers import AMArcFace, AMCosFace
    g = load_golden("margins")
    kind, s, m, a, k = _case(g, tag)
    C, D = g[tag + ".w"].shape
    head = (AMArcFace if kind == "arc" else AMCosFace)(D, C, None, s=s, m=m, a=a, k=k).cuda()
    with torch.no_grad():
        head.weight.copy_(dev(g[tag + ".w"]))
    e = dev(g[tag + ".e"]).requires_grad_(True)
    logits = head(e, dev(g[tag + ".label"]))
    assert logits.shape == (e.shape[0], C) and logits.dtype == torch.float32
    # bf16 operands: |d cos| <= ~2^-8, so |d logit| <= s * 4e-3 away from the margin's steep region
    assert_close(host(logits), g[tag + ".logits"], 2e-2, atol=1e-2 * s, what=tag + " logits")
    logits.backward(dev(g[tag + ".dl"]))
    assert_close(host(e.grad), g[tag + ".de"], 2e-2, atol_frac=1e-2, what=tag + " de")
    assert_close(host(head.weight.grad), g[tag + ".dw"], 2e-2, atol_frac=1e-2, what=tag + " dw")


def test_softmax_head_module_matches_reference():
    need_gpu()
    from msml_b200.headers import Softmax
    g = load_golden("margins")
    C, D = g["softmax.w"].shape
    head = Softmax(D, C, None).cuda()
    with torch.no_grad():
        head.weight.copy_(dev(g["softmax.w"]))
        head.bias.copy_(dev(g["softmax.b"]))
    e = dev(g["softmax.e"]).requires_grad_(True)
    out = head(e, None)
    assert_close(host(out), g["softmax.logits"], 2e-2, atol_frac=1e-2, what="softmax logits")
    # gradients of a plain linear layer: dE = dL W, dW = dL^T E, db = sum dL  (checked against torch on the same operands)
    dl = torch.randn_like(out)
    out.backward(dl)
    assert_close(host(e.grad), host(dl @ head.weight.detach()), 2e-2, atol_frac=1e-2, what="softmax de")
    assert_close(host(head.weight.grad), host(dl.t() @ e.detach()), 2e-2, atol_frac=1e-2, what="softmax dw")
    assert_close(host(head.bias.grad), host(dl.sum(0)), 1e-5, atol=1e-6, what="softmax db")
    with pytest.raises(ValueError):
        Softmax(D, C, [0]).cuda()(e, None)


@pytest.mark.parametrize("kind,smak", [("arc", (64.0, 0.5, 0.0, 0.0)), ("arc", (32.0, 0.45, 1.2, 0.1)),
                                       ("cos", (64.0, 0.4, 0.0, 0.0)), ("cos", (32.0, 0.35, 1.2, 0.1))])
def test_margin_kernels_vs_oracle_at_model_size(kind, smak):
    """BASELINE config-1 head size (B = 8 ... 64 rows x 10,572 classes), rows with label -1 mixed in."""
    need_gpu()
    from msml_b200 import ops
    from oracle import margins as om
    torch.manual_seed(17)
    B, C = 64, 10572
    cos = (torch.rand(B, C, device="cuda") * 1.9 - 0.95).requires_grad_(True)
    label = torch.randint(0, C, (B,), device="cuda")
    label[::5] = -1
    out = ops.margin_logits(cos, label, kind, *smak)
    want = om.margin_apply(host(cos), label.cpu().numpy(), kind, *smak)
    assert_close(host(out), want, 1e-5, atol=1e-5 * smak[0], what="logits")
    dl = torch.randn_like(out)
    out.backward(dl)
    want_d = host(dl) * om.margin_dcos(host(cos), label.cpu().numpy(), kind, *smak)
    assert_close(host(cos.grad), want_d, 1e-4, atol=1e-5 * smak[0], what="dcos")


def test_target_cosine_at_the_clamp_boundary():
    """Documented deviation from the reference's NaN policy: the reference calls acos without a clamp
    (margin_losses.py:413,415), so a target cosine that rounding pushes past 1 becomes NaN; here the target cosine is
    clamped to +-(1 - 1e-6) first (bf16 operands make |cos| > 1 a practical case, not a theoretical one).  Pinned: the
    values AT the boundary are the clamped ones, finite, forward and backward; inside the interval nothing changes."""
    need_gpu()
    from msml_b200 import ops
    from oracle import margins as om
    s, m = 64.0, 0.5
    c = torch.tensor([[1.0, 0.1], [1.0 + 4e-3, 0.1], [-1.0, 0.1], [-1.0 - 4e-3, 0.1], [0.999, 0.1]], device="cuda", requires_grad=True)
    label = torch.zeros(5, dtype=torch.int64, device="cuda")
    for kind in ("arc", "cos"):
        c.grad = None
        out = ops.margin_logits(c, label, kind, s, m)
        out.sum().backward()
        assert torch.isfinite(out).all() and torch.isfinite(c.grad).all()
        lim = 1.0 - 1e-6
        clamped = np.array([[lim, 0.1], [lim, 0.1], [-lim, 0.1], [-lim, 0.1], [0.999, 0.1]])
        want = om.margin_apply(clamped, label.cpu().numpy(), kind, s, m)
        # acos is ill-conditioned at the boundary: theta = 1.4e-3 there, and fp32 holds 1 - 1e-6 to 6e-8
        assert_close(host(out), want, 1e-3, atol=1e-3 * s, what=kind + " boundary logits")
        assert torch.equal(out[:, 1], c.detach()[:, 1] * s)
        inside = om.margin_dcos(clamped[4:], label.cpu().numpy()[4:], kind, s, m)
        assert_close(host(c.grad[4:]), inside, 1e-3, what=kind + " gradient inside the interval")
