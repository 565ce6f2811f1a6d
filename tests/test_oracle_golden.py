"""Pins the CPU oracle (oracle/) against vectors produced by running the reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import consensus, dap, fm_tail, margins, partial_fc

FM = ["fm_c64_sigmoid_mul", "fm_c128_tanh_add", "fm_c64_sigmoid_div", "fm_c256_tanh_sub",
      "fm_c512_sigmoid_mul"]


def close(a, b, rtol=1e-5, atol=1e-6):
    np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol)


@pytest.mark.parametrize("name", FM)
def test_fm_tail_matches_reference(name):
    g = load_golden(name)
    act, arith = str(g["act"]), str(g["arith"])
    close(fm_tail.fm_gate_fwd(g["yf"], g["z"], act, arith), g["out"], 2e-5, 2e-5)
    dyf, dz = fm_tail.fm_gate_bwd(g["dout"], g["yf"], g["z"], act, arith)
    close(dyf, g["dyf_direct"], 1e-4, 2e-5)
    close(dz, g["dz"], 1e-4, 2e-5)


def test_fm_mask_extension_identity_case():
    # with Cm == C and (Hm, Wm) == (H, W) the resize/broadcast extension is the plain tail
    g = load_golden("fm_c64_sigmoid_mul")
    yf = g["yf"].transpose(0, 2, 3, 1)
    z = g["z"].transpose(0, 2, 3, 1)
    d = g["dout"].transpose(0, 2, 3, 1)
    close(fm_tail.fm_mask_fwd(yf, z), g["out"].transpose(0, 2, 3, 1), 2e-5, 2e-5)
    dyf, dm = fm_tail.fm_mask_bwd(d, yf, z)
    close(dm, g["dz"].transpose(0, 2, 3, 1), 1e-4, 2e-5)


def test_dap_matches_reference():
    g = load_golden("dap")
    y = dap.dap_fwd(g["x"])
    close(y, g["y"], 1e-6, 1e-6)
    close(dap.dap_bwd(g["dy"]), g["dx"], 1e-6, 1e-7)
    # argmax on the reference's own fp32 y is bit-exact, including the planted ties
    assert np.array_equal(dap.argmax_mask(g["y"]), g["mask"])
    assert g["mask"][0, 0, 0] == 0 and g["mask"][1, 3, 4] == 0


@pytest.mark.parametrize("tag", ["arc_p", "arc_f", "arc_am_p", "arc_am_f", "cos_p", "cos_f", "cos_am_p", "cos_am_f"])
def test_margin_heads_match_reference(tag):
    g = load_golden("margins")
    kind = "arc" if tag.startswith("arc") else "cos"
    s, m, a, k = g[tag + ".smak"]
    logits, *_ = margins.am_head_fwd(g[tag + ".e"], g[tag + ".w"], g[tag + ".label"], kind, s, m, a, k)
    close(logits, g[tag + ".logits"], 2e-5, 2e-5 * s)
    de, dw = margins.am_head_bwd(g[tag + ".e"], g[tag + ".w"], g[tag + ".label"], g[tag + ".dl"], kind, s, m, a, k)
    close(de, g[tag + ".de"], 2e-4, 1e-5 * s)
    close(dw, g[tag + ".dw"], 2e-4, 1e-5 * s)


def test_softmax_head_matches_reference():
    g = load_golden("margins")
    close(margins.softmax_head_fwd(g["softmax.e"], g["softmax.w"], g["softmax.b"]), g["softmax.logits"], 1e-5, 1e-6)


PFC = ["pfc_w1_full", "pfc_w1_sample", "pfc_w2_full", "pfc_w2_sample", "pfc_w2_am", "pfc_w1_d512",
       "pfc_w1_overflow"]


@pytest.mark.parametrize("name", PFC)
def test_partial_fc_step_matches_reference(name):
    g = load_golden(name)
    W, C, sr = int(g["W"]), int(g["C"]), float(g["sample_rate"])
    kind = str(g["kind"])
    s, m, a, k = g["smak"]
    # step 0 only depends on the initial weights; later steps on the ref's own updated weights
    weights = [g[f"r{r}.w0"] for r in range(W)]
    for step in range(int(g["steps"])):
        p = f"s{step}."
        feats = [g[f"r{r}.{p}feat"] for r in range(W)]
        labels = [g[f"r{r}.{p}label"] for r in range(W)]
        perms = [g[f"r{r}.{p}perm"] for r in range(W)]
        res = partial_fc.step(feats, labels, weights, C, kind, s, m, a, k, sr,
                              perms if int(sr) != 1 else None)
        for r in range(W):
            if int(sr) != 1:   # bit-exact: sampled class indices
                assert np.array_equal(res["index"][r], g[f"r{r}.{p}index"]), (name, step, r)
            close(res["x_grad"][r], g[f"r{r}.{p}x_grad"], 5e-4, 2e-5)
            close(res["w_grad"][r], g[f"r{r}.{p}w_grad"], 5e-4, 2e-5)
            assert abs(res["loss"] - float(g[f"r{r}.{p}loss"])) <= 1e-5 * abs(res["loss"])
        weights = [g[f"r{r}.{p}weight_after"] for r in range(W)]


@pytest.mark.parametrize("name", ["pfc_w1_full", "pfc_w1_d512"])
def test_torch_port_of_the_head_matches_reference(name):
    """oracle/partial_fc_torch.py (the timed CPU port of bench.py's reference arm) against the reference's own outputs."""
    import torch
    from oracle import partial_fc_torch as pt
    g = load_golden(name)
    assert int(g["W"]) == 1 and int(float(g["sample_rate"])) == 1
    s, m, a, k = (float(v) for v in g["smak"])
    assert a == 0.0 and k == 0.0
    w = torch.from_numpy(g["r0.w0"]).clone()
    mom = torch.zeros_like(w)
    for step in range(int(g["steps"])):
        p = f"r0.s{step}."
        x_grad, w_grad, loss = pt.head_step(torch.from_numpy(g[p + "feat"]), torch.from_numpy(g[p + "label"]), w, str(g["kind"]), s, m)
        close(x_grad.numpy(), g[p + "x_grad"], 1e-4, 1e-6)
        close(w_grad.numpy(), g[p + "w_grad"], 1e-4, 1e-6)
        assert abs(float(loss) - float(g[p + "loss"])) <= 1e-5 * abs(float(loss))
        pt.sgd_update(w, mom, w_grad)
        close(w.numpy(), g[p + "weight_after"], 1e-5, 1e-7)
        close(mom.numpy(), g[p + "mom_after"], 1e-5, 1e-7)


def test_shard_geometry():
    # ref partial_fc.py:34-36 at BASELINE config 3 (93,431 classes over 8 ranks)
    geo = [partial_fc.shard_geometry(93431, 8, r) for r in range(8)]
    assert [g[0] for g in geo] == [11679] * 7 + [11678]
    assert geo[7][1] == 11679 * 7 and sum(g[0] for g in geo) == 93431
    assert partial_fc.shard_geometry(1_000_000, 8, 3, 0.1) == (125000, 375000, 12500)


def test_sample_overflow_uses_positive_set():
    g = load_golden("pfc_w1_overflow")
    tl, index = partial_fc.sample(g["r0.s0.label"], g["r0.s0.perm"], 0, 40, 4, 0.1)
    assert np.array_equal(index, np.unique(g["r0.s0.label"]))
    assert np.array_equal(index, g["r0.s0.index"])


def test_model_cpu_matches_reference():
    """oracle.model_cpu (functional torch-CPU restatement) vs the reference MSML's own outputs."""
    import torch
    from oracle import model_cpu
    from oracle.detfill import det_labels, det_tensor, fill_state_dict_
    from msml_b200.backbones import MSML
    g = load_golden("model_iresnet18")
    net = MSML("iresnet18", "unet", (1, 1, 1, 1), 97, header_type="AMArcFace", fm_params=(3, 2, "sigmoid", "mul"))
    fill_state_dict_(net)           # same keys/shapes as the reference => same deterministic weights
    sd = model_cpu.trainable_state(net)
    x = det_tensor("model.x", (2, 3, 112, 112))
    with torch.no_grad():
        feat, seg = model_cpu.msml_forward(sd, x, "iresnet18", training=False)
    close(feat, g["eval_feature"], 1e-4, 1e-4)
    close(seg, g["eval_seg"], 1e-4, 1e-4)
    label = det_labels("model.l", 2, 97)
    feat, seg = model_cpu.msml_forward(sd, x, "iresnet18", training=True)
    cls = model_cpu.am_head(sd["classification.weight"], feat, label, "arc", 64.0, 0.5)
    close(cls.detach(), g["train_cls"], 1e-3, 1e-3)
    loss = torch.nn.functional.cross_entropy(cls, label) + seg.mean()
    assert abs(float(loss) - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    loss.backward()
    for key in ("frb.conv1.weight", "frb.layer4.1.bn3.weight", "osb.deconv5.weight"):
        close(sd[key].grad, g["grad." + key], 2e-3, 2e-3 * float(np.abs(g["grad." + key]).max()))


CONSENSUS = ["binary_missing", "four_blobs", "four_blobs_all_all", "four_blobs_idx_all", "underflow", "seg_shape"]


def consensus_case(g, name):
    alpha, beta, rp, rkl = [str(v) for v in g[name + ".cfg"]]
    return g[name + ".logit"], g[name + ".blobs"].astype(np.int64), g[name + ".target"].astype(np.int64), \
        (float(alpha), float(beta), rp, rkl)


@pytest.mark.parametrize("name", CONSENSUS)
def test_consensus_loss_matches_reference(name):
    """ref tricks/consensus_loss.py:63-178 run by make_golden.py (loss and autograd gradient)."""
    g = load_golden("consensus")
    logit, blobs, target, cfg = consensus_case(g, name)
    # the reference's `p != 0` pattern is that of an fp32 softmax; only the underflow case depends on it
    loss, dz = consensus.consensus_loss(logit, blobs, target, *cfg, softmax_dtype=np.float32 if name == "underflow" else np.float64)
    assert abs(loss - float(g[name + ".loss"])) <= 2e-6 * abs(float(g[name + ".loss"]))
    close(dz, g[name + ".dlogit"], 2e-4, 1e-6 * float(np.abs(g[name + ".dlogit"]).max()))


def test_consensus_loss_gradient_is_the_derivative():
    """Central differences of the oracle's own forward (fp64): pins the closed-form gradient independently of autograd."""
    rng = np.random.default_rng(5)
    logit = rng.normal(size=(2, 3, 4, 5))
    blobs = rng.integers(0, 3, size=(2, 4, 5))
    blobs[1][blobs[1] == 2] = 0                     # blob 2 is missing from sample 1
    target = np.array([2, 0, 1])[blobs]
    for cfg in ((10.0, 5.0, "idx", "idx"), (2.0, 3.0, "all", "all")):
        # 'all' on the first term has no defined gradient for a sample that lacks a blob (the reference yields NaN): merge blob 2 away
        blobs_c = np.where(blobs == 2, 0, blobs) if cfg[2] == "all" else blobs
        target_c = np.array([2, 0, 1])[blobs_c]
        _, dz = consensus.consensus_loss(logit, blobs_c, target_c, *cfg)
        num = np.zeros_like(logit)
        eps = 1e-6
        for i in np.ndindex(*logit.shape):
            zp = logit.copy(); zp[i] += eps
            zm = logit.copy(); zm[i] -= eps
            num[i] = (consensus.consensus_loss(zp, blobs_c, target_c, *cfg, want_grad=False)[0]
                      - consensus.consensus_loss(zm, blobs_c, target_c, *cfg, want_grad=False)[0]) / (2 * eps)
        close(dz, num, 1e-5, 1e-8)


@pytest.mark.parametrize("fill,lo,hi", [("black", 40, 41), ("white", 10, 31), ("gauss", 20, 61), ("black", 0, 1), ("black", 90, 91)])
def test_random_block_matches_reference(fill, lo, hi):
    """msml_b200.datasets.augment.RandomBlock (and its batched twin) against the reference's RandomBlock
    (ref datasets/augment/rand_occ.py:25-72) run on the same images with the same numpy seed: pixel for pixel."""
    from PIL import Image
    from msml_b200.datasets.augment import RandomBlock
    from msml_b200.datasets.augment.rand_occ import random_block_batch
    g = load_golden("rand_occ")
    want = g["%s_%d_%d" % (fill, lo, hi)]
    np.random.seed(1)
    t = RandomBlock(lo, hi, fill)
    got = np.stack([np.asarray(t(Image.fromarray(im))) for im in g["imgs"]])
    assert np.array_equal(got, want)
    np.random.seed(1)
    got_b = random_block_batch(np.ascontiguousarray(g["imgs"].transpose(0, 3, 1, 2)), lo, hi, fill)
    assert np.array_equal(got_b.transpose(0, 2, 3, 1), want)
    if lo > 0:
        assert (got != g["imgs"]).any()
    else:
        assert np.array_equal(got, g["imgs"])
