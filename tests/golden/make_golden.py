#!/usr/bin/env python
"""Generate the golden vectors in this directory by RUNNING THE REFERENCE (CPU, fp32).

Run in the build container only:   python tests/golden/make_golden.py
It imports /root/reference read-only (never copies it), runs the reference modules on
deterministic inputs (oracle/detfill.py) and writes small .npz fixtures next to this script.
The reference has no tests or golden vectors of its own, so these
files are what pins the oracle and, on the GPU box, the CUDA path.

Harness-only shims (NOT reference behaviour):
  * PartialFC.__init__ hard-codes cuda:{local_rank} and a CUDA stream (ref partial_fc.py:27,60),
    so the object is built with __new__ and the same attributes on CPU; torch.cuda.stream /
    current_stream are no-ops; gloo has no reduce_scatter so it is emulated with all_reduce.
  * the margin callable PartialFC expects does not exist in the ref; it is
    obtained by calling the ref AMArcFace/AMCosFace.forward with F.linear/F.normalize patched
    to pass the logits through, so ref headers/margin_losses.py:275-303 / :390-418 run verbatim.
  * torch.rand inside PartialFC.sample is recorded so the draw can be replayed.
"""
import contextlib
import os
import random
import sys
import types

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle.detfill import det_labels, det_tensor, fill_state_dict_  # noqa: E402

PEER_OFF = {"use_ori": False, "use_conv": False, "mask_trans": "conv", "use_decoder": False}


def seeds():
    # ref train.py:34-38
    random.seed(4)
    np.random.seed(1)
    torch.manual_seed(1)


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


# --------------------------------------------------------------------------- FM operator
FM_CASES = [  # name, C, H, B, act, arith, train
    ("fm_c64_sigmoid_mul", 64, 6, 2, "sigmoid", "mul", True),
    ("fm_c128_tanh_add", 128, 5, 2, "tanh", "add", True),
    ("fm_c64_sigmoid_div", 64, 4, 2, "sigmoid", "div", False),
    ("fm_c256_tanh_sub", 256, 3, 2, "tanh", "sub", False),
    ("fm_c512_sigmoid_mul", 512, 7, 1, "sigmoid", "mul", False),
]


def gen_fm():
    from backbones.fm.fmoperator import FMCnn, FMNone
    for name, C, H, B, act, arith, train in FM_CASES:
        fm = FMCnn(H, H, C, kernel_size=3, resblocks=2, activation=act, arith_strategy=arith,
                   peer_params=dict(PEER_OFF))
        fill_state_dict_(fm)
        fm.train(train)
        yf = det_tensor(name + ".yf", (B, C, H, H)).requires_grad_(True)
        yo = det_tensor(name + ".yo", (B, 18, H, H))
        dout = det_tensor(name + ".dout", (B, C, H, H))
        cap = {}

        def hook(_m, _i, o):
            o.retain_grad()
            cap["z"] = o

        def cat_hook(_m, inp):
            inp[0].retain_grad()
            cap["xcat"] = inp[0]

        h1 = fm.res_block.register_forward_hook(hook)
        h2 = fm.same_conv.register_forward_pre_hook(cat_hook)
        out, l2 = fm(yf, yo)
        assert l2 is None
        out.backward(dout)
        h1.remove(); h2.remove()
        z = cap["z"]
        dyf_direct = yf.grad - cap["xcat"].grad[:, :C]
        pg = {("pgrad." + k): p.grad for k, p in fm.named_parameters()
              if k in ("res_block.1.prelu3.weight", "res_block.0.bn1.weight")}
        pg["pgrad.same_conv.weight[:2]"] = fm.same_conv.weight.grad[:2]
        save(name, C=C, H=H, B=B, act=act, arith=arith, train=int(train),
             yf=yf, yo=yo, dout=dout, z=z, out=out, dyf_total=yf.grad, dyf_direct=dyf_direct,
             dz=z.grad, **pg)
    o, l = FMNone()(yf, yo)
    assert o is yf and l is None


# --------------------------------------------------------------------------- FM operator, peer-guided branch
FM_PEER_CASES = [  # name, C, H, B, act, arith, use_conv, mask_trans
    ("fm_peer_c64_conv", 64, 5, 2, "sigmoid", "mul", True, "conv"),
    ("fm_peer_c128_invert", 128, 4, 2, "tanh", "add", False, "invert"),
]


def gen_fm_peer():
    """FMCnn with use_ori=True (ref fmoperator.py:129-166, 293-302, 307-308): m_bar = conv_m(gate), f_out = conv1(m_bar*yf),
    l2 = MSE(conv2(m_bar*yt), f_out), out = arith(yf, gate) + f_out + yf.  Random (deterministic) weights: parity only."""
    from backbones.fm.fmoperator import FMCnn
    for name, C, H, B, act, arith, use_conv, mask_trans in FM_PEER_CASES:
        fm = FMCnn(H, H, C, kernel_size=3, resblocks=2, activation=act, arith_strategy=arith,
                   peer_params={"use_ori": True, "use_conv": use_conv, "mask_trans": mask_trans, "use_decoder": False})
        fill_state_dict_(fm)
        fm.train(True)
        yf = det_tensor(name + ".yf", (B, C, H, H)).requires_grad_(True)
        yo = det_tensor(name + ".yo", (B, 18, H, H))
        yt = det_tensor(name + ".yt", (B, C, H, H))
        dout = det_tensor(name + ".dout", (B, C, H, H))
        out, l2 = fm(yf, yo, yt)
        assert l2 is not None
        (out * dout).sum().add(0.5 * l2).backward()
        pg = {("pgrad." + k): p.grad for k, p in fm.named_parameters()
              if k in ("res_block.1.prelu3.weight", "res_block.0.bn1.weight", "conv_m.0.weight", "conv1.0.weight", "conv2.3.weight")
              and p.grad is not None}
        save(name, C=C, H=H, B=B, act=act, arith=arith, use_conv=int(use_conv), mask_trans=mask_trans,
             yf=yf, yo=yo, yt=yt, dout=dout, out=out, l2=l2, dyf=yf.grad, **pg)
        # eval-time call (yt is None): no distillation loss, f_out still added
        fm.eval()
        with torch.no_grad():
            out_e, l2_e = fm(yf.detach(), yo)
        assert l2_e is None
        save(name + "_eval", out=out_e)


# --------------------------------------------------------------------------- DAP
def gen_dap():
    from backbones.osb.unet import unet
    net = unet(backbone="r18", gray=False, input_size=112)
    x = det_tensor("dap.x", (3, 18, 10, 12)).requires_grad_(True)
    # a few exact ties / near ties for the argmax rule
    with torch.no_grad():
        x[0, :, 0, 0] = 0.25
        x[1, 9:, 3, 4] = x[1, :9, 3, 4]
    y = net.DAP(x)
    dy = det_tensor("dap.dy", tuple(y.shape))
    y.backward(dy)
    mask = torch.stack([y[b].max(0)[1] for b in range(y.shape[0])])  # ref train.py:357
    save("dap", x=x, y=y, dy=dy, dx=x.grad, mask=mask)


# --------------------------------------------------------------------------- segmentation criterion
def consensus_cases():
    """(name, logit, blobs, target, (alpha, beta, reduce_pixel, reduce_pixel_kl)); deterministic (numpy default_rng(0))."""
    rng = np.random.default_rng(0)
    cases = []
    N, C, H, W = 3, 2, 12, 12           # ref train.py:258 usage: blobs == target == binary occlusion mask; sample 2 lacks blob 1
    logit = (rng.normal(size=(N, C, H, W)) * 2).astype(np.float32)
    msk = (rng.random((N, H, W)) < 0.3).astype(np.int64)
    msk[2] = 0
    cases.append(("binary_missing", logit, msk, msk, (10.0, 5.0, "idx", "idx")))
    N, C, H, W = 2, 4, 5, 7             # connected-component ids that are not 0..K-1, labels differ from ids
    logit = rng.normal(size=(N, C, H, W)).astype(np.float32)
    ids = np.array([0, 3, 5, 9])
    bl = ids[rng.integers(0, 4, size=(N, H, W))]
    tg = np.vectorize({0: 1, 3: 0, 5: 3, 9: 1}.get)(bl)
    cases.append(("four_blobs", logit, bl, tg, (10.0, 5.0, "idx", "idx")))
    cases.append(("four_blobs_all_all", logit, bl, tg, (10.0, 5.0, "all", "all")))
    cases.append(("four_blobs_idx_all", logit, bl, tg, (3.0, 0.7, "idx", "all")))
    N, C, H, W = 2, 2, 6, 6             # fp32 softmax underflow: p == 0 entries leave the KL term (ref :153-158)
    logit = rng.normal(size=(N, C, H, W)).astype(np.float32)
    logit[0, 0, :2] += 120
    logit[1, 1, 3] -= 200
    msk = (rng.random((N, H, W)) < 0.5).astype(np.int64)
    cases.append(("underflow", logit, msk, msk, (10.0, 5.0, "idx", "idx")))
    N, C, H, W = 4, 2, 112, 112         # the shape of final_seg, (N, 1, H, W) masks
    logit = rng.normal(size=(N, C, H, W)).astype(np.float32)
    msk = np.zeros((N, 1, H, W), np.int64)
    for n in range(N):                  # one rectangular occluder per image, as RandomBlock draws them
        h0, w0 = rng.integers(0, 70, size=2)
        msk[n, 0, h0:h0 + 20 + 10 * n, w0:w0 + 30] = 1
    cases.append(("seg_shape", logit, msk, msk[:, 0], (10.0, 5.0, "idx", "idx")))
    return cases


def gen_consensus():
    import logging
    from tricks.consensus_loss import StructureConsensuLossFunction
    logging.getLogger("tricks.consensus_loss").setLevel(logging.WARNING)
    out = {}
    for name, logit, blobs, target, cfg in consensus_cases():
        crit = StructureConsensuLossFunction(*cfg)
        lt = torch.tensor(logit, requires_grad=True)
        loss = crit(lt, torch.tensor(blobs), torch.tensor(target))      # ref tricks/consensus_loss.py:177-178
        loss.backward()
        out[name + ".logit"] = logit
        out[name + ".blobs"] = blobs.astype(np.int16)
        out[name + ".target"] = target.astype(np.int16)
        out[name + ".cfg"] = np.array([str(c) for c in cfg])
        out[name + ".loss"] = np.float64(loss.item())
        out[name + ".dlogit"] = lt.grad.numpy()
    save("consensus", **out)


# --------------------------------------------------------------------------- margin heads
def gen_margins():
    from headers.margin_losses import AMArcFace, AMCosFace, Softmax
    B, D, C = 6, 16, 8
    label = torch.tensor([-1, 4, -1, 5, 3, -1])  # ref margin_losses.py:439 fixture
    label2 = det_labels("margin.l2", B, C)
    out = {}
    for kind, cls, (s, m, a, k) in [("arc", AMArcFace, (64.0, 0.5, 0.0, 0.0)),
                                    ("arc_am", AMArcFace, (1.0, 0.5, 1.2, 0.1)),
                                    ("cos", AMCosFace, (64.0, 0.4, 0.0, 0.0)),
                                    ("cos_am", AMCosFace, (32.0, 0.35, 1.2, 0.1))]:
        for ltag, lab in (("p", label), ("f", label2)):
            head = cls(D, C, None, s=s, m=m, a=a, k=k)
            with torch.no_grad():
                head.weight.copy_(det_tensor(f"margin.{kind}.w", (C, D)))
            e = det_tensor(f"margin.{kind}.e", (B, D)).requires_grad_(True)
            dl = det_tensor(f"margin.{kind}.dl", (B, C), 0.1)
            logits = head(e, lab)
            logits.backward(dl)
            tag = f"{kind}_{ltag}"
            out.update({tag + ".e": e, tag + ".w": head.weight, tag + ".label": lab,
                        tag + ".smak": np.array([s, m, a, k]), tag + ".logits": logits,
                        tag + ".dl": dl, tag + ".de": e.grad, tag + ".dw": head.weight.grad})
    sm = Softmax(D, C, None)
    with torch.no_grad():
        sm.weight.copy_(det_tensor("margin.sm.w", (C, D)))
        sm.bias.copy_(det_tensor("margin.sm.b", (C,)))
    e = det_tensor("margin.sm.e", (B, D))
    out.update({"softmax.e": e, "softmax.w": sm.weight, "softmax.b": sm.bias,
                "softmax.logits": sm(e, label)})
    try:
        Softmax(D, C, [0])(e, label)
        raise AssertionError("expected ValueError")
    except ValueError:
        pass
    save("margins", **out)


# --------------------------------------------------------------------------- PartialFC harness
def ref_margin_callable(kind, s, m, a, k):
    """Runs ref margin_losses.py:390-418 (arc) / :275-303 (cos) verbatim on given logits."""
    import headers.margin_losses as ml
    cls = ml.AMArcFace if kind == "arc" else ml.AMCosFace
    head = cls(4, 4, None, s=s, m=m, a=a, k=k)

    def margin_softmax(logits, label):
        real_F = ml.F
        ml.F = types.SimpleNamespace(linear=lambda _a, _b: logits, normalize=lambda t: t)
        try:
            return head.forward(None, label)
        finally:
            ml.F = real_F
    return margin_softmax


class _NoStream:
    def wait_stream(self, _s):
        pass


def build_ref_pfc(rank, world_size, batch_size, num_classes, sample_rate, margin, weight, D):
    from headers.partial_fc import PartialFC
    from torch.nn.parameter import Parameter
    pfc = PartialFC.__new__(PartialFC)
    torch.nn.Module.__init__(pfc)
    pfc.num_classes, pfc.rank, pfc.local_rank = num_classes, rank, rank
    pfc.device = torch.device("cpu")
    pfc.world_size, pfc.batch_size = world_size, batch_size
    pfc.margin_softmax, pfc.sample_rate, pfc.embedding_size = margin, sample_rate, D
    pfc.prefix = "./"
    pfc.num_local = num_classes // world_size + int(rank < num_classes % world_size)
    pfc.class_start = num_classes // world_size * rank + min(rank, num_classes % world_size)
    pfc.num_sample = int(sample_rate * pfc.num_local)
    pfc.weight = weight.clone()
    pfc.weight_mom = torch.zeros_like(pfc.weight)
    pfc.stream = None
    pfc.index = None
    if int(sample_rate) == 1:
        pfc.update = lambda: 0
        pfc.sub_weight = Parameter(pfc.weight)
        pfc.sub_weight_mom = pfc.weight_mom
    else:
        pfc.sub_weight = Parameter(torch.empty((0, 0)))
    return pfc


@contextlib.contextmanager
def cpu_shims(record):
    real_stream, real_cur, real_rs, real_rand = (torch.cuda.stream, torch.cuda.current_stream,
                                                 dist.reduce_scatter, torch.rand)
    torch.cuda.stream = lambda _s: contextlib.nullcontext()
    torch.cuda.current_stream = lambda *a, **k: _NoStream()

    def reduce_scatter(out, in_list, *a, **k):
        full = torch.cat([t.detach() for t in in_list], 0).clone()
        dist.all_reduce(full)
        r = dist.get_rank()
        n = out.shape[0]
        with torch.no_grad():
            out.copy_(full[r * n:(r + 1) * n])

    def rand(*a, **k):
        k.pop("device", None)
        t = real_rand(*a, **k)
        record.append(t.clone())
        return t

    dist.reduce_scatter = reduce_scatter
    torch.rand = rand
    try:
        yield
    finally:
        torch.cuda.stream, torch.cuda.current_stream = real_stream, real_cur
        dist.reduce_scatter, torch.rand = real_rs, real_rand


PFC_CASES = [  # name, W, B, C, D, sample_rate, kind, (s,m,a,k), steps
    ("pfc_w1_full", 1, 8, 37, 64, 1.0, "arc", (64.0, 0.5, 0.0, 0.0), 2),
    ("pfc_w1_sample", 1, 8, 101, 64, 0.3, "arc", (64.0, 0.5, 0.0, 0.0), 2),
    ("pfc_w2_full", 2, 4, 37, 64, 1.0, "cos", (64.0, 0.4, 0.0, 0.0), 2),
    ("pfc_w2_sample", 2, 6, 203, 64, 0.25, "arc", (64.0, 0.5, 0.0, 0.0), 2),
    ("pfc_w2_am", 2, 4, 50, 64, 0.5, "arc", (32.0, 0.45, 1.2, 0.1), 1),
    ("pfc_w1_d512", 1, 16, 96, 512, 1.0, "arc", (64.0, 0.5, 0.0, 0.0), 1),
    ("pfc_w1_overflow", 1, 16, 40, 64, 0.1, "arc", (64.0, 0.5, 0.0, 0.0), 1),  # n_pos > num_sample
]


def pfc_worker(rank, world_size, case, port, q):
    name, W, B, C, D, sr, kind, smak, steps = case
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    torch.set_num_threads(1)
    seeds()
    margin = ref_margin_callable(kind, *smak)
    num_local = C // W + int(rank < C % W)
    w0 = det_tensor(f"{name}.w{rank}", (num_local, D), 0.01)  # N(0, 0.01) as ref :56
    pfc = build_ref_pfc(rank, W, B, C, sr, margin, w0, D)
    opt = torch.optim.SGD([{"params": pfc.parameters()}], lr=0.1, momentum=0.9, weight_decay=5e-4)
    res = {"w0": w0.numpy()}
    for step in range(steps):
        feat = torch.nn.functional.normalize(det_tensor(f"{name}.x{rank}.{step}", (B, D)))
        label = det_labels(f"{name}.l{rank}.{step}", B, C)
        if name == "pfc_w1_overflow":
            label = torch.arange(B, dtype=torch.int64) * 2
        rec = []
        with cpu_shims(rec):
            x_grad, loss_v = pfc.forward_backward(label, feat, opt)
        p = f"s{step}."
        res.update({p + "feat": feat.numpy(), p + "label": label.numpy(),
                    p + "x_grad": x_grad.detach().numpy(), p + "loss": loss_v.detach().numpy(),
                    p + "w_grad": pfc.sub_weight.grad.detach().numpy().copy(),
                    p + "perm": rec[0].numpy() if rec else np.zeros(0, np.float32),
                    p + "index": pfc.index.numpy().copy() if pfc.index is not None else np.zeros(0, np.int64)})
        opt.step()
        pfc.update()
        opt.zero_grad()
        res[p + "weight_after"] = pfc.weight.detach().numpy().copy()
        res[p + "mom_after"] = pfc.weight_mom.detach().numpy().copy()
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def gen_pfc():
    port = 29611
    for case in PFC_CASES:
        name, W = case[0], case[1]
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=pfc_worker, args=(r, W, case, port, q)) for r in range(W)]
        port += 1
        for p in procs:
            p.start()
        got = dict(q.get() for _ in range(W))
        for p in procs:
            p.join()
        out = {"W": W, "B": case[2], "C": case[3], "D": case[4], "sample_rate": case[5],
               "kind": case[6], "smak": np.array(case[7]), "steps": case[8]}
        for r in range(W):
            for k, v in got[r].items():
                out[f"r{r}.{k}"] = v
        save(name, **out)


# --------------------------------------------------------------------------- full model
def gen_model():
    import backbones
    for frb in ("iresnet18",):
        net = backbones.MSML(frb, "unet", (1, 1, 1, 1), 97, fp16=False, header_type="AMArcFace",
                             header_params=(64.0, 0.5, 0.0, 0.0), fm_params=(3, 2, "sigmoid", "mul"),
                             peer_params=dict(PEER_OFF))
        fill_state_dict_(net)
        x = det_tensor("model.x", (2, 3, 112, 112))
        label = det_labels("model.l", 2, 97)
        net.eval()
        with torch.no_grad():
            feat, seg = net(x)
        net.train()
        final_cls, final_seg, kd = net(x, label)
        loss = torch.nn.functional.cross_entropy(final_cls, label) + final_seg.mean()
        loss.backward()
        named = dict(net.named_parameters())
        keys = ["frb.conv1.weight", "frb.fm_ops.0.same_conv.weight", "frb.fm_ops.3.res_block.1.conv3.weight",
                "frb.layer4.1.bn3.weight", "osb.deconv5.weight", "classification.weight", "frb.fc.bias"]
        gn = {"gradnorm." + k: named[k].grad.norm() for k in named if named[k].grad is not None}
        save(f"model_{frb}", eval_feature=feat, eval_seg=seg, train_cls=final_cls,
             train_seg=final_seg, kd=np.float32(kd), loss=loss,
             **{"grad." + k: named[k].grad for k in keys if named[k].numel() < 200000},
             **gn)


def gen_rand_occ():
    """ref datasets/augment/rand_occ.py:25-72 RandomBlock on deterministic RGB images, numpy seeded as ref train.py:36.
    The module is loaded by path (the venv's HuggingFace `datasets` shadows the reference's namespace package) with a stub
    for its `eval.preprocess.RealOcc.image_infer` import (unused by RandomBlock) — harness-only, SURVEY 8c."""
    import importlib.util
    from PIL import Image
    stub = types.ModuleType("eval.preprocess.RealOcc.image_infer")
    stub.RealOcc = object
    for name in ("eval", "eval.preprocess", "eval.preprocess.RealOcc"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["eval.preprocess.RealOcc.image_infer"] = stub
    spec = importlib.util.spec_from_file_location("ref_rand_occ", os.path.join(REF, "datasets", "augment", "rand_occ.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.RandomState(7)
    imgs = rng.randint(0, 256, (6, 112, 112, 3)).astype(np.uint8)
    out = {"imgs": imgs}
    for fill, lo, hi in (("black", 40, 41), ("white", 10, 31), ("gauss", 20, 61), ("black", 0, 1), ("black", 90, 91)):
        np.random.seed(1)
        t = mod.RandomBlock(lo, hi, fill)
        res = np.stack([np.asarray(t(Image.fromarray(im))) for im in imgs])
        out["%s_%d_%d" % (fill, lo, hi)] = res
    save("rand_occ", **out)


def gen_model_peer():
    """MSML with the peer-guided branch on (ref config.yaml:22-26 default: use_ori / use_conv / mask_trans conv): the frozen
    teacher of ref backbones/peer/arcface.py with RANDOM deterministic weights (the reference ships none: its
    arcface18() would raise FileNotFoundError, so the constructor is called with pretrained=False — harness-only)."""
    import backbones
    import backbones.peer as peer_mod
    real = peer_mod.arcface18
    peer_mod.arcface18 = lambda *a, **k: real(pretrained=False)
    try:
        net = backbones.MSML("iresnet18", "unet", (1, 1, 1, 1), 97, fp16=False, header_type="AMArcFace",
                             header_params=(64.0, 0.5, 0.0, 0.0), fm_params=(3, 2, "sigmoid", "mul"),
                             peer_params={"use_ori": True, "use_conv": True, "mask_trans": "conv", "use_decoder": False})
    finally:
        peer_mod.arcface18 = real
    fill_state_dict_(net)
    assert not any(p.requires_grad for p in net.frb.peer.parameters())
    x = det_tensor("model.x", (2, 3, 112, 112))
    ori = det_tensor("model.ori", (2, 3, 112, 112))
    label = det_labels("model.l", 2, 97)
    net.eval()
    with torch.no_grad():
        feat, seg = net(x)                                  # eval: no ori, the distillation branch is inactive
        pf, inter = net.frb.peer(ori)
    net.train()
    final_cls, final_seg, kd = net(x, label, ori)
    loss = torch.nn.functional.cross_entropy(final_cls, label) + final_seg.mean()
    loss.backward()
    named = dict(net.named_parameters())
    keys = ["frb.conv1.weight", "frb.fm_ops.0.same_conv.weight", "frb.fm_ops.0.conv_m.0.weight", "frb.fm_ops.1.conv1.0.weight",
            "frb.fm_ops.3.conv2.3.weight", "frb.layer4.1.bn3.weight", "classification.weight"]
    gn = {"gradnorm." + k: named[k].grad.norm() for k in named if named[k].grad is not None}
    save("model_iresnet18_peer", eval_feature=feat, eval_seg=seg, peer_feature=pf, peer_ft3=inter[3], peer_ft0_norm=inter[0].norm(),
         train_cls=final_cls, train_seg=final_seg, kd=np.float32(kd.item()), loss=loss,
         **{"grad." + k: named[k].grad for k in keys if named[k].numel() < 200000}, **gn)


if __name__ == "__main__":
    which = sys.argv[1:] or ["fm", "fm_peer", "dap", "margins", "pfc", "model", "model_peer", "consensus", "rand_occ"]
    seeds()
    for w in which:
        globals()["gen_" + w]()
