"""TrainStep: the CUDA-graph replay must compute the same step as the eager path, and the eager
path must match the oracle's CPU restatement of the reference step."""
import copy

import numpy as np
import pytest
import torch

from gpu_util import assert_close, host, need_gpu

pytestmark = pytest.mark.gpu


def _build(seed=3, classes=1000, B=8, fp16=False, fused=None):
    from msml_b200.backbones import MSML
    from msml_b200.headers import ArcFace, PartialFC
    from oracle.detfill import fill_state_dict_
    torch.manual_seed(seed)
    net = MSML("iresnet18", "unet", (1, 1, 1, 1), classes, fp16=fp16, header_type=None, fm_params=(3, 2, "sigmoid", "mul"))
    fill_state_dict_(net)
    net = net.cuda().train()
    pfc = PartialFC(0, 0, 1, B, False, ArcFace(64.0, 0.5), classes)
    opt = torch.optim.SGD([p for p in net.parameters() if p.requires_grad], lr=0.01, momentum=0.9, weight_decay=5e-4, fused=fused)
    opt_pfc = torch.optim.SGD([{"params": pfc.parameters()}], lr=0.01, momentum=0.9, weight_decay=5e-4, fused=fused)
    return net, pfc, opt, opt_pfc


def test_graph_replay_matches_eager():
    need_gpu()
    from msml_b200.engine import TrainStep
    B = 8
    g = torch.Generator(device="cuda").manual_seed(5)
    imgs = [torch.randn(B, 3, 112, 112, device="cuda", generator=g) for _ in range(3)]
    labels = [torch.randint(0, 1000, (B,), device="cuda", generator=g) for _ in range(3)]
    losses, weights = {}, {}
    for mode in ("eager", "graph"):
        net, pfc, opt, opt_pfc = _build()
        step = TrainStep(net, pfc, opt, opt_pfc, (B, 3, 112, 112), use_graph=(mode == "graph"))
        if mode == "graph":
            before = copy.deepcopy(net.state_dict())
            step.recapture()                          # warm-up + capture must leave model / optimizer state untouched
            for k, v in net.state_dict().items():
                assert torch.equal(v, before[k]), k
        losses[mode] = [float(step(i, l)) for i, l in zip(imgs, labels)]
        weights[mode] = (host(net.frb.conv1.weight), host(net.frb.fm_ops[2].same_conv.weight), host(pfc.weight))
        assert net.osb.conv1.weight.grad is None       # unused branch keeps grad None => optimizer skips it (ref semantics)
    for a, b in zip(losses["eager"], losses["graph"]):
        assert abs(a - b) <= 2e-3 * abs(a), (losses["eager"], losses["graph"])
    # same kernels, same order; only the fp32 atomics of the head (dX split-K, rdot) reorder between runs, and three
    # SGD steps at s=64 amplify that slightly: compare in norm
    for a, b in zip(weights["eager"], weights["graph"]):
        assert np.linalg.norm(a - b) <= 2e-3 * np.linalg.norm(a), np.linalg.norm(a - b) / np.linalg.norm(a)


def test_eager_step_matches_cpu_oracle_step():
    """One full step (ires18, 1000 classes, fp32 backbone) against oracle.model_cpu + oracle.partial_fc."""
    need_gpu()
    from msml_b200.engine import TrainStep
    from oracle import model_cpu, partial_fc as opfc
    B = 8
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        net, pfc, opt, opt_pfc = _build()
        sd = model_cpu.trainable_state(net)
        w0 = host(pfc.weight)
        g = torch.Generator().manual_seed(9)
        img = torch.randn(B, 3, 112, 112, generator=g)
        label = torch.randint(0, 1000, (B,), generator=g)
        step = TrainStep(net, pfc, opt, opt_pfc, (B, 3, 112, 112), use_graph=False, max_norm=None)
        loss = float(step(img.cuda(), label.cuda()))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    feat, _ = model_cpu.msml_forward(sd, img, "iresnet18", training=True)
    featn = torch.nn.functional.normalize(feat)
    res = opfc.step([featn.detach().to(torch.bfloat16).double().numpy()], [label.numpy()], [w0], 1000, "arc", 64.0, 0.5)
    assert abs(loss - res["loss"]) <= 2e-3 * abs(res["loss"]), (loss, res["loss"])
    featn.backward(torch.from_numpy(res["x_grad"][0]).float())
    for key in ("frb.conv1.weight", "frb.fm_ops.0.same_conv.weight", "frb.layer3.0.conv1.weight", "frb.fc.weight"):
        want = sd[key].grad.numpy()
        got_w = host(dict(net.named_parameters())[key])           # w1 = w0 - lr * (g + wd * w0)
        w_init = sd[key].detach().numpy()
        got_grad = (w_init - got_w) / 0.01 - 5e-4 * w_init
        assert_close(got_grad, want, 5e-2, atol_frac=3e-2, what="grad via SGD update " + key)


def test_bf16_engine_step_matches_plain_autograd_step():
    """bf16 autocast path: TrainStep (bf16 shadow weights, direct gradient accumulation into the flat buffer, fused
    multi-tensor weight-gradient add, zero-padded channel counts) against the same model stepped the plain PyTorch way
    (autocast casts, AccumulateGrad, clip_grad_norm_)."""
    need_gpu()
    from msml_b200.engine import TrainStep
    B = 8
    g = torch.Generator(device="cuda").manual_seed(11)
    imgs = [torch.randn(B, 3, 112, 112, device="cuda", generator=g) for _ in range(2)]
    labels = [torch.randint(0, 1000, (B,), device="cuda", generator=g) for _ in range(2)]

    net, pfc, opt, opt_pfc = _build(fp16=True, fused=True)      # fused SGD: the clip coefficient rides in its grad_scale
    step = TrainStep(net, pfc, opt, opt_pfc, (B, 3, 112, 112), use_graph=False)
    assert step._fused_sgd_takes_scale()
    loss_e = [float(step(i, l)) for i, l in zip(imgs, labels)]
    assert hasattr(net.frb.layer1[0].conv1.weight, "_msml_shadow")          # the engine path really ran on shadows

    net2, pfc2, opt2, opt_pfc2 = _build(fp16=True)
    loss_p = []
    for img, lab in zip(imgs, labels):
        feat, _ = net2(img)
        featn = torch.nn.functional.normalize(feat)
        x_grad, loss = pfc2.forward_backward(lab, featn, opt_pfc2)
        featn.backward(x_grad)
        torch.nn.utils.clip_grad_norm_([p for p in net2.parameters() if p.grad is not None], 5.0)
        opt2.step(); opt_pfc2.step(); pfc2.update()
        opt2.zero_grad(set_to_none=True)
        pfc2.sub_weight.grad = None
        loss_p.append(float(loss))
    for a, b in zip(loss_e, loss_p):
        assert abs(a - b) <= 2e-2 * abs(b), (loss_e, loss_p)
    for key in ("frb.conv1.weight", "frb.layer2.0.conv2.weight", "frb.fm_ops.1.same_conv.weight", "frb.layer3.1.bn2.weight",
                "frb.layer4.0.prelu.weight", "frb.fc.weight"):
        a = host(dict(net.named_parameters())[key]); b = host(dict(net2.named_parameters())[key])
        assert np.linalg.norm(a - b) <= 2e-2 * np.linalg.norm(b) + 1e-6, (key, np.linalg.norm(a - b) / np.linalg.norm(b))
    assert np.linalg.norm(host(pfc.weight) - host(pfc2.weight)) <= 2e-2 * np.linalg.norm(host(pfc2.weight))


def test_bf16_graph_replay_matches_eager():
    need_gpu()
    from msml_b200.engine import TrainStep
    B = 8
    g = torch.Generator(device="cuda").manual_seed(13)
    imgs = [torch.randn(B, 3, 112, 112, device="cuda", generator=g) for _ in range(3)]
    labels = [torch.randint(0, 1000, (B,), device="cuda", generator=g) for _ in range(3)]
    out = {}
    for mode in ("eager", "graph"):
        net, pfc, opt, opt_pfc = _build(fp16=True)
        step = TrainStep(net, pfc, opt, opt_pfc, (B, 3, 112, 112), use_graph=(mode == "graph"))
        out[mode] = ([float(step(i, l)) for i, l in zip(imgs, labels)], host(net.frb.layer3[0].conv1.weight), host(pfc.weight))
    for a, b in zip(out["eager"][0], out["graph"][0]):
        assert abs(a - b) <= 1e-2 * abs(a), (out["eager"][0], out["graph"][0])
    for k in (1, 2):
        a, b = out["eager"][k], out["graph"][k]
        assert np.linalg.norm(a - b) <= 1e-2 * np.linalg.norm(a)


def test_bucketed_allreduce_covers_every_gradient_exactly_once(monkeypatch):
    """world_size 2 emulated on one GPU: all_reduce is replaced by `x *= 2` (two identical ranks).  After the engine's
    division by the world size the step must equal the world_size-1 step: a gradient slice that no bucket covered would
    come out halved, one covered twice doubled.  Also checks that the stage buckets really start during backward."""
    need_gpu()
    import msml_b200.engine as eng
    B = 8
    g = torch.Generator(device="cuda").manual_seed(17)
    imgs = [torch.randn(B, 3, 112, 112, device="cuda", generator=g) for _ in range(2)]
    labels = [torch.randint(0, 1000, (B,), device="cuda", generator=g) for _ in range(2)]
    calls = []

    class _Done:
        def wait(self):
            return True

    def fake_all_reduce(t, op=None, group=None, async_op=False):
        calls.append(t.numel())
        t.mul_(2.0)
        return _Done()

    out = {}
    for world in (1, 2):
        net, pfc, opt, opt_pfc = _build(fp16=True)
        step = eng.TrainStep(net, pfc, opt, opt_pfc, (B, 3, 112, 112), world_size=world, use_graph=False)
        if world == 2:
            monkeypatch.setattr(eng.dist, "all_reduce", fake_all_reduce)
        losses = [float(step(i, l)) for i, l in zip(imgs, labels)]
        out[world] = (losses, {k: host(v) for k, v in net.named_parameters() if v.grad is not None}, step)
    step2 = out[2][2]
    assert sum(calls) == 2 * step2.flat.numel()                      # every element reduced once per step, two steps
    assert len(calls) >= 2 * 4                                       # >= 4 non-empty buckets per step (stages 3..0 [+ rest])
    for a, b in zip(out[1][0], out[2][0]):
        assert abs(a - b) <= 1e-3 * abs(a), (out[1][0], out[2][0])
    for k, a in out[1][1].items():
        b = out[2][1][k]
        assert np.linalg.norm(a - b) <= 2e-3 * np.linalg.norm(a) + 1e-7, (k, np.linalg.norm(a - b), np.linalg.norm(a))


@pytest.mark.parametrize("use_graph", [False, True])
def test_prefetched_inputs_give_the_same_step(use_graph):
    need_gpu()
    from msml_b200.engine import TrainStep
    B = 8
    g = torch.Generator().manual_seed(19)
    imgs = [torch.randn(B, 3, 112, 112, generator=g).pin_memory() for _ in range(3)]
    labels = [torch.randint(0, 1000, (B,), generator=g).pin_memory() for _ in range(3)]
    out = {}
    for mode in ("direct", "prefetch"):
        net, pfc, opt, opt_pfc = _build(fp16=True)
        step = TrainStep(net, pfc, opt, opt_pfc, (B, 3, 112, 112), use_graph=use_graph)
        losses = []
        if mode == "direct":
            for i, l in zip(imgs, labels):
                losses.append(float(step(i, l)))
        else:
            step.prefetch(imgs[0], labels[0])
            for k in range(3):
                loss_t = step()
                if k + 1 < 3:
                    step.prefetch(imgs[k + 1], labels[k + 1])
                losses.append(float(loss_t))
            with pytest.raises(RuntimeError):
                step()                                       # nothing staged
        out[mode] = losses
    for a, b in zip(out["direct"], out["prefetch"]):
        assert abs(a - b) <= 1e-2 * abs(a), out


def test_lr_schedule_is_honoured_by_the_captured_graph():
    """The captured step reads the learning rate from a device tensor: a torch LR scheduler that sets it to zero must
    freeze the weights on the next replay, and restoring it must move them again (ref train.py steps its LambdaLR
    every iteration)."""
    need_gpu()
    from msml_b200.engine import TrainStep
    B = 8
    g = torch.Generator(device="cuda").manual_seed(23)
    img = torch.randn(B, 3, 112, 112, device="cuda", generator=g)
    label = torch.randint(0, 1000, (B,), device="cuda", generator=g)
    net, pfc, opt, opt_pfc = _build(fp16=True, fused=True)
    step = TrainStep(net, pfc, opt, opt_pfc, (B, 3, 112, 112), use_graph=True)
    step(img, label)                                   # captures; lr is now a device tensor
    assert isinstance(opt.param_groups[0]["lr"], torch.Tensor) and opt.param_groups[0]["lr"].is_cuda
    factors = [1.0, 0.0, 1.0]
    sch = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: factors[min(s, 2)])
    sch_pfc = torch.optim.lr_scheduler.LambdaLR(opt_pfc, lambda s: factors[min(s, 2)])
    w = lambda: (net.frb.layer2[0].conv1.weight.detach().clone(), pfc.weight.detach().clone())
    w0 = w()
    step(img, label)
    w1 = w()
    assert not torch.equal(w0[0], w1[0]) and not torch.equal(w0[1], w1[1])
    sch.step(); sch_pfc.step()                         # factor 0: lr tensors filled with 0 in place
    assert float(opt.param_groups[0]["lr"]) == 0.0
    for st in opt.state.values():                      # momentum would still move the weights: clear it for the check
        if st.get("momentum_buffer") is not None:
            st["momentum_buffer"].zero_()
    pfc.weight_mom.zero_()
    step(img, label)
    w2 = w()
    assert torch.equal(w1[0], w2[0]) and torch.equal(w1[1], w2[1])
    sch.step(); sch_pfc.step()                         # factor 1 again
    step(img, label)
    w3 = w()
    assert not torch.equal(w2[0], w3[0]) and not torch.equal(w2[1], w3[1])


@pytest.mark.parametrize("fused_head", [False, True])
def test_captured_step_with_a_sampled_head(fused_head):
    """sample_rate < 1 inside the CUDA graph: torch.rand through the graph-registered generator, radix
    select and the row gathers into fixed-capacity buffers are all captured.  Every replay must draw a NEW sample that
    contains the positives of ITS batch (ref partial_fc.py:84-88), train (loss finite, class centres of sampled rows move,
    rows outside every sample stay bit-identical) — with the stock optimizer + update() and with headers.PartialFCSGD."""
    need_gpu()
    from msml_b200.backbones import MSML
    from msml_b200.engine import TrainStep
    from msml_b200.headers import ArcFace, PartialFC, PartialFCSGD
    from oracle.detfill import fill_state_dict_
    torch.manual_seed(3)
    B, C = 8, 1000
    net = MSML("iresnet18", "unet", (1, 1, 1, 1), C, fp16=False, header_type=None, fm_params=(3, 2, "sigmoid", "mul"))
    fill_state_dict_(net)
    net = net.cuda().train()
    pfc = PartialFC(0, 0, 1, B, False, ArcFace(64.0, 0.5), C, sample_rate=0.2)
    opt = torch.optim.SGD([p for p in net.parameters() if p.requires_grad], lr=0.01, momentum=0.9, weight_decay=5e-4, fused=True)
    hp = dict(lr=0.01, momentum=0.9, weight_decay=5e-4)
    opt_pfc = PartialFCSGD(pfc, **hp) if fused_head else torch.optim.SGD([{"params": pfc.parameters()}], fused=True, **hp)
    step = TrainStep(net, pfc, opt, opt_pfc, (B, 3, 112, 112), use_graph=True)
    step.recapture()
    w0 = pfc.weight.clone()
    g = torch.Generator(device="cuda").manual_seed(5)
    touched = torch.zeros(C, dtype=torch.bool, device="cuda")
    seen = []
    for _ in range(3):
        img = torch.randn(B, 3, 112, 112, device="cuda", generator=g)
        label = torch.randint(0, C, (B,), device="cuda", generator=g)
        loss = step(img, label)
        torch.cuda.synchronize()
        assert torch.isfinite(loss)
        idx = pfc.index.clone()
        assert idx.numel() == pfc.num_sample == 200 and bool((idx[1:] > idx[:-1]).all())
        assert bool(torch.isin(label, idx).all())                        # this batch's positives are in this replay's sample
        touched[idx] = True
        seen.append(idx)
    assert not torch.equal(seen[0], seen[1]) and not torch.equal(seen[1], seen[2])     # a new draw every replay
    moved = (pfc.weight != w0).any(dim=1)
    assert bool(moved[touched].all()) and not bool(moved[~touched].any())
    with pytest.raises(ValueError):                                                    # gathered batch > num_sample: refused up front
        TrainStep(net, PartialFC(0, 0, 1, B, False, ArcFace(64.0, 0.5), 70, sample_rate=0.1), opt, opt_pfc, (B, 3, 112, 112), use_graph=True)


# ------------------------------------------------------------------------------------------------ engine.FlatSGD
def _toy(seed):
    torch.manual_seed(seed)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1, bias=False), torch.nn.BatchNorm2d(8), torch.nn.PReLU(8),
                              torch.nn.Conv2d(8, 5, 3, padding=1, bias=True), torch.nn.Flatten(), torch.nn.Linear(5 * 6 * 6, 7))
    return net.cuda().to(memory_format=torch.channels_last)


def _bind(net, opt):
    """What engine.TrainStep does: one flat fp32 gradient buffer, every p.grad a view padded to 4 elements."""
    params = list(net.parameters())
    pad4 = lambda k: (k + 3) // 4 * 4
    flat = torch.zeros(sum(pad4(p.numel()) for p in params), device="cuda")
    off = 0
    for p in params:
        p.grad = torch.as_strided(flat, p.size(), p.stride(), off)
        off += pad4(p.numel())
    opt.bind_flat(params, flat)
    return flat


@pytest.mark.parametrize("nesterov,scale", [(False, None), (True, 2.5), (False, 0.5)])
def test_flat_sgd_matches_torch_sgd(nesterov, scale):
    """engine.FlatSGD (one msml_sgd_flat launch over flat parameter / momentum / gradient buffers) against
    torch.optim.SGD (ref train.py:186-191, 299) on the same gradients for four steps: parameters, momentum buffers and the
    emitted bf16 shadow weights; grad_scale as torch's fused optimizers take it."""
    need_gpu()
    from msml_b200.engine import FlatSGD
    ref_net, net = _toy(5), _toy(5)
    hp = dict(lr=0.05, momentum=0.9, weight_decay=5e-4, nesterov=nesterov)
    ref_opt = torch.optim.SGD(ref_net.parameters(), **hp)
    opt = FlatSGD(net.parameters(), **hp)
    with pytest.raises(RuntimeError, match="not bound"):
        opt.step()
    flat = _bind(net, opt)
    shadows = [opt.shadow_view(p) for p in net.parameters()]
    opt.refresh_shadows()
    for p, sh, q in zip(net.parameters(), shadows, ref_net.parameters()):
        assert torch.equal(p, q) and p.stride() == q.stride()                       # the move into the flat buffer kept values and layout
        assert torch.equal(sh, p.to(torch.bfloat16)) and sh.stride() == p.stride()
    if scale is not None:
        opt.grad_scale = torch.tensor(scale, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    for it in range(4):
        x = torch.randn(4, 3, 6, 6, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
        ref_opt.zero_grad(set_to_none=True)
        opt.zero_grad()
        assert float(flat.abs().sum()) == 0.0 and all(p.grad is not None for p in net.parameters())
        ref_net(x).square().mean().backward()
        if it == 0:
            net(x).square().mean().backward()                                     # autograd accumulates into the flat views
            for q, p in zip(ref_net.parameters(), net.parameters()):
                assert_close(host(p.grad), host(q.grad), 1e-4, atol_frac=1e-5, what="gradient through the flat view")
        for q, p in zip(ref_net.parameters(), net.parameters()):                  # both optimizers see the SAME gradients, so the
            p.grad.copy_(q.grad)                                                  # comparison is of the update rule alone
        if scale is not None:
            for q in ref_net.parameters():
                q.grad.div_(scale)
        ref_opt.step()
        opt.step()
        for (name, q), p, sh in zip(ref_net.named_parameters(), net.parameters(), shadows):
            assert_close(host(p), host(q), 2e-5, atol=2e-6, what="%s step %d" % (name, it))
            assert_close(host(opt.state[p]["momentum_buffer"]), host(ref_opt.state[q]["momentum_buffer"]), 2e-5, atol=2e-6, what="momentum " + name)
            assert torch.equal(sh, p.to(torch.bfloat16)), name
    # state_dict round trip keeps the momentum inside the flat buffer
    sd = copy.deepcopy(opt.state_dict())
    for p in net.parameters():
        opt.state[p]["momentum_buffer"].zero_()
    opt.load_state_dict(sd)
    for q, p in zip(ref_net.parameters(), net.parameters()):
        mb = opt.state[p]["momentum_buffer"]
        assert opt._m.data_ptr() <= mb.data_ptr() < opt._m.data_ptr() + 4 * opt._m.numel()
        assert_close(host(mb), host(ref_opt.state[q]["momentum_buffer"]), 2e-5, atol=2e-6, what="reloaded momentum")


@pytest.mark.parametrize("use_graph", [False, True])
def test_train_step_with_flat_sgd_matches_torch_fused_sgd(use_graph):
    """The whole bf16 step with engine.FlatSGD as the backbone optimizer == the same step with torch.optim.SGD(fused=True):
    losses of four steps and the weights afterwards (the clip coefficient rides in grad_scale for both), and a checkpoint
    loaded between steps reaches the optimizer-emitted bf16 shadow weights."""
    need_gpu()
    from msml_b200.engine import FlatSGD, TrainStep
    B = 8
    g = torch.Generator(device="cuda").manual_seed(5)
    imgs = [torch.randn(B, 3, 112, 112, device="cuda", generator=g) for _ in range(4)]
    labels = [torch.randint(0, 1000, (B,), device="cuda", generator=g) for _ in range(4)]
    out = {}
    for kind in ("torch", "flat"):
        net, pfc, opt, opt_pfc = _build(fp16=True, fused=True)
        if kind == "flat":
            opt = FlatSGD([p for p in net.parameters() if p.requires_grad], lr=0.01, momentum=0.9, weight_decay=5e-4)
        step = TrainStep(net, pfc, opt, opt_pfc, (B, 3, 112, 112), use_graph=use_graph)
        losses = [float(step(i, l)) for i, l in zip(imgs, labels)]
        out[kind] = (losses, host(net.frb.conv1.weight), host(net.frb.layer3[0].conv1.weight), host(net.frb.fc.weight), host(net.frb.bn2.weight))
        if kind == "flat":
            assert opt.is_bound() and step._emitted and not any(w is s for w in step._emitted for s in step._shadow_src)
            w = net.frb.conv1.weight
            assert torch.equal(w._msml_shadow, w.detach().to(torch.bfloat16))     # written by the optimizer kernel
            sd = {k: v.clone() for k, v in net.state_dict().items()}
            sd["frb.conv1.weight"] = sd["frb.conv1.weight"] * 0.5
            net.load_state_dict(sd)                                                 # in place through the Parameter: version moves
            assert not torch.equal(w._msml_shadow, w.detach().to(torch.bfloat16))   # stale until the next call notices
            step._sync_shadows()                                                    # what TrainStep.__call__ does first
            assert torch.equal(w._msml_shadow, w.detach().to(torch.bfloat16))
            assert np.isfinite(float(step(imgs[0], labels[0])))
    for a, b in zip(out["torch"][0], out["flat"][0]):
        assert abs(a - b) <= 3e-3 * abs(a), (out["torch"][0], out["flat"][0])
    for a, b in zip(out["torch"][1:], out["flat"][1:]):
        assert np.linalg.norm(a - b) <= 3e-3 * np.linalg.norm(a), np.linalg.norm(a - b) / np.linalg.norm(a)
