"""GPU parity: tcgen05 GEMM, PartialFC sampling (K-D) and the PartialFC head (K-E..K-H) against the
oracle and the vectors generated from the reference (tests/golden)."""
import threading

import numpy as np
import pytest
import torch

from conftest import load_golden
from gpu_util import assert_close, dev, host, need_gpu
from oracle import partial_fc as opfc

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------- tcgen05 GEMM
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 512), (256, 512, 128), (100, 300, 72),
                                   (1, 8, 8), (300, 1000, 1032), (1024, 2048, 512)])
def test_gemm_bf16_tn_vs_fp32_matmul(M, N, K):
    need_gpu()
    from msml_b200 import ops
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    c = ops.gemm_tn(a, b)
    want = a.double().cpu() @ b.double().cpu().t()
    assert_close(host(c), want.numpy(), 1e-4, atol=1e-3 * K ** 0.5, what="gemm")


def test_gemm_exact_on_small_integers():
    """Integer-valued bf16 operands: the fp32 tensor-core accumulation is exact, so the result
    must equal the integer matmul bit for bit (catches any smem-descriptor / swizzle mismatch)."""
    need_gpu()
    from msml_b200 import ops
    torch.manual_seed(0)
    a = torch.randint(-4, 5, (384, 320), device="cuda").float()
    b = torch.randint(-4, 5, (520, 320), device="cuda").float()
    c = ops.gemm_tn(a, b)
    assert torch.equal(c, a @ b.t())


@pytest.mark.parametrize("a_mn", [False, True])
@pytest.mark.parametrize("b_mn", [False, True])
@pytest.mark.parametrize("block_n", [256, 512])
@pytest.mark.parametrize("M,N,K", [(128, 512, 128), (384, 520, 320), (200, 96, 1000), (1000, 512, 136)])
def test_gemm_operand_layouts_exact(a_mn, b_mn, block_n, M, N, K):
    """All four operand-major combinations (K-major / MN-major UMMA descriptors + their TMA boxes)
    and both accumulator tiles, on integer-valued operands: results must be bit-exact."""
    need_gpu()
    from msml_b200 import ops
    torch.manual_seed(M * 7 + N * 3 + K)
    a = torch.randint(-4, 5, (M, K), device="cuda").float()
    b = torch.randint(-4, 5, (N, K), device="cuda").float()
    a_store = a.t().contiguous() if a_mn else a
    b_store = b.t().contiguous() if b_mn else b
    if (a_store.shape[1] % 8) or (b_store.shape[1] % 8):
        pytest.skip("pitch must be a multiple of 8 elements")
    c = ops.gemm(a_store, b_store, a_mn=a_mn, b_mn=b_mn, block_n=block_n)
    assert torch.equal(c, a @ b.t())


@pytest.mark.parametrize("a_mn", [False, True])
@pytest.mark.parametrize("b_mn", [False, True])
@pytest.mark.parametrize("block_n", [128, 256])
@pytest.mark.parametrize("M,N,K", [(256, 512, 128), (384, 520, 320), (200, 96, 1000), (1000, 512, 136), (1024, 2048, 512), (128, 256, 64)])
def test_gemm_cta_pair_exact(a_mn, b_mn, block_n, M, N, K):
    """The CTA-pair mainloop (clusters of two CTAs, tcgen05.mma.cta_group::2 with M = 256, each CTA staging its 128 rows
    of A and half of the B tile): every operand-major combination on integer-valued operands, bit-exact — including an
    odd number of 128-row blocks (the last pair's second CTA has no rows) and ragged N / K."""
    need_gpu()
    from msml_b200 import ops
    torch.manual_seed(M * 7 + N * 3 + K + block_n)
    a = torch.randint(-4, 5, (M, K), device="cuda").float()
    b = torch.randint(-4, 5, (N, K), device="cuda").float()
    a_store = a.t().contiguous() if a_mn else a
    b_store = b.t().contiguous() if b_mn else b
    if (a_store.shape[1] % 8) or (b_store.shape[1] % 8):
        pytest.skip("pitch must be a multiple of 8 elements")
    c = ops.gemm(a_store, b_store, a_mn=a_mn, b_mn=b_mn, block_n=block_n, pair=True)
    assert torch.equal(c, a @ b.t())


# ------------------------------------------------------------------------------- sampling
def _pfc(rank, W, B, C, sr, D, kind="arc", smak=(64.0, 0.5, 0.0, 0.0), comm=None):
    from msml_b200.headers import MarginSoftmax, PartialFC
    return PartialFC(rank, 0, W, B, False, MarginSoftmax(kind, *smak), C, sample_rate=sr, embedding_size=D, comm=comm)


def test_remap_and_shards_bit_exact():
    need_gpu()
    from msml_b200 import _lib
    import ctypes
    lib = _lib.load()
    torch.manual_seed(1)
    labels = torch.randint(0, 93431, (1024,), device="cuda")
    for rank in range(8):
        num_local, class_start, _ = opfc.shard_geometry(93431, 8, rank)
        tl = labels.clone()
        _lib.check(lib.msml_pfc_remap(ctypes.c_void_p(tl.data_ptr()), tl.numel(), class_start, num_local, _lib.stream_ptr()))
        assert np.array_equal(tl.cpu().numpy(), opfc.remap_labels(labels.cpu().numpy(), class_start, num_local))


@pytest.mark.parametrize("num_local,sr,n_labels", [(1000, 0.3, 64), (125000, 0.1, 1024), (11679, 0.5, 1024),
                                                   (4097, 0.999, 16), (50, 0.1, 16), (1_000_000, 0.1, 128)])
def test_sample_index_bit_exact_vs_oracle(num_local, sr, n_labels):
    need_gpu()
    pfc = _pfc(0, 1, n_labels, num_local, sr, 64)
    torch.manual_seed(11)
    labels = torch.randint(0, num_local, (n_labels,), device="cuda")
    state = torch.cuda.get_rng_state()
    tl = labels.clone()
    pfc.sample(tl)
    torch.cuda.set_rng_state(state)
    perm = torch.rand(size=[num_local], device="cuda").cpu().numpy()     # the draw sample() consumed
    want_tl, want_index = opfc.sample(labels.cpu().numpy(), perm, 0, num_local, pfc.num_sample, sr)
    assert np.array_equal(pfc.index.cpu().numpy(), want_index)
    assert np.array_equal(tl.cpu().numpy(), want_tl)
    idx = pfc.index.cpu().numpy()
    assert (np.diff(idx) > 0).all()                                      # sorted, unique
    assert np.isin(np.unique(labels.cpu().numpy()), idx).all()           # every positive kept
    assert torch.equal(pfc.sub_weight.data, pfc.weight[pfc.index])


def test_sample_ties_at_threshold_take_lowest_index():
    need_gpu()
    import ctypes
    from msml_b200 import _lib
    lib = _lib.load()
    n, k = 10000, 2500
    perm = torch.full((n,), 0.5, device="cuda")
    perm[::7] = 0.75                     # 1429 strictly greater
    perm[5] = 2.0
    ws = torch.empty(lib.msml_pfc_select_workspace(n), dtype=torch.uint8, device="cuda")
    index = torch.empty(k, dtype=torch.int64, device="cuda")
    n_index = torch.zeros(1, dtype=torch.int64, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(lib.msml_pfc_select(p(perm), n, k, p(index), p(n_index), p(ws), ws.numel(), _lib.stream_ptr()))
    want = opfc.select_index(perm.cpu().numpy(), np.array([5]), k)
    assert int(n_index.item()) == k
    assert np.array_equal(index.cpu().numpy(), want)


def assert_rows_close(actual, desired, rtol, floor_frac=1e-3, what=""):
    """Per-ROW relative error: ||a_r - d_r||_2 <= rtol * ||d_r||_2 + floor_frac * median_r ||d_r||_2.
    assert_close's absolute floor is a fraction of the matrix maximum; in w_grad the target rows are 100-1000x larger
    than the rest, so that floor leaves the eps/(n_s - 1) label-smoothing rows (ref partial_fc.py:152-156) unchecked.
    Here every row is measured against its own norm."""
    a = np.asarray(actual, np.float64)
    d = np.asarray(desired, np.float64)
    assert a.shape == d.shape and np.isfinite(a).all(), what
    dn = np.linalg.norm(d, axis=1)
    err = np.linalg.norm(a - d, axis=1)
    tol = rtol * dn + floor_frac * np.median(dn)
    bad = err > tol
    if bad.any():
        i = int(np.argmax(err / np.maximum(tol, 1e-300)))
        raise AssertionError("%s: %d / %d rows out of tolerance; worst row %d: |err| %g, |want| %g (tol %g), median |want| %g" % (
            what, bad.sum(), bad.size, i, err[i], dn[i], tol[i], np.median(dn)))


def check_w_grad_rows(w_grad, want, labels_local, what):
    """Target rows and non-target rows of the class-centre gradient, each against its own scale."""
    n_s = want.shape[0]
    is_t = np.zeros(n_s, bool)
    lab = np.asarray(labels_local)
    is_t[lab[lab >= 0]] = True
    assert_rows_close(w_grad[is_t], want[is_t], 2e-2, what=what + " target rows")
    assert (~is_t).sum() > 0
    assert_rows_close(w_grad[~is_t], want[~is_t], 2e-2, what=what + " non-target rows (softmax + local label smoothing)")


# ------------------------------------------------------------------------------- PartialFC step
PFC = ["pfc_w1_full", "pfc_w1_sample", "pfc_w1_d512", "pfc_w1_overflow", "pfc_w2_full", "pfc_w2_sample", "pfc_w2_am"]


@pytest.mark.parametrize("name", PFC)
def test_partial_fc_step_matches_reference_golden(name, monkeypatch):
    need_gpu()
    from msml_b200.headers import partial_fc as mod
    from msml_b200.headers._comm import LockstepComm
    g = load_golden(name)
    W = int(g["W"])
    # per-rank replay of the recorded torch.rand draw (thread-safe: looked up on the PartialFC object)
    real_rand = torch.rand

    def rand_hook(*a, **k):
        cur = getattr(threading.current_thread(), "pfc_perm", None)
        return cur if cur is not None else real_rand(*a, **k)
    monkeypatch.setattr(mod.torch, "rand", rand_hook)

    results, errors = {}, []
    comms = LockstepComm.create(W) if W > 1 else [None]

    def worker(rank):
        th = threading.current_thread()

        def run():
            B, C, D, sr = int(g["B"]), int(g["C"]), int(g["D"]), float(g["sample_rate"])
            torch.cuda.set_device(0)
            pfc = _pfc(rank, W, B, C, sr, D, str(g["kind"]), tuple(float(v) for v in g["smak"]), comms[rank])
            pfc.weight.copy_(dev(g[f"r{rank}.w0"]))
            opt = torch.optim.SGD([{"params": pfc.parameters()}], lr=0.1, momentum=0.9, weight_decay=5e-4)
            out = []
            for step in range(int(g["steps"])):
                p = f"r{rank}.s{step}."
                if step > 0:     # every step starts from the reference's own state (bf16 noise must not compound)
                    pfc.weight.copy_(dev(g[f"r{rank}.s{step - 1}.weight_after"]))
                    pfc.weight_mom.copy_(dev(g[f"r{rank}.s{step - 1}.mom_after"]))
                perm = g[p + "perm"]
                th.pfc_perm = dev(perm) if perm.size else None
                x_grad, loss = pfc.forward_backward(dev(g[p + "label"]), dev(g[p + "feat"]), opt)
                th.pfc_perm = None
                rec = dict(x_grad=host(x_grad), loss=float(loss.item()), w_grad=host(pfc.sub_weight.grad),
                           index=None if pfc.index is None else pfc.index.cpu().numpy().copy())
                opt.step()
                pfc.update()
                opt.zero_grad()
                rec["weight_after"], rec["mom_after"] = host(pfc.weight), host(pfc.weight_mom)
                out.append(rec)
            results[rank] = out
        try:
            run()
        except BaseException as e:  # noqa: BLE001
            errors.append((rank, e))
            if W > 1:
                comms[rank].shared.barrier.abort()

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(W)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    if errors:
        raise errors[0][1]
    sampled = int(float(g["sample_rate"])) != 1
    for rank in range(W):
        for step, rec in enumerate(results[rank]):
            p = f"r{rank}.s{step}."
            if sampled:      # bit-exact sampled class indices
                assert np.array_equal(rec["index"], g[p + "index"]), (name, rank, step)
            want_loss = float(g[p + "loss"])
            assert abs(rec["loss"] - want_loss) <= 1e-3 * abs(want_loss), (rec["loss"], want_loss)
            assert_close(rec["x_grad"], g[p + "x_grad"], 2e-2, atol_frac=1e-2, what=f"{name} x_grad r{rank} s{step}")
            assert_close(rec["w_grad"], g[p + "w_grad"], 2e-2, atol_frac=1e-2, what=f"{name} w_grad r{rank} s{step}")
            assert_close(rec["weight_after"], g[p + "weight_after"], 2e-2, atol_frac=1e-2, what="weight_after")
            assert_close(rec["mom_after"], g[p + "mom_after"], 2e-2, atol_frac=1e-2, what="mom_after")


@pytest.mark.parametrize("B_tot,C,kind,smak", [(128, 3000, "arc", (64.0, 0.5, 0.0, 0.0)),
                                               (256, 2051, "cos", (64.0, 0.4, 0.0, 0.0)),
                                               (130, 777, "arc", (32.0, 0.45, 1.2, 0.1))])
def test_head_step_vs_oracle_fp64(B_tot, C, kind, smak):
    """Mid-size single-rank step (D=512) against the fp64 oracle on the SAME bf16-rounded inputs."""
    need_gpu()
    torch.manual_seed(B_tot)
    D = 512
    pfc = _pfc(0, 1, B_tot, C, 1.0, D, kind, smak)
    feat = torch.nn.functional.normalize(torch.randn(B_tot, D, device="cuda"))
    label = torch.randint(0, C, (B_tot,), device="cuda")
    # make a few targets well aligned with their class centre so the margin matters
    with torch.no_grad():
        pfc.weight[label[:8]] = feat[:8] * 0.01 + pfc.weight[label[:8]] * 0.2
    x_grad, loss = pfc.forward_backward(label, feat, None)
    xb = feat.to(torch.bfloat16).double().cpu().numpy()
    w = pfc.weight.double().cpu().numpy()
    res = opfc.step([xb], [label.cpu().numpy()], [w], C, kind, *smak)
    assert abs(float(loss) - res["loss"]) <= 1e-3 * abs(res["loss"])
    assert_close(host(x_grad), res["x_grad"][0], 2e-2, atol_frac=1e-2, what="x_grad")
    assert_close(host(pfc.sub_weight.grad), res["w_grad"][0], 2e-2, atol_frac=1e-2, what="w_grad")
    assert_rows_close(host(x_grad), res["x_grad"][0], 2e-2, what="x_grad rows")
    check_w_grad_rows(host(pfc.sub_weight.grad), res["w_grad"][0], res["total_label"][0], "w_grad")


def test_label_smoothing_term_is_visible_in_non_target_rows():
    """The eps/(n_s - 1) local label smoothing (ref partial_fc.py:152-156) must be IN the gradient: the oracle
    evaluated without it (EPSILON = 0) has to FAIL the same row check the kernel passes."""
    need_gpu()
    torch.manual_seed(31)
    B_tot, C, D = 128, 3000, 512
    pfc = _pfc(0, 1, B_tot, C, 1.0, D)
    feat = torch.nn.functional.normalize(torch.randn(B_tot, D, device="cuda"))
    label = torch.randint(0, C, (B_tot,), device="cuda")
    pfc.forward_backward(label, feat, None)
    xb = feat.to(torch.bfloat16).double().cpu().numpy()
    w = pfc.weight.double().cpu().numpy()
    res = opfc.step([xb], [label.cpu().numpy()], [w], C, "arc", 64.0, 0.5)
    got = host(pfc.sub_weight.grad)
    check_w_grad_rows(got, res["w_grad"][0], res["total_label"][0], "w_grad")
    saved = opfc.EPSILON
    opfc.EPSILON = 0.0
    try:
        res0 = opfc.step([xb], [label.cpu().numpy()], [w], C, "arc", 64.0, 0.5)
    finally:
        opfc.EPSILON = saved
    with pytest.raises(AssertionError):
        check_w_grad_rows(got, res0["w_grad"][0], res["total_label"][0], "w_grad without smoothing")


@pytest.mark.parametrize("B_tot,n_s", [(1024, 11679), (1024, 11678), (1024, 20011), (300, 40003)])
def test_head_step_at_config3_w8_rank_shape_vs_oracle(B_tot, n_s):
    """The GEMM shapes one rank of the 8-GPU BASELINE config-3 run sees (B_tot = 8 x 128 gathered rows against an
    11,679 / 11,678-class shard; 8 M-blocks, ragged last class tile) against the fp64 oracle; the two larger shards are
    long enough for the resident-A forward kernel (tc_gemm.cuh: gemm_ares_kernel) to be selected, one of them with a ragged
    last row block and an odd number of row blocks for the CTA-pair dX kernel."""
    need_gpu()
    torch.manual_seed(n_s)
    D = 512
    pfc = _pfc(0, 1, B_tot, n_s, 1.0, D)
    feat = torch.nn.functional.normalize(torch.randn(B_tot, D, device="cuda"))
    label = torch.randint(0, n_s, (B_tot,), device="cuda")
    label[::3] = torch.randint(n_s - 40, n_s, (label[::3].numel(),), device="cuda")      # load the ragged last tile
    with torch.no_grad():
        pfc.weight[label[:64]] = feat[:64] * 0.01 + pfc.weight[label[:64]] * 0.2
    x_grad, loss = pfc.forward_backward(label, feat, None)
    res = opfc.step([feat.to(torch.bfloat16).double().cpu().numpy()], [label.cpu().numpy()], [pfc.weight.double().cpu().numpy()],
                    n_s, "arc", 64.0, 0.5)
    assert abs(float(loss) - res["loss"]) <= 1e-3 * abs(res["loss"])
    assert_close(host(x_grad), res["x_grad"][0], 2e-2, atol_frac=1e-2, what="x_grad")
    assert_rows_close(host(x_grad), res["x_grad"][0], 2e-2, what="x_grad rows")
    check_w_grad_rows(host(pfc.sub_weight.grad), res["w_grad"][0], res["total_label"][0], "w_grad")


def test_head_config3_shape_properties():
    """BASELINE config 3 single-GPU shape (B=128, 93,431 classes): size-independent checks —
    softmax gradient rows sum to ~0 through the margin-free columns, loss equals -log p_target
    recomputed from a dense fp32 torch evaluation of the same bf16 operands."""
    need_gpu()
    torch.manual_seed(5)
    B, C, D = 128, 93431, 512
    pfc = _pfc(0, 1, B, C, 1.0, D)
    feat = torch.nn.functional.normalize(torch.randn(B, D, device="cuda"))
    label = torch.randint(0, C, (B,), device="cuda")
    x_grad, loss = pfc.forward_backward(label, feat, None)
    xb = feat.to(torch.bfloat16).float()
    wn = torch.nn.functional.normalize(pfc.weight).to(torch.bfloat16).float()
    cos = xb @ wn.t()
    theta = torch.acos(cos[torch.arange(B), label])
    logits = cos * 64.0
    logits[torch.arange(B), label] = torch.cos(theta + 0.5) * 64.0
    want = torch.nn.functional.cross_entropy(logits.double(), label)
    assert abs(float(loss) - float(want)) <= 1e-3 * float(want)
    p = torch.softmax(logits.double(), 1)
    t = torch.full_like(p, 0.1 / (C - 1))
    t[torch.arange(B), label] = 0.9
    gl = (p - t) / B * 64.0
    gl[torch.arange(B), label] *= (torch.sin(theta + 0.5) / torch.sin(theta)).double()
    assert_close(host(x_grad), (gl @ wn.double()).cpu().numpy(), 2e-2, atol_frac=1e-2, what="x_grad cfg3")
    # class-centre gradient through the normalise backward (ref :115,169), every row against its own norm
    dwn = gl.t() @ xb.double()
    w64 = pfc.weight.double()
    nrm = w64.norm(dim=1, keepdim=True)
    wn64 = w64 / nrm
    want_dw = ((dwn - wn64 * (wn64 * dwn).sum(1, keepdim=True)) / nrm).cpu().numpy()
    check_w_grad_rows(host(pfc.sub_weight.grad), want_dw, label.cpu().numpy(), "w_grad cfg3")
