"""Edge cases of the PartialFC head: a rank none of whose rows has its class on the shard, a batch of one, one class
repeated through the whole batch, shard boundaries with an odd class count, a two-class shard.  Checker: the fp64 oracle
on the same bf16-rounded inputs (oracle/partial_fc.py follows ref headers/partial_fc.py:118-177)."""
import threading

import numpy as np
import pytest
import torch

from gpu_util import assert_close, host, need_gpu
from oracle import partial_fc as opfc

pytestmark = pytest.mark.gpu


def run_ranks(W, B, C, D, labels, feats, weights, kind="arc", smak=(64.0, 0.5, 0.0, 0.0)):
    """W ranks as W host threads on one GPU (headers._comm.LockstepComm) -> per-rank (x_grad, loss, w_grad)."""
    from msml_b200.headers import MarginSoftmax, PartialFC
    from msml_b200.headers._comm import LockstepComm
    comms = LockstepComm.create(W) if W > 1 else [None]
    out, errors = {}, []

    def worker(rank):
        try:
            torch.cuda.set_device(0)
            pfc = PartialFC(rank, 0, W, B, False, MarginSoftmax(kind, *smak), C, sample_rate=1.0, embedding_size=D, comm=comms[rank])
            pfc.weight.copy_(weights[rank])
            x_grad, loss = pfc.forward_backward(labels[rank], feats[rank], None)
            torch.cuda.synchronize()
            out[rank] = (host(x_grad), float(loss), host(pfc.sub_weight.grad))
        except BaseException as e:  # noqa: BLE001
            errors.append(e)
            if W > 1:
                comms[rank].shared.barrier.abort()
    threads = [threading.Thread(target=worker, args=(r,)) for r in range(W)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    if errors:
        raise errors[0]
    return out


def check(W, B, C, D, labels, seed):
    torch.manual_seed(seed)
    feats = [torch.nn.functional.normalize(torch.randn(B, D, device="cuda")) for _ in range(W)]
    geo = [opfc.shard_geometry(C, W, r) for r in range(W)]
    weights = [torch.randn(g[0], D, device="cuda") * 0.01 for g in geo]
    got = run_ranks(W, B, C, D, labels, feats, weights)
    res = opfc.step([f.to(torch.bfloat16).double().cpu().numpy() for f in feats], [l.cpu().numpy() for l in labels],
                    [w.double().cpu().numpy() for w in weights], C, "arc", 64.0, 0.5)
    for r in range(W):
        x_grad, loss, w_grad = got[r]
        assert abs(loss - res["loss"]) <= 1e-3 * abs(res["loss"]), (r, loss, res["loss"])
        assert_close(x_grad, res["x_grad"][r], 2e-2, atol_frac=1e-2, what="x_grad r%d" % r)
        assert_close(w_grad, res["w_grad"][r], 2e-2, atol_frac=1e-2, what="w_grad r%d" % r)


def test_rank_without_any_positive_row():
    """W = 2, every label on shard 0: rank 1 remaps all rows to -1 (ref :79-81), contributes no target logit, no
    one-hot rows (ref :149-156) and a pure softmax gradient."""
    need_gpu()
    W, B, C, D = 2, 16, 600, 512
    labels = [torch.randint(0, 300, (B,), device="cuda") for _ in range(W)]
    check(W, B, C, D, labels, 21)


def test_batch_of_one_and_repeated_class():
    need_gpu()
    check(1, 1, 257, 512, [torch.tensor([200], device="cuda")], 22)                       # B_tot = 1
    check(1, 24, 1000, 512, [torch.full((24,), 999, dtype=torch.int64, device="cuda")], 23)      # the last class, 24 times
    check(2, 8, 1001, 512, [torch.tensor([0, 500, 501, 1000, 0, 500, 501, 1000], device="cuda")] * 2, 24)   # shard boundaries, odd C


def test_two_class_shard():
    """n_s = 2: the label-smoothing denominator n_s - 1 is 1 (ref :152), one class tile, mostly padding."""
    need_gpu()
    check(1, 8, 2, 512, [torch.tensor([0, 1, 1, 0, 0, 0, 1, 1], device="cuda")], 25)
