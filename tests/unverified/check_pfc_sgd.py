"""GPU parity checks of the fused PartialFC optimizer (csrc/pfc_sgd_kernels.cuh, headers/pfc_sgd.py) against the
reference recipe: stock torch.optim.SGD over module.parameters() + PartialFC.update() (ref train.py:188-191,299-300).

NOT collected by the default test run (never executed on a GPU yet); tests/test_gpu_unverified.py runs this file in a
subprocess and reports xfail / xpass.  The kernel's logic is covered on CPU by tests/test_emu_kernels.py.
"""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from gpu_util import assert_close, host, need_gpu  # noqa: E402


def run_steps(fused, sample_rate, nesterov=False, steps=3):
    from msml_b200.headers import ArcFace, PartialFC, PartialFCSGD
    torch.manual_seed(7)
    B, C, D = 16, 1000, 512
    pfc = PartialFC(0, 0, 1, B, False, ArcFace(64.0, 0.5), C, sample_rate=sample_rate, embedding_size=D)
    hp = dict(lr=0.1, momentum=0.9, weight_decay=5e-4, nesterov=nesterov)
    opt = PartialFCSGD(pfc, **hp) if fused else torch.optim.SGD([{"params": pfc.parameters()}], **hp)
    gen = torch.Generator(device="cuda").manual_seed(11)
    for _ in range(steps):
        feat = torch.nn.functional.normalize(torch.randn(B, D, device="cuda", generator=gen))
        label = torch.randint(0, C, (B,), device="cuda", generator=gen)
        pfc.forward_backward(label, feat, opt)
        opt.step()
        pfc.update()
    torch.cuda.synchronize()
    return pfc.weight.clone(), pfc.weight_mom.clone()


@pytest.mark.parametrize("sample_rate", [1.0, 0.3])
@pytest.mark.parametrize("nesterov", [False, True])
def test_fused_pfc_sgd_matches_stock_sgd_plus_update(sample_rate, nesterov):
    need_gpu()
    w_ref, m_ref = run_steps(False, sample_rate, nesterov)
    w, m = run_steps(True, sample_rate, nesterov)
    assert_close(host(w), host(w_ref), 1e-5, atol=1e-7, what="weight")
    assert_close(host(m), host(m_ref), 1e-5, atol=1e-7, what="weight_mom")
    assert not torch.equal(m_ref, torch.zeros_like(m_ref))


def test_fused_pfc_sgd_tensor_lr_emit_and_errors():
    need_gpu()
    from msml_b200 import _lib
    from msml_b200.headers import ArcFace, PartialFC, PartialFCSGD
    lib = _lib.load()
    torch.manual_seed(3)
    n, D = 37, 512
    w = torch.randn(n, D, device="cuda") * 0.01
    mom = torch.randn(n, D, device="cuda") * 0.001
    dw = torch.randn(n, D, device="cuda") * 0.1
    lr = torch.tensor(0.05, device="cuda")
    w0, m0 = w.clone(), mom.clone()
    wn = torch.empty(n, D, device="cuda", dtype=torch.bfloat16)
    inv = torch.empty(n, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.msml_pfc_sgd_update(w.data_ptr(), mom.data_ptr(), dw.data_ptr(), None, n, n, D, lr.data_ptr(), 0.0, 0.9, 5e-4, 0.0, 0,
                                       wn.data_ptr(), inv.data_ptr(), st))
    d = dw + 5e-4 * w0
    m_want = 0.9 * m0 + d
    w_want = w0 - 0.05 * m_want
    assert_close(host(mom), host(m_want), 1e-5, atol=1e-8, what="mom")
    assert_close(host(w), host(w_want), 1e-5, atol=1e-8, what="w")
    assert_close(host(wn), host(torch.nn.functional.normalize(w)), 1e-2, atol=1e-4, what="wn")
    assert_close(host(inv), host(1.0 / w.norm(dim=1)), 1e-5, what="inv_norm")
    assert lib.msml_pfc_sgd_update(w.data_ptr(), mom.data_ptr(), dw.data_ptr(), None, n, n, 100, None, 0.1, 0.9, 0.0, 0.0, 0, None, None, st) != 0
    assert lib.msml_pfc_sgd_update(w.data_ptr(), mom.data_ptr(), dw.data_ptr(), None, n, n, D, None, 0.1, 0.0, 0.0, 0.0, 1, None, None, st) != 0
    with pytest.raises(ValueError):
        PartialFCSGD(PartialFC(0, 0, 1, 4, False, ArcFace(), 64), lr=0.1, momentum=0.0, nesterov=True)
    with pytest.raises(TypeError):
        PartialFCSGD(torch.nn.Linear(4, 4), lr=0.1)
