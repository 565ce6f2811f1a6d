"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares, the drop-in classes keep the reference's surface, and errors are loud (no fallback)."""
import ctypes
import json
import os
import re

import pytest
import torch

from conftest import GOLDEN, ROOT


def test_library_exports_every_declared_symbol():
    from msml_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "msml_b200.h")).read()
    declared = set(re.findall(r"\b(msml_[a-z0-9_]+)\s*\(", header))
    declared -= {"msml_status", "msml_dtype", "msml_act", "msml_arith", "msml_margin", "msml_resize", "msml_margin_params"}
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.msml_abi_version() == 1
    assert lib.msml_head_workspace(1024, 11679, 512) > 0
    assert lib.msml_pfc_select_workspace(125000) > 0


def test_argument_errors_are_reported_not_swallowed():
    from msml_b200 import _lib
    lib = _lib.load()
    # null pointers / bad enum -> negative status + message, no crash, no compute
    code = lib.msml_fm_gate_fwd(None, None, None, None, 16, 0, 1, 3, None)
    assert code < 0 and b"null" in lib.msml_last_error()
    code = lib.msml_fm_gate_fwd(ctypes.c_void_p(16), ctypes.c_void_p(16), None, ctypes.c_void_p(16), 16, 0, 7, 3, None)
    assert code < 0 and b"activation" in lib.msml_last_error()
    code = lib.msml_fm_gate_fwd(ctypes.c_void_p(8), ctypes.c_void_p(16), None, ctypes.c_void_p(16), 16, 0, 1, 3, None)
    assert code == -2      # MSML_EALIGN
    with pytest.raises(RuntimeError):
        _lib.check(code)


def test_ops_refuse_cpu_tensors():
    from msml_b200 import ops
    x = torch.randn(2, 8, 4, 4)
    with pytest.raises(RuntimeError, match="only a CUDA"):
        ops.fm_gate(x, x)
    with pytest.raises(RuntimeError, match="only a CUDA"):
        ops.dap(torch.randn(1, 18, 4, 4))
    with pytest.raises(ValueError, match="activation type error"):
        ops.fm_gate(x, x, act="relu")
    with pytest.raises(ValueError, match="arith type error"):
        ops.fm_gate(x, x, arith="pow")


def test_new_entry_points_validate_their_arguments_without_a_gpu():
    """msml_sgd_flat, msml_bn_fwd_ex, the peer-branch kernels and the MSE: argument errors come back as negative status codes
    with a message before anything touches the device."""
    from msml_b200 import _lib
    lib = _lib.load()
    p16 = ctypes.c_void_p(16)
    assert lib.msml_sgd_flat(p16, p16, p16, None, 6, p16, None, 0.9, 0.0, 0, None) < 0 and b"multiple of 4" in lib.msml_last_error()
    assert lib.msml_sgd_flat(p16, None, p16, None, 8, p16, None, 0.9, 0.0, 0, None) < 0 and b"null" in lib.msml_last_error()
    assert lib.msml_sgd_flat(p16, p16, p16, None, 8, p16, None, 0.0, 0.0, 1, None) < 0 and b"nesterov" in lib.msml_last_error()
    assert lib.msml_sgd_flat(ctypes.c_void_p(8), p16, p16, None, 8, p16, None, 0.9, 0.0, 0, None) == -2      # MSML_EALIGN
    assert lib.msml_sgd_flat(p16, p16, p16, None, 0, p16, None, 0.9, 0.0, 0, None) == 0                       # nothing to do
    assert lib.msml_fm_peer_mul_fwd(p16, p16, p16, p16, None, 16, 1, 0, 1, None) < 0 and b"come together" in lib.msml_last_error()
    assert lib.msml_fm_peer_mul_fwd(p16, p16, None, p16, None, 16, 1, 2, 1, None) < 0 and b"mode" in lib.msml_last_error()
    assert lib.msml_fm_peer_mul_bwd(p16, None, p16, p16, p16, p16, p16, 16, 1, 1, 1, None) < 0
    assert lib.msml_mse_fwd(p16, p16, 0, 1, p16, p16, 1 << 20, None) < 0
    assert lib.msml_mse_fwd(p16, p16, 16, 1, p16, p16, 4, None) < 0 and b"workspace" in lib.msml_last_error()
    assert lib.msml_mse_workspace() >= 148 * 4
    ws = lib.msml_bn_workspace(64, 32)
    # eval mode has no batch statistics to chain; a next-op workspace must be big enough and distinct from the op's own
    assert lib.msml_bn_fwd_ex(p16, None, p16, None, None, None, p16, p16, None, p16, p16, 64, 32, 1, 0, 0.1, 1e-5, p16, ws,
                              ctypes.c_void_p(32), ws, 0, None) < 0 and b"training" in lib.msml_last_error()
    assert lib.msml_bn_fwd_ex(p16, None, p16, None, None, None, None, None, None, p16, p16, 64, 32, 1, 1, 0.1, 1e-5, p16, ws,
                              ctypes.c_void_p(32), 8, 0, None) < 0 and b"next-op" in lib.msml_last_error()
    assert lib.msml_bn_fwd_ex(p16, None, p16, None, None, None, None, None, None, p16, p16, 64, 32, 1, 1, 0.1, 1e-5, p16, ws,
                              p16, ws, 0, None) < 0 and b"alias" in lib.msml_last_error()


def test_flat_sgd_surface_without_a_gpu():
    """engine.FlatSGD is a torch.optim.SGD (param_groups, state_dict, schedulers) that refuses what its kernel does not
    implement and has no unbound / CPU fallback."""
    from msml_b200.engine import FlatSGD
    net = torch.nn.Linear(4, 4)
    opt = FlatSGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
    assert isinstance(opt, torch.optim.SGD) and not opt.is_bound()
    assert opt.param_groups[0]["capturable_lr"] is True and opt.param_groups[0]["momentum"] == 0.9
    with pytest.raises(RuntimeError, match="not bound"):
        opt.step()
    with pytest.raises(ValueError, match="dampening"):
        FlatSGD(net.parameters(), lr=0.1, dampening=0.5)
    with pytest.raises(ValueError, match="one parameter group"):
        FlatSGD([{"params": [net.weight]}, {"params": [net.bias], "lr": 0.2}], lr=0.1)
    with pytest.raises(RuntimeError, match="only a CUDA"):
        net.weight.grad, net.bias.grad = torch.zeros(4, 4), torch.zeros(4)
        opt.bind_flat(list(net.parameters()), torch.zeros(20))
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda i: 0.5)
    assert abs(opt.param_groups[0]["lr"] - 0.05) < 1e-12
    sd = opt.state_dict()
    assert sd["param_groups"][0]["lr"] == opt.param_groups[0]["lr"] and sched is not None
    opt.zero_grad()                                 # unbound: torch's behaviour
    assert net.weight.grad is None


def test_msml_state_dict_matches_reference_layout():
    """Key names and shapes recorded from the reference models (tests/golden/state_keys.json)."""
    from msml_b200.backbones import MSML
    ref = json.load(open(os.path.join(GOLDEN, "state_keys.json")))
    for frb, classes in (("iresnet18", 10572), ("iresnet50", 97)):
        net = MSML(frb, "unet", (1, 1, 1, 1), classes, header_type="AMArcFace", fm_params=(3, 2, "sigmoid", "mul"))
        got = {k: list(v.shape) for k, v in net.state_dict().items()}
        assert got == ref[frb], set(got) ^ set(ref[frb])
    net = MSML("iresnet18", "unet", (1, 0, 1, 0), 10, header_type=None)
    assert net.classification is None and type(net.frb.fm_ops[1]).__name__ == "FMNone"


def test_msml_constructor_errors_mirror_reference():
    from msml_b200.backbones import MSML
    with pytest.raises(AssertionError):
        MSML("iresnet18", "unet", (1, 1, 1), 10)
    with pytest.raises(ValueError, match="FRB type error"):
        MSML("vgg", "unet", (1, 1, 1, 1), 10)
    with pytest.raises(ValueError, match="OSB type error"):
        MSML("iresnet18", "fcn", (1, 1, 1, 1), 10)
    with pytest.raises(ValueError, match="FM Operators type error"):
        MSML("iresnet18", "unet", (1, 2, 1, 1), 10)
    with pytest.raises(ValueError, match="not found"):
        MSML("iresnet99", "unet", (1, 1, 1, 1), 10)


def test_margin_heads_surface():
    from msml_b200.headers import AMArcFace, AMCosFace, ArcFace, CosFace, Softmax
    h = AMArcFace(16, 8, None, s=64.0, m=0.5, a=0.0, k=0.0)
    assert tuple(h.weight.shape) == (8, 16) and (h.kind, h.s, h.m) == ("arc", 64.0, 0.5)
    assert AMCosFace(16, 8, None).kind == "cos"
    with pytest.raises(ValueError, match="DataParallel"):
        Softmax(16, 8, [0])(torch.randn(2, 16), torch.zeros(2, dtype=torch.long))
    with pytest.raises(ValueError, match="DataParallel"):
        AMArcFace(16, 8, [0])(torch.randn(2, 16), torch.zeros(2, dtype=torch.long))
    assert ArcFace().kind == "arc" and CosFace().m == 0.4


def test_partial_fc_rejects_unfusable_margin_callable():
    from msml_b200.headers import PartialFC
    with pytest.raises(TypeError, match="kind, s, m, a, k"):
        PartialFC(0, 0, 1, 8, False, lambda logits, label: logits, 100)


def test_cpu_emulation_macro_never_reaches_the_product_build():
    """MSML_CPU_EMU (plain-C++ stand-ins for inline PTX, used by tests/emu to run the kernels on the host) must be defined by the
    emulation harness only: not by the build flags, not by any source under msml_b200/."""
    import glob
    import os
    from msml_b200 import _build
    assert not any("MSML_CPU_EMU" in f for f in _build.NVCC_FLAGS)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for path in glob.glob(os.path.join(root, "msml_b200", "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".cu", ".cuh", ".h", ".py")):
            text = open(path, encoding="utf-8", errors="ignore").read()
            assert "#define MSML_CPU_EMU" not in text and "-DMSML_CPU_EMU" not in text, path
    defs = [p for p in glob.glob(os.path.join(root, "tests", "emu", "*.cpp")) if "#define MSML_CPU_EMU" in open(p).read()]
    assert defs, "the emulation harness is where the macro is defined"
