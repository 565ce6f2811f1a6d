"""GPU parity: mask-fusion kernels (K-A), DAP (K-B) vs the oracle and the reference goldens."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from gpu_util import assert_close, dev, host, need_gpu
from oracle import dap as odap
from oracle import fm_tail

pytestmark = pytest.mark.gpu

FM = ["fm_c64_sigmoid_mul", "fm_c128_tanh_add", "fm_c64_sigmoid_div", "fm_c256_tanh_sub", "fm_c512_sigmoid_mul"]


@pytest.mark.parametrize("name", FM)
@pytest.mark.parametrize("cl", [False, True])
def test_fm_gate_fp32_matches_reference_golden(name, cl):
    need_gpu()
    from msml_b200 import ops
    g = load_golden(name)
    act, arith = str(g["act"]), str(g["arith"])
    fmt = torch.channels_last if cl else torch.contiguous_format
    yf = dev(g["yf"]).contiguous(memory_format=fmt).requires_grad_(True)
    z = dev(g["z"]).contiguous(memory_format=fmt).requires_grad_(True)
    out = ops.fm_gate(yf, z, act, arith)
    out.backward(dev(g["dout"]))
    assert_close(host(out), g["out"], 1e-5, atol=1e-5, what="out")
    assert_close(host(yf.grad), g["dyf_direct"], 1e-4, atol=1e-5, what="dyf")
    assert_close(host(z.grad), g["dz"], 1e-4, atol=1e-5, what="dz")


@pytest.mark.parametrize("act", ["sigmoid", "tanh"])
@pytest.mark.parametrize("arith", ["add", "sub", "div", "mul"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_fm_gate_all_modes_vs_oracle(act, arith, dtype):
    need_gpu()
    from msml_b200 import ops
    torch.manual_seed(1)
    shape = (3, 64, 9, 7)     # 12096 elements: vector body + ragged tail for every dtype
    yf = torch.randn(shape, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
    z = (torch.randn(shape, device="cuda") + (1.5 if arith == "div" else 0.0)).to(dtype).contiguous(memory_format=torch.channels_last)
    if arith == "div" and act == "tanh":
        z = z.abs() + 0.5
    fo = torch.randn(shape, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
    d = torch.randn(shape, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
    yf.requires_grad_(True); z.requires_grad_(True); fo.requires_grad_(True)
    out = ops.fm_gate(yf, z, act, arith, fo)
    out.backward(d)
    rt = 2e-2 if dtype != torch.float32 else 2e-5
    want = fm_tail.fm_gate_fwd(host(yf), host(z), act, arith, host(fo))
    assert_close(host(out), want, rt, atol=1e-5 if dtype == torch.float32 else 1e-2, what="out")
    wdyf, wdz = fm_tail.fm_gate_bwd(host(d), host(yf), host(z), act, arith)
    assert_close(host(yf.grad), wdyf, rt, atol_frac=1e-5 if dtype == torch.float32 else 4e-3, what="dyf")
    assert_close(host(z.grad), wdz, rt, atol_frac=1e-5 if dtype == torch.float32 else 4e-3, what="dz")
    assert_close(host(fo.grad), host(d), 0, what="df_out")


@pytest.mark.parametrize("n", [1, 7, 8, 4096 + 3])
def test_fm_gate_ragged_sizes(n):
    need_gpu()
    from msml_b200 import ops
    torch.manual_seed(n)
    yf = torch.randn(n, device="cuda", dtype=torch.bfloat16).view(1, n, 1, 1)
    z = torch.randn(n, device="cuda", dtype=torch.bfloat16).view(1, n, 1, 1)
    out = ops.fm_gate(yf, z, "sigmoid", "mul")
    assert_close(host(out), fm_tail.fm_gate_fwd(host(yf), host(z)), 2e-2, atol=1e-2, what="ragged")


def test_fm_gate_multi_scale_single_launch():
    need_gpu()
    from msml_b200 import ops
    torch.manual_seed(2)
    shapes = [(4, 64, 56, 56), (4, 128, 28, 28), (4, 256, 14, 14), (4, 512, 7, 7)]
    mk = lambda s: torch.randn(s, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    yfs, zs, ds = [mk(s) for s in shapes], [mk(s) for s in shapes], [mk(s) for s in shapes]
    ops.launch_count_reset()
    outs = ops.fm_gate_multi_fwd(yfs, zs)
    dyfs, dzs = ops.fm_gate_multi_bwd(ds, yfs, zs)
    assert ops.launch_count() == 2
    for yf, z, d, o, dy, dz in zip(yfs, zs, ds, outs, dyfs, dzs):
        assert_close(host(o), fm_tail.fm_gate_fwd(host(yf), host(z)), 2e-2, atol=1e-2, what="multi out")
        wdy, wdz = fm_tail.fm_gate_bwd(host(d), host(yf), host(z))
        assert_close(host(dy), wdy, 2e-2, atol_frac=4e-3, what="multi dyf")
        assert_close(host(dz), wdz, 2e-2, atol_frac=4e-3, what="multi dz")


def test_fm_gate_linearity_at_full_size():
    """Size-independent property at the BASELINE config-2 scale shape (B=512, 64x56x56):
    with arith=mul the op is linear in yf:  f(a*yf, z) == a * f(yf, z)  up to rounding."""
    need_gpu()
    from msml_b200 import ops
    torch.manual_seed(3)
    shape = (512, 64, 56, 56)
    yf = torch.randn(shape, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    z = torch.randn(shape, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    a = ops.fm_gate(yf, z, "sigmoid", "mul")
    b = ops.fm_gate(yf * 2, z, "sigmoid", "mul")       # exact doubling in bf16
    assert torch.equal(b, a * 2)
    ref = yf.float() * (1 + torch.sigmoid(z.float()))
    assert (a.float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()


@pytest.mark.parametrize("C,Co,H,W", [(64, 18, 9, 7), (128, 18, 5, 5), (512, 18, 7, 7), (8, 3, 4, 3), (16, 8, 3, 5)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_fm_cat_matches_concat(C, Co, H, W, dtype):
    """K-C (ref fmoperator.py:277-279): cat(yf, yo) + zero pad is a pure copy -> bit-exact against numpy's concatenate;
    backward = column slice of the concat's gradient + the tail's gradient, summed in fp32 and rounded once."""
    need_gpu()
    from msml_b200 import ops
    torch.manual_seed(11)
    B = 3
    yf0 = torch.randn(B, C, H, W, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
    yo0 = torch.randn(B, Co, H, W, device="cuda")                       # fp32 NCHW, as the OSB emits outside autocast
    yf = yf0.clone().requires_grad_(True)
    yo = yo0.clone().requires_grad_(True)
    x, pad, yf_tail = ops.fm_cat(yf, yo)
    Ct = x.shape[1]
    assert Ct % 8 == 0 and pad == Ct - C - Co and 0 <= pad < 8
    assert x.is_contiguous(memory_format=torch.channels_last)
    want = np.concatenate([host(yf0), host(yo0.to(dtype)), np.zeros((B, pad, H, W))], axis=1)
    assert np.array_equal(host(x), want)
    assert torch.equal(yf_tail, yf0)
    dcat = torch.randn(B, Ct, H, W, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
    dtail = torch.randn(B, C, H, W, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
    torch.autograd.backward([x, yf_tail], [dcat, dtail])
    want_dyf = (dcat[:, :C].float().cpu() + dtail.float().cpu()).to(dtype)
    assert torch.equal(yf.grad.cpu(), want_dyf)
    assert yo.grad.dtype == torch.float32
    assert torch.equal(yo.grad.cpu(), dcat[:, C:C + Co].float().cpu())
    # only one of the two consumers carries a gradient
    yf = yf0.clone().requires_grad_(True)
    x, _, yf_tail = ops.fm_cat(yf, yo0)
    x.backward(dcat)
    assert torch.equal(yf.grad, dcat[:, :C])
    yf = yf0.clone().requires_grad_(True)
    x, _, yf_tail = ops.fm_cat(yf, yo0)
    yf_tail.backward(dtail)
    assert torch.equal(yf.grad, dtail)
    # no autograd: same forward, yf handed through
    with torch.no_grad():
        x2, pad2, t2 = ops.fm_cat(yf0, yo0)
    assert torch.equal(x2, x) and pad2 == pad and torch.equal(t2, yf0)


def test_fm_cat_full_size_and_argument_errors():
    """Stage-1 shape of BASELINE config 3 (128 x 64 x 56 x 56 + 18 maps): column checks on the device; bad shapes raise."""
    need_gpu()
    from msml_b200 import _lib, ops
    torch.manual_seed(12)
    yf = torch.randn(128, 64, 56, 56, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    yo = torch.randn(128, 18, 56, 56, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    x, pad, _ = ops.fm_cat(yf, yo)
    assert x.shape == (128, 88, 56, 56) and pad == 6
    assert torch.equal(x[:, :64], yf) and torch.equal(x[:, 64:82], yo) and not x[:, 82:].any()
    with pytest.raises(ValueError):
        ops.fm_cat(yf, yo[:, :, :28])
    with pytest.raises(ValueError):                      # no torch.cat fallback for channel counts the kernel does not take
        ops.fm_cat(yf[:, :60], yo)
    with pytest.raises(RuntimeError):                    # ... nor for CPU tensors
        ops.fm_cat(yf.cpu(), yo.cpu())
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    assert lib.msml_fm_cat_fwd(yf.data_ptr(), yo.data_ptr(), x.data_ptr(), 10, 64, 18, 80, _lib.BF16, st) != 0    # Ct < C + Co
    assert b"smaller" in lib.msml_last_error()
    assert lib.msml_fm_cat_fwd(yf.data_ptr(), yo.data_ptr(), x.data_ptr(), 10, 60, 18, 88, _lib.BF16, st) != 0    # C % 8
    assert lib.msml_fm_cat_bwd(None, None, x.data_ptr(), None, 10, 64, 18, 88, _lib.BF16, st) != 0                # null dcat


@pytest.mark.parametrize("Cm,Hm,Wm", [(1, 8, 6), (1, 4, 3), (64, 8, 6), (64, 4, 3)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fm_mask_resize_broadcast_extension(Cm, Hm, Wm, dtype):
    need_gpu()
    from msml_b200 import ops
    torch.manual_seed(5)
    B, C, H, W = 3, 64, 8, 6
    yf = torch.randn(B, C, H, W, device="cuda").to(dtype).requires_grad_(True)
    m = torch.randn(B, Cm, Hm, Wm, device="cuda").to(dtype).requires_grad_(True)
    d = torch.randn(B, C, H, W, device="cuda").to(dtype)
    out = ops.fm_mask(yf, m, "sigmoid", "mul")
    out.backward(d)
    nhwc = lambda t: host(t).transpose(0, 2, 3, 1)
    want = fm_tail.fm_mask_fwd(nhwc(yf), nhwc(m))
    wdyf, wdm = fm_tail.fm_mask_bwd(nhwc(d), nhwc(yf), nhwc(m))
    rt = 2e-2 if dtype == torch.bfloat16 else 2e-5
    assert_close(nhwc(out), want, rt, atol=1e-2 if dtype == torch.bfloat16 else 1e-5, what="mask out")
    assert_close(nhwc(yf.grad), wdyf, rt, atol_frac=4e-3 if dtype == torch.bfloat16 else 1e-5, what="mask dyf")
    assert_close(nhwc(m.grad), wdm, rt, atol_frac=1e-2 if dtype == torch.bfloat16 else 1e-5, what="mask dm")


@pytest.mark.parametrize("cl", [False, True])
def test_dap_matches_reference_golden(cl):
    need_gpu()
    from msml_b200 import ops
    g = load_golden("dap")
    x = dev(g["x"])
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    y, mask = ops.dap_with_mask(x, 3)
    y.backward(dev(g["dy"]))
    assert_close(host(y), g["y"], 1e-6, atol=1e-6, what="dap y")
    assert_close(host(x.grad), g["dx"], 1e-6, atol=1e-7, what="dap dx")
    assert np.array_equal(mask.cpu().numpy(), g["mask"])      # bit-exact, planted ties included
    y2 = ops.dap(x.detach(), 3)
    assert torch.equal(y2, y)


def test_dap_full_size_properties():
    """112x112, batch 128: constant-in-group input is reproduced exactly; mask == 2-way argmax."""
    need_gpu()
    from msml_b200 import ops
    torch.manual_seed(7)
    base = torch.randn(128, 2, 112, 112, device="cuda", dtype=torch.bfloat16)
    x = base.repeat_interleave(9, dim=1)
    y, mask = ops.dap_with_mask(x.float(), 3)
    assert_close(host(y), host(base), 1e-6, atol=1e-6, what="dap const groups")
    assert torch.equal(mask, (base[:, 1].float() > base[:, 0].float()).long())
    assert np.array_equal(odap.argmax_mask(host(y)), mask.cpu().numpy())


# ------------------------------------------------------------------------------- fused BN / PReLU / residual
class _BNRef(torch.nn.Module):
    def __init__(self, C, prelu):
        super().__init__()
        self.bn = torch.nn.BatchNorm2d(C, eps=1e-05)
        self.prelu = torch.nn.PReLU(C) if prelu else None


@pytest.mark.parametrize("C,H", [(32, 9), (64, 14), (128, 7), (512, 7), (256, 3)])
@pytest.mark.parametrize("prelu", [False, True])
@pytest.mark.parametrize("res", [False, True])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_bn_act_train_vs_oracle(C, H, prelu, res, dtype):
    need_gpu()
    from msml_b200 import ops
    from oracle import bn_act as obn
    torch.manual_seed(C + H)
    B = 5
    m = _BNRef(C, prelu).cuda().train()
    with torch.no_grad():
        m.bn.weight.uniform_(0.5, 1.5); m.bn.bias.normal_(0, 0.2)
        m.bn.running_mean.normal_(0, 0.1); m.bn.running_var.uniform_(0.5, 1.5)
        if prelu:
            m.prelu.weight.uniform_(0.1, 0.4)
    rm0, rv0 = host(m.bn.running_mean), host(m.bn.running_var)
    x = (torch.randn(B, C, H, H, device="cuda") * 1.3 + 0.4).to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    r = torch.randn(B, C, H, H, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True) if res else None
    dy = torch.randn(B, C, H, H, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
    y = ops.bn_act(x, m.bn, m.prelu, r)
    y.backward(dy)
    flat = lambda t: host(t).transpose(0, 2, 3, 1).reshape(-1, C)
    g, b = host(m.bn.weight), host(m.bn.bias)
    a = host(m.prelu.weight) if prelu else None
    want, st = obn.bn_act_fwd(flat(x), g, b, a, flat(r) if res else None, True, rm0, rv0)
    wb = obn.bn_act_bwd(flat(dy), flat(x), g, b, a, flat(r) if res else None)
    lo = dtype == torch.bfloat16
    assert_close(flat(y), want, 2e-2 if lo else 2e-5, atol=2e-2 if lo else 2e-5, what="bn y")
    assert_close(host(m.bn.running_mean), st["running_mean"], 1e-5, atol=1e-6, what="running_mean")
    assert_close(host(m.bn.running_var), st["running_var"], 1e-4, atol=1e-6, what="running_var")
    assert int(m.bn.num_batches_tracked) == 1
    assert_close(flat(x.grad), wb["dx"], 2e-2 if lo else 1e-4, atol_frac=1e-2 if lo else 1e-5, what="bn dx")
    assert_close(host(m.bn.weight.grad), wb["dgamma"], 2e-2 if lo else 1e-4, atol_frac=1e-2 if lo else 1e-5, what="dgamma")
    assert_close(host(m.bn.bias.grad), wb["dbeta"], 2e-2 if lo else 1e-4, atol_frac=1e-2 if lo else 1e-5, what="dbeta")
    if prelu:
        assert_close(host(m.prelu.weight.grad), wb["dprelu"], 2e-2 if lo else 1e-4, atol_frac=1e-2 if lo else 1e-5, what="dprelu")
    if res:
        assert_close(flat(r.grad), wb["dres"], 2e-2 if lo else 1e-5, atol_frac=1e-2 if lo else 1e-6, what="dres")


def test_bn_act_eval_uses_running_stats():
    need_gpu()
    from msml_b200 import ops
    from oracle import bn_act as obn
    torch.manual_seed(3)
    C = 64
    m = _BNRef(C, True).cuda().eval()
    with torch.no_grad():
        m.bn.running_mean.normal_(0, 0.3); m.bn.running_var.uniform_(0.5, 2.0); m.bn.weight.uniform_(0.5, 1.5)
    x = torch.randn(3, C, 6, 5, device="cuda").contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        y = ops.bn_act(x, m.bn, m.prelu)
    flat = lambda t: host(t).transpose(0, 2, 3, 1).reshape(-1, C)
    want, _ = obn.bn_act_fwd(flat(x), host(m.bn.weight), host(m.bn.bias), host(m.prelu.weight), None, False,
                             host(m.bn.running_mean), host(m.bn.running_var))
    assert_close(flat(y), want, 2e-5, atol=2e-5, what="bn eval")
    assert int(m.bn.num_batches_tracked) == 0


def test_bn_act_full_size_statistics():
    """Stage-1 shape at batch 128 (64 x 56 x 56, bf16): the output is normalised per channel."""
    need_gpu()
    from msml_b200 import ops
    torch.manual_seed(4)
    m = _BNRef(64, False).cuda().train()
    x = (torch.randn(128, 64, 56, 56, device="cuda") * 3 + 5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y = ops.bn_act(x, m.bn).float()
    assert y.mean((0, 2, 3)).abs().max().item() < 5e-3
    assert (y.var((0, 2, 3), unbiased=False) - 1).abs().max().item() < 1e-2


def test_accum_bf16_multi_matches_torch():
    """dst_f32 += src_bf16 over many ragged / unaligned segments in one launch (csrc/optim.cu)."""
    need_gpu()
    import ctypes
    from msml_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(3)
    sizes = [1, 7, 8, 9, 4096, 8191, 8192, 8193, 100003, 3 * 3 * 64 * 64] + [257 + 13 * i for i in range(110)]   # > one table
    big_d = torch.randn(sum(sizes) + 8 * len(sizes) + 8, device="cuda", generator=g)
    big_s = torch.randn(sum(sizes) + 8 * len(sizes) + 8, device="cuda", generator=g).to(torch.bfloat16)
    want = big_d.clone()
    dst, src, off = [], [], 0
    for i, n in enumerate(sizes):
        o = off + (i % 3)                      # some segments start off a 16-byte boundary
        dst.append(big_d[o:o + n]); src.append(big_s[o + 1:o + 1 + n] if i % 5 == 0 else big_s[o:o + n])
        want[o:o + n] += src[-1].float()
        off = o + n + 5
    k = len(sizes)
    d = (ctypes.c_void_p * k)(*[t.data_ptr() for t in dst]); s = (ctypes.c_void_p * k)(*[t.data_ptr() for t in src])
    n = (ctypes.c_int64 * k)(*sizes)
    _lib.check(lib.msml_accum_bf16_multi(k, d, s, n, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert torch.equal(big_d, want)


def test_padded_channel_paths_match_unpadded(monkeypatch):
    """Zero-padded channel counts (FM concat, image stems, U-Net decoder) must not change the bf16 forward beyond
    rounding: same model with the padding helpers disabled."""
    need_gpu()
    from msml_b200 import ops
    from msml_b200.backbones import MSML
    import importlib
    unet_mod = importlib.import_module("msml_b200.backbones.osb.unet")
    from oracle.detfill import det_tensor, fill_state_dict_
    net = MSML("iresnet18", "unet", (1, 1, 1, 1), 97, fp16=True, header_type=None, fm_params=(3, 2, "sigmoid", "mul"))
    fill_state_dict_(net)
    net = net.cuda().eval()
    x = det_tensor("model.x", (2, 3, 112, 112)).cuda()
    with torch.no_grad():
        f1, s1 = net(x)
        monkeypatch.setattr(ops, "cat_channels_padded", lambda parts, multiple=8: (torch.cat(list(parts), 1), 0))
        monkeypatch.setattr(ops, "fm_cat", lambda yf, yo, multiple=8: (torch.cat([yf, yo.to(yf.dtype)], 1), 0, yf))
        monkeypatch.setattr(unet_mod.Unet, "pad_channels", False)
        f2, s2 = net(x)
    cos = torch.nn.functional.cosine_similarity(f1.double(), f2.double())
    assert (cos > 0.999).all(), cos
    assert_close(host(s1), host(s2), 3e-2, atol_frac=3e-2, what="seg padded vs unpadded")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_bn_act_fork_fuses_the_skip_gradient(dtype):
    """(bn(x), x) with the skip gradient added inside the backward kernel == bn(x) and x used separately."""
    need_gpu()
    from msml_b200 import ops
    torch.manual_seed(4)
    bn = torch.nn.BatchNorm2d(64).cuda().train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.5, 0.5)
    x0 = torch.randn(4, 64, 14, 14, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
    d1 = torch.randn_like(x0); d2 = torch.randn_like(x0)
    res = {}
    for mode in ("plain", "fork"):
        x = x0.clone().requires_grad_(True)
        bn.zero_grad(set_to_none=True)
        if mode == "plain":
            y, xs = ops.bn_act(x, bn), x
        else:
            y, xs = ops.bn_act_fork(x, bn)
        (y * d1).sum().backward(retain_graph=True) if False else ((y * d1).sum() + (xs * d2).sum()).backward()
        res[mode] = (host(y), host(x.grad), host(bn.weight.grad), host(bn.bias.grad))
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    for a, b, what in zip(res["fork"], res["plain"], ("y", "dx", "dgamma", "dbeta")):
        assert_close(a, b, tol, atol_frac=tol, what=what)
    # only the skip branch used: gradient passes straight through
    x = x0.clone().requires_grad_(True)
    _y, xs = ops.bn_act_fork(x, bn)
    (xs * d2).sum().backward()
    assert_close(host(x.grad), host(d2), 0.0, atol=0.0, what="skip only")


@pytest.mark.parametrize("C,H,B", [(64, 14, 5), (128, 7, 6), (512, 7, 4), (64, 56, 32)])
@pytest.mark.parametrize("consumer", ["fork", "plain_prelu"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_bn_chained_statistics_match_unchained(C, H, B, consumer, dtype):
    """msml_bn_fwd_ex: `bn3(out) + identity` of one residual unit hands the batch statistics of its output to the next
    unit's bn1 (ref iresnet.py:56-67).  Outputs, saved / running statistics and every gradient must equal the unchained
    sequence (same kernels, statistics pass included) up to the fp32 summation order of the slab partials."""
    need_gpu()
    from msml_b200 import ops
    torch.manual_seed(C + H)
    mk = lambda: torch.randn(B, C, H, H, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
    out0, skip0, d1, d2 = mk() * 1.7 + 0.3, mk(), mk(), mk()
    res = {}
    for chained in (False, True):
        torch.manual_seed(1)
        bn_a = torch.nn.BatchNorm2d(C, eps=1e-5).cuda().train()
        bn_b = torch.nn.BatchNorm2d(C, eps=1e-5).cuda().train()
        pr = torch.nn.PReLU(C).cuda()
        with torch.no_grad():
            for bn in (bn_a, bn_b):
                bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.5, 0.5)
            pr.weight.uniform_(0.1, 0.4)
        out, skip = out0.clone().requires_grad_(True), skip0.clone().requires_grad_(True)
        y1 = ops.bn_act(out, bn_a, None, skip, emit_next_stats=chained)
        assert hasattr(y1, ops._CHAIN_ATTR) == chained
        if consumer == "fork":
            y2, xs = ops.bn_act_fork(y1, bn_b)
            loss = (y2.float() * d1.float()).sum() + (xs.float() * d2.float()).sum()
        else:
            y2 = ops.bn_act(y1, bn_b, pr)
            loss = (y2.float() * d1.float()).sum()
        assert not hasattr(y1, ops._CHAIN_ATTR)                 # consumed exactly once
        loss.backward()
        res[chained] = dict(y1=host(y1), y2=host(y2), dout=host(out.grad), dskip=host(skip.grad), rm=host(bn_b.running_mean),
                            rv=host(bn_b.running_var), dg_a=host(bn_a.weight.grad), dg_b=host(bn_b.weight.grad),
                            db_b=host(bn_b.bias.grad), nbt=int(bn_b.num_batches_tracked))
    assert np.array_equal(res[True]["y1"], res[False]["y1"])    # the producer's own output does not change at all
    assert res[True]["nbt"] == res[False]["nbt"] == 1
    lo = dtype == torch.bfloat16
    assert_close(res[True]["rm"], res[False]["rm"], 1e-5, atol=1e-6, what="running_mean")
    assert_close(res[True]["rv"], res[False]["rv"], 1e-5, atol=1e-6, what="running_var")
    for k in ("y2", "dout", "dskip", "dg_a", "dg_b", "db_b"):
        assert_close(res[True][k], res[False][k], 1e-2 if lo else 2e-5, atol_frac=4e-3 if lo else 1e-5, what=k)


def test_bn_chained_statistics_are_dropped_when_stale():
    """The attached statistics describe one tensor at one version: an in-place write, eval mode or MSML_BN_CHAIN=0
    must fall back to reading the tensor."""
    need_gpu()
    from msml_b200 import ops
    torch.manual_seed(2)
    bn_a = torch.nn.BatchNorm2d(64).cuda().train()
    bn_b = torch.nn.BatchNorm2d(64).cuda().train()
    x = torch.randn(4, 64, 9, 9, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        y1 = ops.bn_act(x, bn_a, emit_next_stats=True)
        assert hasattr(y1, ops._CHAIN_ATTR)
        y1.mul_(3.0).add_(1.0)                                   # the statistics on y1 are now wrong
        y2 = ops.bn_act(y1, bn_b).float()
    assert y2.mean((0, 2, 3)).abs().max().item() < 2e-2
    assert (y2.var((0, 2, 3), unbiased=False) - 1).abs().max().item() < 3e-2
    with torch.no_grad():
        y1 = ops.bn_act(x, bn_a, emit_next_stats=True)
        bn_b.eval()
        want = torch.nn.functional.batch_norm(y1.float(), bn_b.running_mean, bn_b.running_var, bn_b.weight, bn_b.bias, False, 0.0, bn_b.eps)
        got = ops.bn_act(y1, bn_b).float()
    assert not hasattr(y1, ops._CHAIN_ATTR)
    assert_close(host(got), host(want), 2e-2, atol=2e-2, what="eval consumer ignores batch statistics")
    bn_a.eval()
    with torch.no_grad():
        assert not hasattr(ops.bn_act(x, bn_a, emit_next_stats=True), ops._CHAIN_ATTR)     # eval producer emits nothing


def test_bn_cooperative_single_launch_matches_default(tmp_path):
    """MSML_BN_FUSED=1 (one cooperative launch with two grid barriers, kept for the measured comparison) must compute the
    same BN forward / backward as the default three-launch path.  The mode is read once per process, so the
    cooperative run happens in a subprocess."""
    need_gpu()
    import os
    import subprocess
    import sys
    script = r'''
import sys, torch
sys.path.insert(0, %r)
from msml_b200 import ops
torch.manual_seed(7)
bn = torch.nn.BatchNorm2d(128).cuda().train(); pr = torch.nn.PReLU(128).cuda()
with torch.no_grad():
    bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.5, 0.5); pr.weight.uniform_(0.1, 0.4)
x = torch.randn(16, 128, 14, 14, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_()
r = torch.randn_like(x)
y = ops.bn_act(x, bn, pr, r)
y.backward(torch.randn_like(y))
torch.save({"y": y.detach().float().cpu(), "dx": x.grad.float().cpu(), "dg": bn.weight.grad.cpu(), "db": bn.bias.grad.cpu(),
            "dp": pr.weight.grad.cpu(), "rm": bn.running_mean.cpu(), "rv": bn.running_var.cpu()}, sys.argv[1])
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for mode in ("0", "1"):
        path = str(tmp_path / ("bn_%s.pt" % mode))
        env = dict(os.environ, MSML_BN_FUSED=mode)
        subprocess.run([sys.executable, "-c", script, path], check=True, env=env, timeout=300)
        outs[mode] = torch.load(path)
    for k in outs["0"]:
        assert_close(outs["1"][k].double().numpy(), outs["0"][k].double().numpy(), 2e-2 if k in ("y", "dx") else 1e-4,
                     atol_frac=2e-2 if k in ("y", "dx") else 1e-4, what="cooperative vs default: " + k)


@pytest.mark.parametrize("mode,act", [("given", "sigmoid"), ("invert", "sigmoid"), ("invert", "tanh")])
@pytest.mark.parametrize("with_teacher", [True, False])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fm_peer_mul_and_mse_match_autograd(mode, act, with_teacher, dtype):
    """K-P: ops.fm_peer_mul + ops.mse_loss against the reference's expressions (ref fmoperator.py:293-302: `m_bar * identity`,
    `m_bar * yt`, `MSELoss()(f_occ, f_out)`; `1 - x` over the activated mask for mask_trans 'invert', ref :160-166) run by
    torch autograd in fp64 on the same (rounded) inputs: values, the loss, and the gradients of the mask source and yf."""
    need_gpu()
    from msml_b200 import ops
    torch.manual_seed(11)
    B, C, H = 3, 64, 14
    mk = lambda: torch.randn(B, C, H, H, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
    src, yf, yt, w1, w2 = mk(), mk(), mk(), mk(), mk()
    src_a, yf_a = src.clone().requires_grad_(True), yf.clone().requires_grad_(True)
    out = ops.fm_peer_mul(src_a, yf_a, yt if with_teacher else None, mode, act)
    pf, pt = out if with_teacher else (out, None)
    l2 = ops.mse_loss(pt * 1.0, pf) if with_teacher else ops.mse_loss(pf, w2)
    assert l2.dtype == torch.float32 and l2.dim() == 0
    ((pf.float() * w1.float()).sum() + 3.0 * l2).backward()
    src_r, yf_r = src.double().requires_grad_(True), yf.double().requires_grad_(True)
    gate = src_r if mode == "given" else (torch.sigmoid(src_r) if act == "sigmoid" else torch.tanh(src_r))
    m_bar = gate if mode == "given" else 1 - gate
    pf_r = m_bar * yf_r
    rd = lambda t: t.to(dtype).double()                      # the kernels hand the products on in the storage dtype
    if with_teacher:
        pt_r = m_bar * yt.double()
        l2_r = torch.nn.MSELoss()(pt_r, pf_r)
    else:
        l2_r = torch.nn.MSELoss()(pf_r, w2.double())
    ((pf_r * w1.double()).sum() + 3.0 * l2_r).backward()
    lo = dtype == torch.bfloat16
    assert_close(host(pf), host(rd(pf_r)), 1e-2 if lo else 2e-5, atol=1e-2 if lo else 2e-6, what="pf")
    if with_teacher:
        assert_close(host(pt), host(rd(pt_r)), 1e-2 if lo else 2e-5, atol=1e-2 if lo else 2e-6, what="pt")
    assert abs(float(l2) - float(l2_r)) <= (1e-2 if lo else 2e-5) * abs(float(l2_r))
    assert_close(host(yf_a.grad), host(yf_r.grad), 3e-2 if lo else 1e-4, atol_frac=1e-2 if lo else 1e-5, what="dyf")
    assert_close(host(src_a.grad), host(src_r.grad), 3e-2 if lo else 1e-4, atol_frac=1e-2 if lo else 1e-5, what="dsrc")


def test_fm_peer_ops_reject_bad_arguments():
    need_gpu()
    from msml_b200 import ops
    a = torch.randn(2, 8, 4, 4, device="cuda")
    with pytest.raises(ValueError):
        ops.fm_peer_mul(a, a, None, "flip")
    with pytest.raises(ValueError):
        ops.fm_peer_mul(a, a[:, :4], None)
    with pytest.raises(ValueError):
        ops.mse_loss(a, a[:1])
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.mse_loss(a.cpu(), a.cpu())
    assert float(ops.mse_loss(a, a)) == 0.0


@pytest.mark.parametrize("name", ["fm_peer_c64_conv", "fm_peer_c128_invert"])
def test_fmcnn_peer_branch_matches_reference_golden(name):
    """FMCnn with the peer-guided branch on (use_ori=True; ref fmoperator.py:129-166,293-302,307-308), random
    deterministic weights, against vectors produced by the reference's FMCnn: output, distillation loss, dyf and
    parameter gradients in train mode; the eval call (yt=None) adds f_out without a loss."""
    need_gpu()
    from msml_b200.backbones.fm import FMCnn
    from oracle.detfill import fill_state_dict_
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        g = load_golden(name)
        C, H = int(g["C"]), int(g["H"])
        fm = FMCnn(H, H, C, kernel_size=3, resblocks=2, activation=str(g["act"]), arith_strategy=str(g["arith"]),
                   peer_params={"use_ori": True, "use_conv": bool(int(g["use_conv"])), "mask_trans": str(g["mask_trans"]),
                                "use_decoder": False})
        fill_state_dict_(fm)
        fm = fm.cuda().train()
        yf = dev(g["yf"]).requires_grad_(True)
        out, l2 = fm(yf, dev(g["yo"]), dev(g["yt"]))
        assert l2 is not None
        (out * dev(g["dout"])).sum().add(0.5 * l2).backward()
        assert_close(host(out), g["out"], 2e-4, atol_frac=2e-5, what="out")
        assert abs(float(l2) - float(g["l2"])) <= 1e-4 * abs(float(g["l2"])) + 1e-7
        assert_close(host(yf.grad), g["dyf"], 2e-3, atol_frac=2e-4, what="dyf")
        params = dict(fm.named_parameters())
        for k in g:
            if k.startswith("pgrad."):
                assert_close(host(params[k[6:]].grad), g[k], 2e-3, atol_frac=5e-4, what=k)
        ge = load_golden(name + "_eval")
        fm.eval()
        with torch.no_grad():
            out_e, l2_e = fm(dev(g["yf"]), dev(g["yo"]))
        assert l2_e is None
        assert_close(host(out_e), ge["out"], 2e-4, atol_frac=2e-5, what="eval out")
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
