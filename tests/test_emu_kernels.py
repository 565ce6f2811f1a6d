"""CUDA kernels executed under a CPU emulation of the CUDA execution model (tests/emu/cuda_emu.h): a LOGIC check that
needs no GPU.  The kernel header is compiled for the host with g++; every CTA runs as blockDim.x OS threads with real
barriers and warp-shuffle exchanges.  It does not replace the GPU parity tests (no memory-model, alignment or
performance coverage); it exists so that kernels written without GPU access are not shipped unexecuted.

Covered: csrc/seg_loss_kernels.cuh (consensus segmentation loss) against the reference goldens
(tests/golden/consensus.npz) and the oracle; csrc/pfc_sgd_kernels.cuh (fused PartialFC SGD) against the
reference recipe gather -> torch.optim.SGD -> scatter (ref headers/partial_fc.py:93-94,101-104,112-114) on CPU;
csrc/fm_cat_kernels.cuh (FM concat) — a kernel that IS verified on a B200 with the same assertions
(tests/test_gpu_fusion.py::test_fm_cat_matches_concat), run here to cross-check the emulation itself;
csrc/bn_act_kernels.cuh (fused BatchNorm + residual + PReLU, also verified on a B200) against oracle/bn_act.py;
csrc/fm_gate_kernels.cuh (K-A, the north-star mask-fusion tail, verified on a B200) against the reference goldens and
oracle/fm_tail.py; csrc/pfc_sample_kernels.cuh (K-D: remap, radix select, searchsorted, row gather / scatter — the
bit-exact integer path of PartialFC.sample, verified on a B200) against the reference's recorded draws and the oracle.
"""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import load_golden
from oracle import bn_act as obn
from oracle import consensus
from oracle import dap as odap
from oracle import fm_tail
from oracle import margins as omarg
from oracle import partial_fc as opfc

HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_INC = "/usr/local/cuda/include"
c_p, c_i64, c_int, c_f = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float
F32, BF16 = 0, 1


def build_emu(tmp_path_factory, source):
    if shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    out = str(tmp_path_factory.mktemp("emu") / ("lib" + source.replace(".cpp", ".so")))
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-shared", "-fPIC", "-pthread", "-I" + CUDA_INC,
                        os.path.join(HERE, "emu", source), "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return ctypes.CDLL(out)


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    lib = build_emu(tmp_path_factory, "emu_seg_loss.cpp")
    lib.emu_consensus_workspace.restype = ctypes.c_size_t
    lib.emu_consensus_workspace.argtypes = [c_i64] * 4
    lib.emu_consensus_fwd.argtypes = [c_p, c_p, c_p] + [c_i64] * 4 + [c_int, c_int, c_f, c_f, c_int, c_int, c_p, c_p, c_p]
    lib.emu_consensus_bwd.argtypes = [c_p] * 5 + [c_i64] * 4 + [c_int, c_int]
    lib.emu_last_error.restype = ctypes.c_char_p
    return lib


def to_bf16_bits(a):
    """fp32 -> bf16 bit patterns (round to nearest even), as uint16."""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    return (((u + 0x7FFF + ((u >> 16) & 1)) >> 16) & 0xFFFF).astype(np.uint16)


def from_bf16_bits(b):
    return (b.astype(np.uint32) << 16).view(np.float32)


def run(lib, logit, blobs, target, K, alpha=10.0, beta=5.0, rp="idx", rkl="idx", cl=False, gout=1.0, dtype=F32, expect_rc=0):
    N, C, H, W = logit.shape
    z = np.ascontiguousarray(logit.transpose(0, 2, 3, 1) if cl else logit, np.float32)
    zbuf = to_bf16_bits(z) if dtype == BF16 else z
    b = np.ascontiguousarray(np.asarray(blobs).reshape(N, H * W), np.int64)
    t = np.ascontiguousarray(np.asarray(target).reshape(N, H * W), np.int64)
    ws = np.zeros(lib.emu_consensus_workspace(N, C, H * W, K) // 4 + 1, np.float32)
    loss = np.zeros(1, np.float32)
    coef = np.zeros(2 * K * N * C, np.float32)
    rc = lib.emu_consensus_fwd(zbuf.ctypes.data, b.ctypes.data, t.ctypes.data, N, C, H * W, K, int(cl), dtype, alpha, beta,
                               int(rp == "all"), int(rkl == "all"), loss.ctypes.data, coef.ctypes.data, ws.ctypes.data)
    if expect_rc is None:
        return rc
    assert rc == expect_rc, lib.emu_last_error()
    dbuf = np.zeros_like(zbuf)
    g = np.array([gout], np.float32)
    assert lib.emu_consensus_bwd(zbuf.ctypes.data, b.ctypes.data, coef.ctypes.data, g.ctypes.data, dbuf.ctypes.data, N, C, H * W, K,
                                 int(cl), dtype) == 0
    dz = from_bf16_bits(dbuf) if dtype == BF16 else dbuf
    return float(loss[0]), (dz.transpose(0, 3, 1, 2) if cl else dz)


@pytest.mark.parametrize("name,cl", [("binary_missing", False), ("binary_missing", True), ("four_blobs", False), ("four_blobs", True),
                                     ("four_blobs_all_all", False), ("four_blobs_idx_all", True), ("underflow", False),
                                     ("underflow", True), ("seg_shape", False)])
def test_consensus_kernels_match_reference_golden(emu, name, cl):
    g = load_golden("consensus")
    alpha, beta, rp, rkl = [str(v) for v in g[name + ".cfg"]]
    blobs = g[name + ".blobs"].astype(np.int64)
    ids, dense = np.unique(blobs, return_inverse=True)                  # the kernels take ids 0 .. K-1
    loss, dz = run(emu, g[name + ".logit"], dense.reshape(blobs.shape), g[name + ".target"], len(ids), float(alpha), float(beta), rp, rkl, cl)
    want = float(g[name + ".loss"])
    assert abs(loss - want) <= 2e-6 * abs(want), (loss, want)
    gd = g[name + ".dlogit"]
    np.testing.assert_allclose(dz, gd, rtol=2e-4, atol=2e-6 * np.abs(gd).max())


def test_consensus_kernels_bf16_three_classes_ignore_and_gout(emu):
    rng = np.random.default_rng(3)
    N, C, H, W = 3, 3, 33, 41                                           # 1353 pixels: a ragged second chunk
    z = from_bf16_bits(to_bf16_bits(rng.normal(size=(N, C, H, W)).astype(np.float32))).reshape(N, C, H, W)
    blobs = rng.integers(0, 3, size=(N, H, W))
    blobs[2][blobs[2] == 1] = 2                                         # blob 1 is missing from the last sample
    target = np.array([2, 0, 1])[blobs]
    want, dwant = consensus.consensus_loss(z, blobs, target, 10.0, 5.0)
    loss, dz = run(emu, z, blobs, target, 3, gout=2.5, dtype=BF16)
    assert abs(loss - want) <= 1e-5 * abs(want)
    np.testing.assert_allclose(dz, 2.5 * dwant, rtol=1e-2, atol=1e-2 * np.abs(dwant).max())      # the gradient is rounded to bf16
    # pixels of blob 2 marked -1 (member of no blob) == the loss over blobs {0, 1} only; K larger than the ids present is fine
    ign = np.where(blobs == 2, -1, blobs)
    want2, dwant2 = consensus.consensus_loss(z, ign, target, 10.0, 5.0, ids=[0, 1])
    loss2, dz2 = run(emu, z, ign, target, 5)
    assert abs(loss2 - want2) <= 2e-6 * abs(want2)
    np.testing.assert_allclose(dz2, dwant2, rtol=2e-4, atol=2e-6 * np.abs(dwant2).max())
    assert not dz2[np.broadcast_to((ign == -1)[:, None], dz2.shape)].any()


def test_consensus_kernels_poison_and_argument_errors(emu):
    rng = np.random.default_rng(4)
    z = rng.normal(size=(2, 2, 8, 8)).astype(np.float32)
    msk = (rng.random((2, 8, 8)) < 0.4).astype(np.int64)
    loss, _ = run(emu, z, msk, msk, 2)
    assert np.isfinite(loss)
    bad = msk.copy(); bad[0, 0, 0] = 7
    assert np.isnan(run(emu, z, bad, msk, 2)[0])                        # id outside [0, K)
    lab = msk.copy(); lab[1, 3, 3] = 1 - lab[1, 3, 3]
    assert np.isnan(run(emu, z, msk, lab, 2)[0])                        # labels differ inside a blob (ref :103 asserts)
    assert np.isnan(run(emu, z, msk, msk * 5, 2)[0])                    # label outside [0, C)
    assert run(emu, np.zeros((2, 7, 8, 8), np.float32), msk, msk, 2, expect_rc=None) != 0      # C > 4
    assert b"classes" in emu.emu_last_error()
    assert run(emu, z, msk, msk, 40, expect_rc=None) != 0               # K > 32


# ------------------------------------------------------------------------------------------------ fused PartialFC SGD
@pytest.fixture(scope="module")
def emu_sgd(tmp_path_factory):
    lib = build_emu(tmp_path_factory, "emu_pfc_sgd.cpp")
    lib.emu_pfc_sgd_update.argtypes = [c_p, c_p, c_p, c_p, c_i64, c_i64, c_i64, c_p, c_f, c_f, c_f, c_f, c_int, c_p, c_p]
    lib.emu_sgd_last_error.restype = ctypes.c_char_p
    return lib


@pytest.mark.parametrize("D", [128, 512])
@pytest.mark.parametrize("sampled", [False, True])
@pytest.mark.parametrize("nesterov,dampening", [(False, 0.0), (True, 0.0), (False, 0.3)])
def test_pfc_sgd_kernel_matches_reference_recipe(emu_sgd, D, sampled, nesterov, dampening):
    """gather weight[index] / weight_mom[index] -> stock SGD with the supplied momentum buffer -> scatter back
    (ref partial_fc.py:93-94,112-114,101-104 + train.py:188-191,299-300) vs ONE in-place pass of the kernel."""
    import torch
    torch.manual_seed(0)
    num_local = 50
    n_s = 21 if sampled else num_local
    W = torch.randn(num_local, D) * 0.01
    M = torch.randn(num_local, D) * 0.001
    index = torch.sort(torch.randperm(num_local)[:n_s]).values if sampled else torch.arange(num_local)
    dw = torch.randn(n_s, D) * 0.1
    sub = torch.nn.Parameter(W[index].clone())
    subm = M[index].clone()
    opt = torch.optim.SGD([sub], lr=0.1, momentum=0.9, weight_decay=5e-4, dampening=dampening, nesterov=nesterov)
    opt.state[sub]["momentum_buffer"] = subm
    sub.grad = dw.clone()
    opt.step()
    Wref, Mref = W.clone(), M.clone()
    Wref[index] = sub.data
    Mref[index] = subm
    w, m, g = W.numpy().copy(), M.numpy().copy(), dw.numpy().copy()
    idx = index.numpy().copy()
    wn = np.zeros((n_s, D), np.uint16)
    inv = np.zeros(n_s, np.float32)
    lr = np.array([0.1], np.float32)                                    # read through the device-scalar path
    rc = emu_sgd.emu_pfc_sgd_update(w.ctypes.data, m.ctypes.data, g.ctypes.data, idx.ctypes.data if sampled else None, n_s, num_local, D,
                                    lr.ctypes.data, 0.0, 0.9, 5e-4, dampening, int(nesterov), wn.ctypes.data, inv.ctypes.data)
    assert rc == 0, emu_sgd.emu_sgd_last_error()
    np.testing.assert_allclose(w, Wref.numpy(), rtol=2e-6, atol=1e-8)      # fma vs mul + add: a few ulps of the 1e-2 scale
    np.testing.assert_allclose(m, Mref.numpy(), rtol=2e-6, atol=1e-8)
    if sampled:                                                         # rows outside the sample are untouched, bit for bit
        rest = np.setdiff1d(np.arange(num_local), idx)
        assert np.array_equal(w[rest], W.numpy()[rest]) and np.array_equal(m[rest], M.numpy()[rest])
    want_wn = torch.nn.functional.normalize(torch.from_numpy(w[idx])).to(torch.bfloat16).float().numpy()
    assert np.abs(from_bf16_bits(wn) - want_wn).max() <= 2.0 ** -8 * np.abs(want_wn).max()      # at most one bf16 ulp
    np.testing.assert_allclose(inv, 1.0 / np.linalg.norm(w[idx], axis=1), rtol=1e-6)


def test_pfc_sgd_kernel_skips_bad_rows_and_rejects_bad_shapes(emu_sgd):
    w = np.ones((4, 128), np.float32)
    m = np.zeros((4, 128), np.float32)
    g = np.ones((2, 128), np.float32)
    idx = np.array([1, 9], np.int64)                                    # 9 is outside the shard: skipped, never dereferenced
    assert emu_sgd.emu_pfc_sgd_update(w.ctypes.data, m.ctypes.data, g.ctypes.data, idx.ctypes.data, 2, 4, 128, None, 0.5, 0.0, 0.0, 0.0, 0,
                                      None, None) == 0
    assert np.allclose(w[1], 0.5) and np.array_equal(w[[0, 2, 3]], np.ones((3, 128), np.float32))
    assert emu_sgd.emu_pfc_sgd_update(w.ctypes.data, m.ctypes.data, g.ctypes.data, None, 2, 4, 100, None, 0.5, 0.0, 0.0, 0.0, 0, None, None) != 0
    assert b"multiples of 128" in emu_sgd.emu_sgd_last_error()
    assert emu_sgd.emu_pfc_sgd_update(w.ctypes.data, m.ctypes.data, g.ctypes.data, None, 9, 4, 128, None, 0.5, 0.0, 0.0, 0.0, 0, None, None) != 0


# ------------------------------------------------------------------------------------------------ FM concat (GPU-verified kernel)
@pytest.fixture(scope="module")
def emu_cat(tmp_path_factory):
    lib = build_emu(tmp_path_factory, "emu_fm_cat.cpp")
    lib.emu_fm_cat_fwd.argtypes = [c_p, c_p, c_p, c_i64, c_i64, c_i64, c_i64, c_int, c_int]
    lib.emu_fm_cat_bwd.argtypes = [c_p, c_p, c_p, c_p, c_i64, c_i64, c_i64, c_i64, c_int, c_int]
    return lib


@pytest.mark.parametrize("C,Co,P", [(64, 18, 3 * 9 * 7), (128, 18, 75), (8, 3, 36), (16, 8, 45), (64, 18, 5000)])
@pytest.mark.parametrize("dtype", [F32, BF16])
def test_fm_cat_kernels_match_concat(emu_cat, C, Co, P, dtype):
    """Same assertions as the GPU test: the forward is a pure copy (bit-exact vs numpy), the backward is the column slice
    plus the tail's gradient summed in fp32 and rounded once; sms=1 forces several trips of the grid-stride loops."""
    rng = np.random.default_rng(11)
    Ct = -(-(C + Co) // 8) * 8
    rnd = lambda *shape: from_bf16_bits(to_bf16_bits(rng.normal(size=shape).astype(np.float32))).reshape(shape)   # bf16-representable
    yf, yo, dcat, dtail = rnd(P, C), rnd(P, Co), rnd(P, Ct), rnd(P, C)
    enc = (lambda a: to_bf16_bits(a).reshape(a.shape)) if dtype == BF16 else (lambda a: np.ascontiguousarray(a, np.float32))
    dec = (lambda b: from_bf16_bits(b).reshape(b.shape)) if dtype == BF16 else (lambda b: b)
    yf_b, yo_b, dcat_b, dtail_b = enc(yf), enc(yo), enc(dcat), enc(dtail)
    cat_b = np.full_like(enc(np.zeros((P, Ct), np.float32)), 0x7F if dtype == BF16 else 7)       # poisoned: every element must be written
    assert emu_cat.emu_fm_cat_fwd(yf_b.ctypes.data, yo_b.ctypes.data, cat_b.ctypes.data, P, C, Co, Ct, dtype, 1) == 0
    want = np.concatenate([yf, yo, np.zeros((P, Ct - C - Co), np.float32)], axis=1)
    assert np.array_equal(dec(cat_b), want)
    dyf_b, dyo_b = np.zeros_like(yf_b), np.zeros_like(yo_b)
    assert emu_cat.emu_fm_cat_bwd(dcat_b.ctypes.data, dtail_b.ctypes.data, dyf_b.ctypes.data, dyo_b.ctypes.data, P, C, Co, Ct, dtype, 1) == 0
    want_dyf = dcat[:, :C] + dtail
    if dtype == BF16:
        want_dyf = from_bf16_bits(to_bf16_bits(want_dyf)).reshape(P, C)
    assert np.array_equal(dec(dyf_b), want_dyf)
    assert np.array_equal(dec(dyo_b), dcat[:, C:C + Co])
    dyf2 = np.zeros_like(yf_b)
    assert emu_cat.emu_fm_cat_bwd(dcat_b.ctypes.data, None, dyf2.ctypes.data, None, P, C, Co, Ct, dtype, 2) == 0
    assert np.array_equal(dec(dyf2), dcat[:, :C])


# ------------------------------------------------------------------------------------------------ memcheck / racecheck on CPU
SANITIZERS = {
    "asan": (["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"], "--seed-oob", "AddressSanitizer"),
    "tsan": (["-fsanitize=thread"], "--seed-race", "ThreadSanitizer: data race"),
}


@pytest.fixture(scope="module")
def sanitizer_builds(tmp_path_factory):
    """Both sanitizer builds of tests/emu/sanitize_main.cpp, compiled side by side."""
    if shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    d = tmp_path_factory.mktemp("sanitize")
    procs = {}
    for tag, (flags, _, _) in SANITIZERS.items():
        exe = str(d / tag)
        procs[tag] = (exe, subprocess.Popen(["g++", "-std=c++20", "-O1", "-pthread", "-I" + CUDA_INC] + flags +      # add -g to symbolise a report
                                            [os.path.join(HERE, "emu", "sanitize_main.cpp"), "-o", exe],
                                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    out = {}
    for tag, (exe, p) in procs.items():
        _, err = p.communicate()
        out[tag] = (exe, p.returncode, err)
    return out


@pytest.mark.parametrize("tag", ["asan", "tsan"])
def test_emulated_kernels_are_clean_under_sanitizers(sanitizer_builds, tag):
    """tests/emu/sanitize_main.cpp runs the consensus, PartialFC-SGD, FM-concat, BatchNorm and mask-fusion kernels on
    exact-size heap buffers with real threads per CTA: AddressSanitizer plays compute-sanitizer's memcheck, ThreadSanitizer
    its racecheck.  The same binary with a seeded defect must be REPORTED, otherwise a clean run would mean nothing."""
    exe, rc, err = sanitizer_builds[tag]
    _, seed, needle = SANITIZERS[tag]
    if rc != 0 and ("cannot find" in err or "unrecognized" in err):
        pytest.skip("sanitizer runtime not installed")
    assert rc == 0, err[-3000:]
    env = dict(os.environ, TSAN_OPTIONS="suppressions=" + os.path.join(HERE, "emu", "tsan.supp"))     # one examined, benign report
    clean = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=env)
    assert clean.returncode == 0 and "Sanitizer" not in clean.stderr, clean.stderr[-3000:]
    assert "ran to completion, rc=0" in clean.stdout
    seeded = subprocess.run([exe, seed], capture_output=True, text=True, timeout=600)
    assert needle in seeded.stderr and seeded.returncode != 0
    if tag == "asan":       # UBSan's alignment check: a 128-bit access through a 4-byte-aligned pointer (cudaErrorMisalignedAddress on a GPU)
        mis = subprocess.run([exe, "--seed-misaligned"], capture_output=True, text=True, timeout=600)
        assert "misaligned address" in mis.stderr and mis.returncode != 0


# ------------------------------------------------------------------------------------------------ fused BN (+res) (+PReLU) (GPU-verified)
kBN_MAX_CTAS = 148 * 4          # csrc/bn_act_kernels.cuh kBnMaxCtas
@pytest.fixture(scope="module")
def emu_bn(tmp_path_factory):
    lib = build_emu(tmp_path_factory, "emu_bn.cpp")
    lib.emu_bn_fwd.argtypes = [c_p] * 11 + [c_i64, c_i64, c_int, c_f, c_f, c_int, c_int]
    lib.emu_bn_bwd.argtypes = [c_p] * 14 + [c_i64, c_i64, c_int, c_int, c_int, c_int, c_int]
    lib.emu_bn_fwd_chain.argtypes = [c_p] * 14 + [c_i64, c_i64, c_int, c_f, c_f, c_int, c_int, c_int]
    return lib


@pytest.mark.parametrize("P,C,G1,G3a,G3b,prelu,dtype", [
    (162, 32, 3, 5, 4, False, BF16), (401, 64, 7, 4, 9, True, BF16), (98, 256, 2, 2, 3, False, BF16),
    (5, 16, 9, 7, 2, True, BF16),                # empty producer slabs: zero-weight partials
    (2, 8, 1, 1, 1, False, BF16), (37, 4, 2, 3, 2, True, F32),
    (700, 64, 9, kBN_MAX_CTAS, 5, False, BF16),      # producer grid = the whole partial stride
])
def test_bn_chained_statistics(emu_bn, P, C, G1, G3a, G3b, prelu, dtype):
    """msml_bn_fwd_ex: the apply pass of `bn_a(x) + res` (ref iresnet.py:66-67, the end of one residual unit) also emits the
    slab statistics of its output, and the next unit's `bn1` (iresnet.py:57) starts at the merge.  The second op must agree
    with the oracle run on the first op's ROUNDED output, statistics included, although it never read that tensor for them;
    the consumer workspace starts out as NaN, so slots behind the producer's grid must carry zero weight."""
    rng = np.random.default_rng(7 * P + C)
    q = (lambda a: from_bf16_bits(to_bf16_bits(a)).reshape(a.shape)) if dtype == BF16 else (lambda a: a.astype(np.float32))
    enc = (lambda t: to_bf16_bits(t).reshape(t.shape)) if dtype == BF16 else (lambda t: np.ascontiguousarray(t, np.float32))
    dec = (lambda b: from_bf16_bits(b).reshape(b.shape)) if dtype == BF16 else (lambda b: b)
    ptr = lambda t: t.ctypes.data if t is not None else None
    x = q(rng.normal(-0.5, 1.5, size=(P, C)).astype(np.float32))
    r = q(rng.normal(0.7, 1.0, size=(P, C)).astype(np.float32))
    ga, ba = rng.uniform(0.5, 1.5, C).astype(np.float32), rng.uniform(-0.5, 0.5, C).astype(np.float32)
    gb, bb = rng.uniform(0.5, 1.5, C).astype(np.float32), rng.uniform(-0.5, 0.5, C).astype(np.float32)
    a = rng.uniform(0.1, 0.4, C).astype(np.float32) if prelu else None
    rm0, rv0 = rng.normal(size=C).astype(np.float32), rng.uniform(0.5, 2.0, C).astype(np.float32)
    xb, rb = enc(x), enc(r)
    y1b, y2b = np.zeros_like(xb), np.zeros_like(xb)
    rm, rv, nbt = rm0.copy(), rv0.copy(), np.array([2], np.int64)
    mean, invstd = np.zeros(C, np.float32), np.zeros(C, np.float32)
    assert emu_bn.emu_bn_fwd_chain(ptr(xb), ptr(rb), ptr(y1b), ptr(y2b), ptr(ga), ptr(ba), ptr(gb), ptr(bb), ptr(a), ptr(rm), ptr(rv),
                                   ptr(nbt), ptr(mean), ptr(invstd), P, C, dtype, 0.1, 1e-5, G1, G3a, G3b) == 0
    y1_want, _ = obn.bn_act_fwd(x, ga, ba, None, r, True, np.zeros(C, np.float32), np.ones(C, np.float32))
    tol = dict(rtol=2e-2, atol=2e-2) if dtype == BF16 else dict(rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(dec(y1b), y1_want, **tol)
    y1 = dec(y1b).astype(np.float32)                                    # what the unchained second op would have read
    y2_want, st = obn.bn_act_fwd(y1, gb, bb, a, None, True, rm0, rv0)
    np.testing.assert_allclose(mean, st["mean"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(invstd, st["invstd"], rtol=2e-5)
    np.testing.assert_allclose(rm, st["running_mean"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(rv, st["running_var"], rtol=2e-5)
    np.testing.assert_allclose(dec(y2b), y2_want, **tol)
    assert nbt[0] == 3


@pytest.mark.parametrize("P,C,G1,G3,prelu,res,dtype", [
    (162, 32, 3, 5, False, False, F32), (162, 32, 3, 5, True, False, BF16), (162, 32, 3, 5, False, True, BF16), (162, 32, 3, 5, True, True, F32),
    (401, 64, 7, 4, True, True, BF16), (401, 64, 7, 4, False, False, BF16),
    (98, 256, 2, 2, True, True, BF16),           # 32 channel vectors per row: the warp-per-slot fold; one finalize CTA per channel
    (5, 16, 9, 7, True, True, BF16),             # more CTAs than rows: empty slabs must merge as zero-weight partials
    (2, 8, 1, 1, False, False, BF16),            # one channel vector per row, two rows
    (3, 4, 2, 3, True, False, F32),              # fp32, one vector per row
])
def test_bn_kernels_match_oracle(emu_bn, P, C, G1, G3, prelu, res, dtype):
    """The three-launch forward and backward (slab statistics -> per-channel finalize -> apply) on slabs that do not
    divide the rows evenly, vs the numpy fp64 oracle (ref iresnet.py:56-67 / fmoperator.py:52-68 semantics)."""
    rng = np.random.default_rng(P + C)
    q = (lambda a: from_bf16_bits(to_bf16_bits(a)).reshape(a.shape)) if dtype == BF16 else (lambda a: a.astype(np.float32))
    x = q(rng.normal(1.0, 2.0, size=(P, C)).astype(np.float32))
    r = q(rng.normal(size=(P, C)).astype(np.float32)) if res else None
    dy = q(rng.normal(size=(P, C)).astype(np.float32))
    dadd = q(rng.normal(size=(P, C)).astype(np.float32))
    gamma = rng.uniform(0.5, 1.5, C).astype(np.float32)
    beta = rng.uniform(-0.5, 0.5, C).astype(np.float32)
    a = rng.uniform(0.1, 0.4, C).astype(np.float32) if prelu else None
    rm0, rv0 = rng.normal(size=C).astype(np.float32), rng.uniform(0.5, 2.0, C).astype(np.float32)
    enc = (lambda t: to_bf16_bits(t).reshape(t.shape)) if dtype == BF16 else (lambda t: np.ascontiguousarray(t, np.float32))
    dec = (lambda b: from_bf16_bits(b).reshape(b.shape)) if dtype == BF16 else (lambda b: b)
    ptr = lambda t: t.ctypes.data if t is not None else None
    xb, rb, dyb, daddb = enc(x), (enc(r) if res else None), enc(dy), enc(dadd)
    yb = np.zeros_like(xb)
    rm, rv, nbt = rm0.copy(), rv0.copy(), np.array([4], np.int64)
    mean, invstd = np.zeros(C, np.float32), np.zeros(C, np.float32)
    assert emu_bn.emu_bn_fwd(ptr(xb), ptr(rb), ptr(yb), ptr(gamma), ptr(beta), ptr(a), ptr(rm), ptr(rv), ptr(nbt), ptr(mean), ptr(invstd),
                             P, C, dtype, 0.1, 1e-5, G1, G3) == 0
    y_want, st = obn.bn_act_fwd(x, gamma, beta, a, r, True, rm0, rv0)
    tol = dict(rtol=2e-2, atol=2e-2) if dtype == BF16 else dict(rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(dec(yb), y_want, **tol)
    np.testing.assert_allclose(mean, st["mean"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(invstd, st["invstd"], rtol=1e-5)
    np.testing.assert_allclose(rm, st["running_mean"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(rv, st["running_var"], rtol=1e-5)
    assert nbt[0] == 5
    both = prelu and res
    dxb, dresb = np.zeros_like(xb), (np.zeros_like(xb) if both else None)
    grads = np.full((3, C), 0.25, np.float32)                          # accumulate mode adds on top of what is there
    assert emu_bn.emu_bn_bwd(ptr(dyb), ptr(xb), ptr(rb) if both else None, ptr(gamma), ptr(beta), ptr(a), ptr(mean), ptr(invstd), ptr(dxb),
                             ptr(dresb), ptr(daddb), ptr(grads[0]), ptr(grads[1]), ptr(grads[2]) if prelu else None, P, C, dtype, 1, 1,
                             G1, G3) == 0
    want = obn.bn_act_bwd(dy, x, gamma, beta, a, r, True)
    gt = dict(rtol=3e-2, atol=3e-2) if dtype == BF16 else dict(rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(dec(dxb), want["dx"] + dadd, **gt)
    if both:
        np.testing.assert_allclose(dec(dresb), want["dres"], **gt)
    scale = np.abs(want["dgamma"]).max() + 1.0
    np.testing.assert_allclose(grads[0] - 0.25, want["dgamma"], rtol=1e-3, atol=1e-4 * scale)
    np.testing.assert_allclose(grads[1] - 0.25, want["dbeta"], rtol=1e-3, atol=1e-4 * scale)
    if prelu:
        np.testing.assert_allclose(grads[2] - 0.25, want["dprelu"], rtol=1e-3, atol=1e-4 * scale)


# ------------------------------------------------------------------------------------------------ K-A mask-fusion tail (GPU-verified)
ACTS = {"tanh": 0, "sigmoid": 1}
ARITHS = {"add": 0, "sub": 1, "div": 2, "mul": 3}


@pytest.fixture(scope="module")
def emu_gate(tmp_path_factory):
    lib = build_emu(tmp_path_factory, "emu_fm_gate.cpp")
    lib.emu_fm_gate_fwd_multi.argtypes = [c_int, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int]
    lib.emu_fm_gate_bwd_multi.argtypes = [c_int, c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int]
    return lib


def gate_run(lib, yfs, zs, douts, act, arith, dtype, sms, fouts=None):
    """One forward and one backward launch over len(yfs) segments -> (outs, dyfs, dzs) as fp32 arrays."""
    enc = (lambda t: to_bf16_bits(t).reshape(t.shape)) if dtype == BF16 else (lambda t: np.ascontiguousarray(t, np.float32))
    dec = (lambda b: from_bf16_bits(b).reshape(b.shape)) if dtype == BF16 else (lambda b: b)
    n = len(yfs)
    arr = lambda bufs: (c_p * n)(*[b.ctypes.data for b in bufs])
    yb, zb, db = [enc(t) for t in yfs], [enc(t) for t in zs], [enc(t) for t in douts]
    fb = [enc(t) for t in fouts] if fouts is not None else None
    ob, dyb, dzb = [np.zeros_like(b) for b in yb], [np.zeros_like(b) for b in yb], [np.zeros_like(b) for b in yb]
    sizes = (c_i64 * n)(*[b.size for b in yb])
    assert lib.emu_fm_gate_fwd_multi(n, arr(yb), arr(zb), arr(fb) if fb else None, arr(ob), sizes, dtype, ACTS[act], ARITHS[arith], sms) == 0
    assert lib.emu_fm_gate_bwd_multi(n, arr(db), arr(yb), arr(zb), arr(dyb), arr(dzb), sizes, dtype, ACTS[act], ARITHS[arith], sms) == 0
    return [dec(b) for b in ob], [dec(b) for b in dyb], [dec(b) for b in dzb]


@pytest.mark.parametrize("name", ["fm_c64_sigmoid_mul", "fm_c128_tanh_add", "fm_c64_sigmoid_div", "fm_c256_tanh_sub"])
def test_fm_gate_kernels_match_reference_golden(emu_gate, name):
    """ref backbones/fm/fmoperator.py:288,304-310 run by make_golden.py: out, the direct dyf and dz (fp32 mode)."""
    g = load_golden(name)
    act, arith = str(g["act"]), str(g["arith"])
    outs, dyfs, dzs = gate_run(emu_gate, [g["yf"]], [g["z"]], [g["dout"]], act, arith, F32, sms=1)
    np.testing.assert_allclose(outs[0], g["out"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dyfs[0], g["dyf_direct"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(dzs[0], g["dz"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_fm_gate_kernels_four_scales_one_launch_with_ragged_tails(emu_gate, dtype):
    """Four segments of very different sizes through ONE launch (CTAs dealt in proportion), sizes that are not whole vectors,
    the peer term f_out, all on 2 'SMs' so that the persistent loops take several trips; vs oracle/fm_tail.py."""
    rng = np.random.default_rng(8)
    q = (lambda a: from_bf16_bits(to_bf16_bits(a)).reshape(a.shape)) if dtype == BF16 else (lambda a: a)
    sizes = [64 * 7 * 7 * 3 + 5, 128 * 5 * 5, 19, 256 * 9 + 3]
    mk = lambda: [q(rng.normal(size=n).astype(np.float32)) for n in sizes]
    yfs, zs, ds, fos = mk(), mk(), mk(), mk()
    tol = dict(rtol=2e-2, atol=2e-2) if dtype == BF16 else dict(rtol=2e-5, atol=2e-5)
    for act, arith in (("sigmoid", "mul"), ("tanh", "add"), ("sigmoid", "sub")):
        outs, dyfs, dzs = gate_run(emu_gate, yfs, zs, ds, act, arith, dtype, sms=2)
        for i in range(4):
            np.testing.assert_allclose(outs[i], fm_tail.fm_gate_fwd(yfs[i], zs[i], act, arith), **tol)
            wdyf, wdz = fm_tail.fm_gate_bwd(ds[i], yfs[i], zs[i], act, arith)
            np.testing.assert_allclose(dyfs[i], wdyf, **tol)
            np.testing.assert_allclose(dzs[i], wdz, **tol)
    outs, _, _ = gate_run(emu_gate, yfs[:1], zs[:1], ds[:1], "sigmoid", "mul", dtype, sms=1, fouts=fos[:1])
    np.testing.assert_allclose(outs[0], fm_tail.fm_gate_fwd(yfs[0], zs[0], "sigmoid", "mul") + fos[0], **tol)


# ------------------------------------------------------------------------------------------------ K-D PartialFC sampling (GPU-verified)
@pytest.fixture(scope="module")
def emu_pfc(tmp_path_factory):
    lib = build_emu(tmp_path_factory, "emu_pfc_sample.cpp")
    lib.emu_pfc_remap.argtypes = [c_p, c_i64, c_i64, c_i64]
    lib.emu_pfc_mark_positive.argtypes = [c_p, c_p, c_i64, c_i64]
    lib.emu_pfc_select.argtypes = [c_p, c_i64, c_i64, c_p, c_p]
    lib.emu_pfc_searchsorted.argtypes = [c_p, c_i64, c_p, c_p]
    lib.emu_gather_rows.argtypes = [c_p, c_p, c_p, c_i64, c_i64]
    lib.emu_scatter_rows.argtypes = [c_p, c_p, c_p, c_i64, c_i64]
    for f in ("emu_pfc_remap", "emu_pfc_mark_positive", "emu_pfc_select", "emu_pfc_searchsorted", "emu_gather_rows", "emu_scatter_rows"):
        getattr(lib, f).restype = None
    return lib


def emu_sample(lib, total_label, perm, class_start, num_local, num_sample):
    """PartialFC.sample through the kernels (ref :77-94): remap -> perm[positive] = 2 -> select -> searchsorted."""
    tl = np.ascontiguousarray(total_label, np.int64).copy()
    lib.emu_pfc_remap(tl.ctypes.data, tl.size, class_start, num_local)
    remapped = tl.copy()
    p = np.ascontiguousarray(perm, np.float32).copy()
    lib.emu_pfc_mark_positive(p.ctypes.data, tl.ctypes.data, tl.size, num_local)
    index = np.full(max(num_sample, tl.size, 1), -7, np.int64)
    n_index = np.zeros(1, np.int64)
    lib.emu_pfc_select(p.ctypes.data, num_local, num_sample, index.ctypes.data, n_index.ctypes.data)
    index = index[:int(n_index[0])].copy()
    lib.emu_pfc_searchsorted(tl.ctypes.data, tl.size, index.ctypes.data, n_index.ctypes.data)
    return remapped, index, tl


@pytest.mark.parametrize("name,W", [("pfc_w1_sample", 1), ("pfc_w2_sample", 2), ("pfc_w1_overflow", 1)])
def test_pfc_sampling_kernels_bit_exact_vs_reference_draws(emu_pfc, name, W):
    """The torch.rand draws the reference consumed (recorded by make_golden.py) -> the SAME sampled class indices, bit for bit
    (ref :84-90, including the branch where the positives outnumber num_sample)."""
    g = load_golden(name)
    C, B, sr = int(g["C"]), int(g["B"]), float(g["sample_rate"])
    for step in range(int(g["steps"])):
        total = np.concatenate([g[f"r{r}.s{step}.label"] for r in range(W)])
        for rank in range(W):
            num_local, class_start, num_sample = opfc.shard_geometry(C, W, rank, sr)
            perm = g[f"r{rank}.s{step}.perm"]
            if perm.size == 0:                                  # the reference took the n_pos > num_sample branch without drawing
                perm = np.zeros(num_local, np.float32)
            remapped, index, tl = emu_sample(emu_pfc, total, perm, class_start, num_local, num_sample)
            assert np.array_equal(index, g[f"r{rank}.s{step}.index"]), (name, rank, step)
            want_tl, want_index = opfc.sample(total, perm, class_start, num_local, num_sample, sr)
            assert np.array_equal(index, want_index) and np.array_equal(tl, want_tl)
            assert np.array_equal(remapped, opfc.remap_labels(total, class_start, num_local))


@pytest.mark.parametrize("num_local,sr,n_labels", [(1000, 0.3, 64), (11679, 0.5, 1024), (4097, 0.1, 700), (8192, 0.999, 16)])
def test_pfc_sampling_kernels_bit_exact_vs_oracle(emu_pfc, num_local, sr, n_labels):
    """Several CTAs (4096 keys each), ragged last tile, ties at the threshold (quantised draws): lowest class index wins."""
    rng = np.random.default_rng(num_local)
    class_start = 5000
    num_sample = int(sr * num_local)
    total = rng.integers(0, class_start + 2 * num_local, n_labels)
    perm = (np.floor(rng.random(num_local) * 257) / 257).astype(np.float32)        # many exact ties
    remapped, index, tl = emu_sample(emu_pfc, total, perm, class_start, num_local, num_sample)
    want_tl, want_index = opfc.sample(total, perm, class_start, num_local, num_sample, sr)
    assert index.size == want_index.size and np.array_equal(index, want_index)
    assert np.array_equal(tl, want_tl)
    # gather the sampled rows and scatter them back (ref :93-94, :101-104)
    D = 64
    w = rng.normal(size=(num_local, D)).astype(np.float32)
    sub = np.zeros((index.size, D), np.float32)
    emu_pfc.emu_gather_rows(w.ctypes.data, index.ctypes.data, sub.ctypes.data, index.size, D)
    assert np.array_equal(sub, w[index])
    w2 = np.zeros_like(w)
    emu_pfc.emu_scatter_rows(w2.ctypes.data, index.ctypes.data, sub.ctypes.data, index.size, D)
    want = np.zeros_like(w)
    want[index] = w[index]
    assert np.array_equal(w2, want)


# ------------------------------------------------------------------------------------------------ K-B DAP + argmax mask (GPU-verified)
@pytest.fixture(scope="module")
def emu_dap(tmp_path_factory):
    lib = build_emu(tmp_path_factory, "emu_dap.cpp")
    lib.emu_dap_fwd.argtypes = [c_p, c_p, c_p, c_i64, c_i64, c_i64, c_i64, c_int, c_int, c_int]
    lib.emu_dap_bwd.argtypes = [c_p, c_p, c_i64, c_i64, c_i64, c_i64, c_int, c_int, c_int]
    lib.emu_dap_fwd.restype = lib.emu_dap_bwd.restype = None
    return lib


@pytest.mark.parametrize("cl", [False, True])
def test_dap_kernels_match_reference_golden(emu_dap, cl):
    """ref backbones/osb/unet.py:158-161,223 + train.py:357 run by make_golden.py (ties planted): y, dx and the argmax mask
    bit for bit (first index on ties)."""
    g = load_golden("dap")
    x, dy = g["x"], g["dy"]
    B, CK, H, W = x.shape
    G, kk = dy.shape[1], CK // dy.shape[1]
    lay = (lambda a: np.ascontiguousarray(a.transpose(0, 2, 3, 1))) if cl else (lambda a: np.ascontiguousarray(a))
    unlay = (lambda a, C: a.reshape(B, H, W, C).transpose(0, 3, 1, 2)) if cl else (lambda a, C: a.reshape(B, C, H, W))
    xb, dyb = lay(x.astype(np.float32)), lay(dy.astype(np.float32))
    yb, dxb = np.zeros(B * G * H * W, np.float32), np.zeros(B * CK * H * W, np.float32)
    mask = np.full(B * H * W, -1, np.int64)
    emu_dap.emu_dap_fwd(xb.ctypes.data, yb.ctypes.data, mask.ctypes.data, B, G, kk, H * W, int(cl), F32, 2)
    emu_dap.emu_dap_bwd(dyb.ctypes.data, dxb.ctypes.data, B, G, kk, H * W, int(cl), F32, 2)
    np.testing.assert_allclose(unlay(yb, G), g["y"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(unlay(dxb, CK), g["dx"], rtol=1e-6, atol=1e-7)
    assert np.array_equal(mask.reshape(B, H, W), g["mask"])
    assert np.array_equal(mask.reshape(B, H, W), odap.argmax_mask(unlay(yb, G)))


# ------------------------------------------------------------------------------------------------ head: margin math, statistics merge, weight normalisation
@pytest.fixture(scope="module")
def emu_head(tmp_path_factory):
    lib = build_emu(tmp_path_factory, "emu_head_small.cpp")
    lib.emu_wnorm_cast.argtypes = [c_p, c_p, c_p, c_i64, c_i64, c_int]
    lib.emu_head_local_stats.argtypes = [c_p, c_p, c_p, c_p, c_int, c_int, c_p]
    lib.emu_head_merge_stats.argtypes = [c_p, c_int, c_int, c_p, c_p]
    lib.emu_margin_fwd.argtypes = [c_p, c_p, c_i64, c_i64, c_i64, c_int, c_f, c_f, c_f, c_f]
    lib.emu_margin_bwd.argtypes = [c_p, c_p, c_p, c_i64, c_i64, c_i64, c_int, c_f, c_f, c_f, c_f]
    for f in ("emu_wnorm_cast", "emu_head_local_stats", "emu_head_merge_stats", "emu_margin_fwd", "emu_margin_bwd"):
        getattr(lib, f).restype = None
    return lib


@pytest.mark.parametrize("tag", ["arc_p", "arc_f", "arc_am_p", "arc_am_f", "cos_p", "cos_f", "cos_am_p", "cos_am_f"])
def test_margin_kernels_match_reference_golden(emu_head, tag):
    """ref headers/margin_losses.py:390-418 (AMArcFace) / :275-303 (AMCosFace) run by make_golden.py on the reference's own
    6 x 8 fixture (labels with -1) and random inputs, k != 0 included: logits from the cosine matrix, and d logits -> d cos."""
    g = load_golden("margins")
    kind = 0 if tag.startswith("arc") else 1
    s_, m_, a_, k_ = [float(v) for v in g[tag + ".smak"]]
    cos = omarg.l2_normalize(g[tag + ".e"].astype(np.float64)) @ omarg.l2_normalize(g[tag + ".w"].astype(np.float64)).T
    label = np.ascontiguousarray(g[tag + ".label"], np.int64)
    B, C = cos.shape
    buf = np.ascontiguousarray(cos, np.float32)
    emu_head.emu_margin_fwd(buf.ctypes.data, label.ctypes.data, B, C, C, kind, s_, m_, a_, k_)
    np.testing.assert_allclose(buf, g[tag + ".logits"], rtol=2e-5, atol=2e-5 * s_)
    dl = np.array(g[tag + ".dl"], np.float32)                   # a copy: the kernel works in place
    cos32 = np.ascontiguousarray(cos, np.float32)
    emu_head.emu_margin_bwd(dl.ctypes.data, cos32.ctypes.data, label.ctypes.data, B, C, C, kind, s_, m_, a_, k_)
    want = g[tag + ".dl"].astype(np.float64) * omarg.margin_dcos(cos, label, "arc" if kind == 0 else "cos", s_, m_, a_, k_)
    np.testing.assert_allclose(dl, want, rtol=2e-4, atol=1e-4 * np.abs(want).max())


def test_head_statistics_merge_and_loss(emu_head):
    """Per-tile partials -> per-rank (max, sum, target logit) -> merge over W ranks -> loss = -mean log p_target
    (ref partial_fc.py:135-144,159-163: three all-reduces there, one gathered (W, 3, B_tot) array here)."""
    rng = np.random.default_rng(9)
    W, B_tot, n_blocks = 3, 37, 11
    log2e = 1.4426950408889634
    logits = [rng.normal(0, 8, size=(B_tot, n_blocks * 16)) for _ in range(W)]               # per rank, 16 classes per tile
    owner = rng.integers(0, W, B_tot)                                                        # the rank holding each row's target class
    tcol = rng.integers(0, n_blocks * 16, B_tot)
    gathered = np.zeros((W, 3, B_tot), np.float32)
    for r in range(W):
        tiles = logits[r].reshape(B_tot, n_blocks, 16)
        pmax = np.ascontiguousarray((tiles.max(axis=2) * log2e).T, np.float32)               # (n_blocks, B_tot), log2 units
        psum = np.ascontiguousarray(np.exp2(tiles * log2e - tiles.max(axis=2, keepdims=True) * log2e).sum(axis=2).T, np.float32)
        tl = np.where(owner == r, tcol, -1).astype(np.int64)
        tgt = logits[r][np.arange(B_tot), tcol].astype(np.float32)
        stats = np.zeros((3, B_tot), np.float32)
        emu_head.emu_head_local_stats(pmax.ctypes.data, psum.ctypes.data, tgt.ctypes.data,
                                      tl.ctypes.data, n_blocks, B_tot, stats.ctypes.data)
        np.testing.assert_allclose(stats[0], logits[r].max(axis=1), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(stats[1], np.exp(logits[r] - logits[r].max(axis=1, keepdims=True)).sum(axis=1), rtol=1e-4)
        assert np.array_equal(np.isinf(stats[2]), owner != r)
        gathered[r] = stats
    gstats = np.zeros((2, B_tot), np.float32)
    loss = np.zeros(1, np.float32)
    emu_head.emu_head_merge_stats(gathered.ctypes.data, W, B_tot, gstats.ctypes.data, loss.ctypes.data)
    full = np.concatenate(logits, axis=1)
    gmax = full.max(axis=1)
    gsum = np.exp(full - gmax[:, None]).sum(axis=1)
    p_t = np.exp(np.array([logits[owner[i]][i, tcol[i]] for i in range(B_tot)]) - gmax) / gsum
    np.testing.assert_allclose(gstats[0], gmax, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(gstats[1], gsum, rtol=1e-4)
    want = -np.mean(np.log(np.maximum(p_t, 1e-30)))
    assert abs(float(loss[0]) - want) <= 1e-4 * abs(want)


def test_weight_normalise_cast_kernel(emu_head):
    """ref partial_fc.py:115 F.normalize(sub_weight) -> unit-norm bf16 rows + 1 / max(||w||, 1e-12), a zero row included."""
    rng = np.random.default_rng(10)
    n, D = 21, 512
    w = (rng.normal(size=(n, D)) * 0.01).astype(np.float32)
    w[5] = 0.0
    wn = np.zeros((n, D), np.uint16)
    inv = np.zeros(n, np.float32)
    emu_head.emu_wnorm_cast(w.ctypes.data, wn.ctypes.data, inv.ctypes.data, n, D, 1)
    want = omarg.l2_normalize(w.astype(np.float64))
    got = from_bf16_bits(wn).reshape(n, D)
    assert np.abs(got - want).max() <= 2.0 ** -8 * np.abs(want).max()
    np.testing.assert_allclose(inv, 1.0 / np.maximum(np.linalg.norm(w.astype(np.float64), axis=1), 1e-12), rtol=1e-5)
    assert not got[5].any()


# ------------------------------------------------------------------------------------------------ K-A extension: resized / broadcast mask
@pytest.fixture(scope="module")
def emu_mask(tmp_path_factory):
    lib = build_emu(tmp_path_factory, "emu_fm_mask.cpp")
    lib.emu_fm_mask.argtypes = [c_p] * 6 + [c_i64] * 7 + [c_int, c_int]
    return lib


@pytest.mark.parametrize("Cm,Hm,Wm", [(1, 8, 6), (1, 4, 3), (64, 8, 6), (64, 4, 3)])
@pytest.mark.parametrize("dtype,sigmoid_mul", [(F32, True), (BF16, True), (F32, False)])
def test_fm_mask_kernels_match_oracle(emu_mask, Cm, Hm, Wm, dtype, sigmoid_mul):
    """north_star: "mask logits resized to each feature scale, normalised into gates and multiplied into the feature maps, forward
    and backward" — nearest resize (2x and ragged), single-channel broadcast or per-channel mask, dMask reduced over the channels
    and over the resize fan-out; vs oracle/fm_tail.py."""
    rng = np.random.default_rng(Cm + Hm)
    B, H, W, C = 2, 8, 6, 64
    q = (lambda a: from_bf16_bits(to_bf16_bits(a)).reshape(a.shape)) if dtype == BF16 else (lambda a: a)
    yf = q(rng.normal(size=(B, H, W, C)).astype(np.float32))
    m = q(rng.normal(size=(B, Hm, Wm, Cm)).astype(np.float32))
    dout = q(rng.normal(size=(B, H, W, C)).astype(np.float32))
    enc = (lambda t: to_bf16_bits(t).reshape(t.shape)) if dtype == BF16 else (lambda t: np.ascontiguousarray(t, np.float32))
    dec = (lambda b: from_bf16_bits(b).reshape(b.shape)) if dtype == BF16 else (lambda b: b)
    yb, mb, db = enc(yf), enc(m), enc(dout)
    ob, dyb = np.zeros_like(yb), np.zeros_like(yb)
    dm = np.zeros((B, Hm, Wm, Cm), np.float32)                          # zero-filled by the caller (C-ABI contract)
    assert emu_mask.emu_fm_mask(db.ctypes.data, yb.ctypes.data, mb.ctypes.data, ob.ctypes.data, dyb.ctypes.data, dm.ctypes.data,
                                B, H, W, C, Hm, Wm, Cm, dtype, int(sigmoid_mul)) == 0
    act, arith = ("sigmoid", "mul") if sigmoid_mul else ("tanh", "add")
    tol = dict(rtol=2e-2, atol=2e-2) if dtype == BF16 else dict(rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(dec(ob), fm_tail.fm_mask_fwd(yf, m, act, arith), **tol)
    wdyf, wdm = fm_tail.fm_mask_bwd(dout, yf, m, act, arith)
    np.testing.assert_allclose(dec(dyb), wdyf, **tol)
    np.testing.assert_allclose(dm, wdm, rtol=1e-4, atol=1e-4 * np.abs(wdm).max())


# ------------------------------------------------------------------------------------------------ K-P peer-branch products + MSE
@pytest.fixture(scope="module")
def emu_peer(tmp_path_factory):
    lib = build_emu(tmp_path_factory, "emu_fm_peer.cpp")
    lib.emu_fm_peer_mul_fwd.argtypes = [c_p] * 5 + [c_i64, c_int, c_int, c_int, c_int]
    lib.emu_fm_peer_mul_bwd.argtypes = [c_p] * 7 + [c_i64, c_int, c_int, c_int, c_int]
    lib.emu_mse.argtypes = [c_p, c_p, c_i64, c_int, c_p, c_f, c_p, c_p, c_int]
    return lib


@pytest.mark.parametrize("n,blocks,dtype,mode,act,has_t", [
    (4096, 2, BF16, 0, "sigmoid", True), (4099, 3, BF16, 1, "sigmoid", True), (1003, 1, F32, 1, "tanh", True),
    (8 * 700 + 5, 4, BF16, 1, "tanh", False), (7, 2, F32, 0, "sigmoid", False), (2048, 9, F32, 0, "sigmoid", True),
])
def test_fm_peer_mul_kernels(emu_peer, n, blocks, dtype, mode, act, has_t):
    """pf = m_bar*yf, pt = m_bar*yt (ref fmoperator.py:295-299) with m_bar given or formed as 1 - act(z) (ref :160-166), and the
    backward dm_bar = dpf*yf + dpt*yt, dyf = dpf*m_bar, against fp64 numpy; ragged tails and more CTAs than work."""
    rng = np.random.default_rng(n + blocks)
    q = (lambda a: from_bf16_bits(to_bf16_bits(a)).reshape(a.shape)) if dtype == BF16 else (lambda a: a.astype(np.float32))
    enc = (lambda t: to_bf16_bits(t).reshape(t.shape)) if dtype == BF16 else (lambda t: np.ascontiguousarray(t, np.float32))
    dec = (lambda b: from_bf16_bits(b).reshape(b.shape)) if dtype == BF16 else (lambda b: b)
    ptr = lambda t: t.ctypes.data if t is not None else None
    src, yf, yt, dpf, dpt = (q(rng.normal(size=n).astype(np.float32)) for _ in range(5))
    sb, fb, tb, gfb, gtb = enc(src), enc(yf), (enc(yt) if has_t else None), enc(dpf), (enc(dpt) if has_t else None)
    pfb, ptb = np.zeros_like(fb), (np.zeros_like(fb) if has_t else None)
    assert emu_peer.emu_fm_peer_mul_fwd(ptr(sb), ptr(fb), ptr(tb), ptr(pfb), ptr(ptb), n, dtype, mode, ACTS[act], blocks) == 0
    s64 = src.astype(np.float64)
    if mode == 0:
        m, dm_dsrc = s64, np.ones(n)
    else:
        gate = 1 / (1 + np.exp(-s64)) if act == "sigmoid" else np.tanh(s64)
        m = 1 - gate
        dm_dsrc = -(gate * (1 - gate) if act == "sigmoid" else 1 - gate * gate)
    tol = dict(rtol=1e-2, atol=1e-2) if dtype == BF16 else dict(rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(dec(pfb), m * yf, **tol)
    if has_t:
        np.testing.assert_allclose(dec(ptb), m * yt, **tol)
    dsb, dyb = np.zeros_like(fb), np.zeros_like(fb)
    assert emu_peer.emu_fm_peer_mul_bwd(ptr(gfb), ptr(gtb), ptr(sb), ptr(fb), ptr(tb), ptr(dsb), ptr(dyb), n, dtype, mode, ACTS[act], blocks) == 0
    dm = dpf.astype(np.float64) * yf + (dpt.astype(np.float64) * yt if has_t else 0.0)
    gt = dict(rtol=2e-2, atol=2e-2) if dtype == BF16 else dict(rtol=5e-5, atol=5e-6)
    np.testing.assert_allclose(dec(dsb), dm * dm_dsrc, **gt)
    np.testing.assert_allclose(dec(dyb), dpf * m, **gt)


@pytest.mark.parametrize("n,blocks,dtype,both", [(5000, 3, BF16, True), (8 * 256 * 5 + 3, 7, BF16, False), (13, 4, F32, True), (40000, 148 * 8, F32, True)])
def test_mse_kernels(emu_peer, n, blocks, dtype, both):
    """mean((a-b)^2) with fp32 accumulation from the storage dtype (ref fmoperator.py:300 under autocast) and its gradient
    2(a-b)/n * g; unused partial slots start as NaN and must not be read."""
    rng = np.random.default_rng(n)
    q = (lambda a: from_bf16_bits(to_bf16_bits(a)).reshape(a.shape)) if dtype == BF16 else (lambda a: a.astype(np.float32))
    enc = (lambda t: to_bf16_bits(t).reshape(t.shape)) if dtype == BF16 else (lambda t: np.ascontiguousarray(t, np.float32))
    dec = (lambda b: from_bf16_bits(b).reshape(b.shape)) if dtype == BF16 else (lambda b: b)
    a, b = q(rng.normal(size=n).astype(np.float32)), q(rng.normal(0.3, 1.0, size=n).astype(np.float32))
    ab, bb = enc(a), enc(b)
    out = np.zeros(1, np.float32)
    da, db = np.zeros_like(ab), (np.zeros_like(ab) if both else None)
    assert emu_peer.emu_mse(ab.ctypes.data, bb.ctypes.data, n, dtype, out.ctypes.data, 0.7, da.ctypes.data, db.ctypes.data if both else None, blocks) == 0
    d = a.astype(np.float64) - b
    assert abs(out[0] - (d * d).mean()) <= 2e-6 * (d * d).mean()
    want = 2 * d / n * 0.7
    tol = dict(rtol=1e-2, atol=1e-2 * np.abs(want).max()) if dtype == BF16 else dict(rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(dec(da), want, **tol)
    if both:
        np.testing.assert_allclose(dec(db), -want, **tol)


# ------------------------------------------------------------------------------------------------ flat momentum SGD (+ bf16 shadow)
@pytest.mark.parametrize("n,blocks,momentum,wd,nesterov,scale", [
    (4096, 3, 0.9, 5e-4, 0, 1.7), (4 * 1031, 2, 0.9, 5e-4, 1, None), (8, 5, 0.0, 0.0, 0, 0.5), (4 * 5000, 1, 0.5, 1e-2, 0, 3.0),
])
def test_sgd_flat_kernel(tmp_path_factory, n, blocks, momentum, wd, nesterov, scale):
    """msml_sgd_flat == torch.optim.SGD's update rule (ref train.py:186-191, 299; dampening 0) written out in fp64, over
    three steps (the momentum buffer starts at zero: the first step is m = g), with the GradScaler-style division and the
    bf16 shadow of the new weights."""
    lib = build_emu(tmp_path_factory, "emu_sgd_flat.cpp")
    lib.emu_sgd_flat.argtypes = [c_p, c_p, c_p, c_p, c_i64, c_f, c_p, c_f, c_f, c_int, c_int]
    rng = np.random.default_rng(n + blocks)
    w = rng.normal(size=n).astype(np.float32)
    m = np.zeros(n, np.float32)
    shadow = np.zeros(n, np.uint16)
    w64, m64 = w.astype(np.float64), np.zeros(n)
    sc = np.array([scale], np.float32) if scale is not None else None
    lr = 0.05
    for step in range(3):
        g = rng.normal(size=n).astype(np.float32)
        assert lib.emu_sgd_flat(w.ctypes.data, m.ctypes.data, g.ctypes.data, shadow.ctypes.data, n, lr, sc.ctypes.data if sc is not None else None,
                                momentum, wd, nesterov, blocks) == 0
        g64 = g.astype(np.float64) / (float(sc[0]) if sc is not None else 1.0) + wd * w64
        if momentum != 0.0:
            m64 = momentum * m64 + g64
            g64 = g64 + momentum * m64 if nesterov else m64
        w64 = w64 - lr * g64
        np.testing.assert_allclose(w, w64, rtol=2e-6, atol=2e-6)
        if momentum != 0.0:
            np.testing.assert_allclose(m, m64, rtol=2e-6, atol=2e-6)
        assert np.array_equal(shadow, to_bf16_bits(w).reshape(-1))          # bit-exact round-to-nearest-even of the new weights
        w64, m64 = w.astype(np.float64), m.astype(np.float64)                # follow the fp32 trajectory


# ------------------------------------------------------------------------------------------------ multi-tensor gradient accumulate
def test_accum_bf16_multi_kernel(tmp_path_factory):
    """dst_f32 += float(src_bf16) over many tensors in one launch: sizes around the 8192-element block, ragged tails, a tensor whose
    fp32 view is only 4-byte aligned (the scalar path), more tensors than one launch takes."""
    lib = build_emu(tmp_path_factory, "emu_accum.cpp")
    lib.emu_accum_bf16_multi.argtypes = [c_int, c_p, c_p, c_p]
    rng = np.random.default_rng(12)
    sizes = [8192, 8193, 5, 16384 + 7, 1, 300] * 17                      # 102 tensors: two launches (96 segments each)
    flat = rng.normal(size=sum(sizes) + len(sizes) * 4 + 1).astype(np.float32)
    want = flat.copy()
    dsts, srcs, keep = [], [], []
    off = 1                                                              # the first view starts 4 bytes into the buffer
    for n in sizes:
        g = to_bf16_bits(rng.normal(size=n).astype(np.float32))
        keep.append(g)
        dsts.append(flat[off:off + n].ctypes.data)
        srcs.append(g.ctypes.data)
        want[off:off + n] += from_bf16_bits(g)
        off += (n + 3) // 4 * 4
    k = len(sizes)
    assert lib.emu_accum_bf16_multi(k, (c_p * k)(*dsts), (c_p * k)(*srcs), (c_i64 * k)(*sizes)) == 0
    assert np.array_equal(flat, want)
