"""Phase timestamps of the fused BN kernels (CTA 0, %globaltimer): MSML_BN_SKIP_PHASES=8 python tools/bn_phases.py"""
import os, sys, ctypes
os.environ["MSML_BN_SKIP_PHASES"] = "8"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msml_b200 import _lib
from msml_b200._lib import load, check
lib = load()
B = 128
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
for C, H in [(64, 112), (64, 56), (128, 28), (256, 14), (512, 7)]:
    Pn = B * H * H
    x = torch.randn(Pn, C, device="cuda").to(torch.bfloat16); dy = torch.randn_like(x); y = torch.empty_like(x); dx = torch.empty_like(x)
    g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda"); a = torch.full((C,), 0.25, device="cuda")
    rm = torch.zeros(C, device="cuda"); rv = torch.ones(C, device="cuda"); st = torch.empty(2, C, device="cuda"); gr = torch.empty(3, C, device="cuda")
    wsb = lib.msml_bn_workspace(Pn, C); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    off = (592 * 3 * C + 592 + 3 * C) * 4
    for name in ("fwd", "bwd"):
        res = []
        for it in range(6):
            if name == "fwd":
                check(lib.msml_bn_fwd(P(x), None, P(y), P(g), P(b), P(a), P(rm), P(rv), None, P(st[0]), P(st[1]), Pn, C, 1, 1, 0.1, 1e-5, P(ws), wsb, s))
            else:
                check(lib.msml_bn_bwd(P(dy), P(x), None, P(g), P(b), P(a), P(st[0]), P(st[1]), P(dx), None, P(gr[0]), P(gr[1]), P(gr[2]), Pn, C, 1, 1, 0, P(ws), wsb, s))
            torch.cuda.synchronize()
            t = ws[off:off + 48].view(torch.int64).cpu().tolist()
            res.append([(t[i + 1] - t[i]) / 1e3 for i in range(5)])
        r = res[-1]
        print("C=%3d H=%3d %s  phase1 %5.1f  sync %5.1f  phase2 %5.1f  sync %5.1f  phase3 %5.1f  (us, CTA 0)" % (C, H, name, *r))
