"""Micro-benchmark of the fused BN(+res)(+PReLU) kernels at the layer shapes of the ires50_msml step (B=128, bf16 NHWC).
CUDA-graph replays of N calls (no Python / launch overhead in the numbers); rotates over several input sets.
MSML_BN_SKIP_PHASES=1|2|4 skips the work of a phase (debug attribution)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch import nn
from msml_b200 import ops

B = int(os.environ.get("B", 128))
shapes = [(64, 112), (64, 56), (128, 28), (256, 14), (512, 7), (32, 56), (128, 14)]
N = 12
print("skip mask:", os.environ.get("MSML_BN_SKIP_PHASES", "0"))


def timed_graph(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(); fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 5 / N * 1e3       # us per call


for C, H in shapes:
    for prelu, res in ((False, False), (True, False), (False, True)):
        bn = nn.BatchNorm2d(C).cuda().train()
        pr = nn.PReLU(C).cuda() if prelu else None
        nset = 3
        xs = [torch.randn(B, C, H, H, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_() for _ in range(nset)]
        rs = [torch.randn_like(x) for x in xs] if res else [None] * nset
        dys = [torch.randn_like(x) for x in xs]

        def fwd_only():
            with torch.no_grad():
                for i in range(N):
                    ops.bn_act(xs[i % nset], bn, pr, rs[i % nset])

        def fwd_bwd():
            for i in range(N):
                ops.bn_act(xs[i % nset], bn, pr, rs[i % nset]).backward(dys[i % nset])
                xs[i % nset].grad = None

        tf = timed_graph(fwd_only)
        tb = timed_graph(fwd_bwd) - tf
        n = B * C * H * H * 2
        fb = n * (3 if res else 2); bb = n * (3 + (2 if (res and prelu) else 0))
        print("C=%3d H=%3d prelu=%d res=%d  %6.1f MB  fwd %6.1f us (%5.0f GB/s)  bwd %6.1f us (%5.0f GB/s)" % (
            C, H, prelu, res, n / 1e6, tf, fb / tf / 1e3, tb, bb / tb / 1e3))
