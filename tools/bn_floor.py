"""Floor of a streaming kernel at BN layer sizes inside a CUDA graph: eval-mode BN (coef kernel + one elementwise pass), torch mul."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch import nn
from msml_b200 import ops
N = 12
def timed_graph(fn):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(); fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 5 / N * 1e3
for C, H in [(64, 56), (128, 28), (256, 14), (512, 7), (128, 14)]:
    bn = nn.BatchNorm2d(C).cuda(); pr = nn.PReLU(C).cuda()
    xs = [torch.randn(128, C, H, H, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last) for _ in range(3)]
    def ev():
        bn.eval()
        with torch.no_grad():
            for i in range(N): ops.bn_act(xs[i % 3], bn, pr)
    def tr():
        bn.train()
        with torch.no_grad():
            for i in range(N): ops.bn_act(xs[i % 3], bn, pr)
    def mul():
        with torch.no_grad():
            for i in range(N): xs[i % 3] * 2
    def amax():
        with torch.no_grad():
            for i in range(N): xs[i % 3].amax()
    print("C=%3d H=%3d %5.1f MB: eval BN (2 kernels) %5.1f us | train BN %5.1f us | torch mul %5.1f us | torch amax %5.1f us" % (
        C, H, xs[0].numel() * 2 / 1e6, timed_graph(ev), timed_graph(tr), timed_graph(mul), timed_graph(amax)))
