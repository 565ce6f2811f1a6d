"""Which ATen kernels are left in the training step, and which line of this package launches them.
Runs the benchmark's step eagerly under torch.profiler (one step, shapes + python stacks) and prints the
non-convolution ATen ops by device time.  Diagnostic only (needs a GPU): python tools/diag_glue.py > gpurun_out/glue.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from msml_b200.backbones import MSML  # noqa: E402
from msml_b200.engine import TrainStep  # noqa: E402
from msml_b200.headers import ArcFace, PartialFC, PartialFCSGD  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(1)
    B = bench.BATCH
    net = MSML("iresnet50", "unet", (1, 1, 1, 1), bench.NUM_CLASSES, fp16=True, header_type=None, fm_params=bench.FM_PARAMS).to(dev).train()
    pfc = PartialFC(0, 0, 1, B, False, ArcFace(bench.S, bench.M), bench.NUM_CLASSES, sample_rate=1.0, embedding_size=512)
    lr = 0.1 * B / 512
    opt = torch.optim.SGD([p for p in net.parameters() if p.requires_grad], lr=lr, momentum=0.9, weight_decay=5e-4, fused=True)
    opt_pfc = PartialFCSGD(pfc, lr=lr, momentum=0.9, weight_decay=5e-4, emit_normalized=True, fuse_projection=True)
    step = TrainStep(net, pfc, opt, opt_pfc, (B, 3, 112, 112), world_size=1, max_norm=5.0, use_graph=False, wgrad_side_stream=False)
    img = torch.randn(B, 3, 112, 112, device=dev).contiguous(memory_format=torch.channels_last)
    lab = torch.randint(0, bench.NUM_CLASSES, (B,), device=dev)
    for _ in range(4):
        step(img, lab)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
        step(img, lab)
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages(group_by_input_shape=True, group_by_stack_n=12):
        t = getattr(e, "self_device_time_total", None)
        if t is None:
            t = getattr(e, "self_cuda_time_total", 0)
        if t <= 0 or not e.key.startswith("aten::") or "conv" in e.key or "cudnn" in e.key:
            continue
        stack = [f for f in (e.stack or []) if "msml_b200" in f or "bench.py" in f or "engine.py" in f]
        rows.append((t, e.count, e.key, str(e.input_shapes)[:160], " <- ".join(s.strip()[-90:] for s in stack[:3])))
    rows.sort(reverse=True)
    total = sum(r[0] for r in rows)
    print("# non-convolution ATen ops of one eager training step: self device us, calls, op, shapes, innermost package frames")
    print("# total %.1f us" % total)
    for t, n, k, sh, st in rows[:70]:
        print("%9.1f %4d %-28s %s\n            %s" % (t, n, k, sh, st))


if __name__ == "__main__":
    main()
