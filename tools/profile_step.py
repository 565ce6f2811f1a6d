#!/usr/bin/env python
"""Kernel-time breakdown of one training step (torch.profiler / CUPTI): where the 1-GPU step goes.
    python tools/profile_step.py [--batch 128] [--frb iresnet50] > gpurun_out/step_profile.txt"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from msml_b200.backbones import MSML  # noqa: E402
from msml_b200.headers import ArcFace, PartialFC  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--frb", default="iresnet50")
ap.add_argument("--classes", type=int, default=93431)
ap.add_argument("--rows", type=int, default=45)
ap.add_argument("--graph", action="store_true", help="profile CUDA-graph replays instead of eager steps")
args = ap.parse_args()

torch.backends.cudnn.benchmark = True
torch.manual_seed(1)
net = MSML(args.frb, "unet", (1, 1, 1, 1), args.classes, fp16=True, header_type=None, fm_params=(3, 2, "sigmoid", "mul")).cuda().train()
pfc = PartialFC(0, 0, 1, args.batch, False, ArcFace(), args.classes)
opt = torch.optim.SGD([p for p in net.parameters() if p.requires_grad], lr=0.02, momentum=0.9, weight_decay=5e-4, fused=True)
opt_pfc = torch.optim.SGD([{"params": pfc.parameters()}], lr=0.02, momentum=0.9, weight_decay=5e-4, fused=True)
img = torch.randn(args.batch, 3, 112, 112, device="cuda")
label = torch.randint(0, args.classes, (args.batch,), device="cuda")


from msml_b200.engine import TrainStep  # noqa: E402
_ts = TrainStep(net, pfc, opt, opt_pfc, (args.batch, 3, 112, 112), use_graph=args.graph)


def step():
    _ts(img, label)


for _ in range(4):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=args.rows, max_name_column_width=90))
ev = [e for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = sum(e.self_device_time_total for e in ev)
print("total CUDA kernel time for 2 steps: %.2f ms over %d kernel kinds, %d launches" % (tot / 1e3, len(ev), sum(e.count for e in ev)))
