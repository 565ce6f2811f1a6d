"""Which ATen ops (with input shapes) are left in one eager training step: guides the removal of layout / dtype copies."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from msml_b200.backbones import MSML
from msml_b200.engine import TrainStep
from msml_b200.headers import ArcFace, PartialFC

torch.backends.cudnn.benchmark = True
torch.manual_seed(1)
B, C = 128, 93431
net = MSML("iresnet50", "unet", (1, 1, 1, 1), C, fp16=True, header_type=None, fm_params=(3, 2, "sigmoid", "mul")).cuda().train()
pfc = PartialFC(0, 0, 1, B, False, ArcFace(), C)
opt = torch.optim.SGD([p for p in net.parameters() if p.requires_grad], lr=0.02, momentum=0.9, weight_decay=5e-4, fused=True)
opt_pfc = torch.optim.SGD([{"params": pfc.parameters()}], lr=0.02, momentum=0.9, weight_decay=5e-4, fused=True)
img = torch.randn(B, 3, 112, 112, device="cuda").contiguous(memory_format=torch.channels_last)
label = torch.randint(0, C, (B,), device="cuda")
ts = TrainStep(net, pfc, opt, opt_pfc, (B, 3, 112, 112), use_graph=False)
for _ in range(3):
    ts(img, label)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    ts(img, label)
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True) if e.self_device_time_total > 0 and e.key.startswith("aten::")]
rows.sort(key=lambda e: -e.self_device_time_total)
for e in rows[:70]:
    print("%8.1f us %4d  %-28s %s" % (e.self_device_time_total, e.count, e.key, str(e.input_shapes)[:150]))
