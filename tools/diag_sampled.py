"""Where does a sampled PartialFC step (sample_rate < 1) spend its time?  Host wall clock and CUDA-event time of each
phase of `forward_backward -> opt.step -> update` at BASELINE config-4 per-rank shapes (ref headers/partial_fc.py:77-116).

    python tools/diag_sampled.py [--classes 125000] [--batch 1024] [--sample-rate 0.1] [--steps 8]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from msml_b200.headers import ArcFace, PartialFC  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--classes", type=int, default=1000000)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--sample-rate", type=float, default=0.1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--fused", action="store_true")
    a = ap.parse_args()
    torch.manual_seed(1)
    pfc = PartialFC(0, 0, 1, a.batch, False, ArcFace(64.0, 0.5), a.classes, sample_rate=a.sample_rate, embedding_size=512)
    if a.fused:
        from msml_b200.headers import PartialFCSGD
        opt = PartialFCSGD(pfc, lr=0.1, momentum=0.9, weight_decay=5e-4)
    else:
        opt = torch.optim.SGD([{"params": pfc.parameters()}], lr=0.1, momentum=0.9, weight_decay=5e-4)
    g = torch.Generator(device="cuda").manual_seed(1)
    feat = torch.nn.functional.normalize(torch.randn(a.batch, 512, device="cuda", generator=g))
    label = torch.randint(0, a.classes, (a.batch,), device="cuda", generator=g)
    rows = []
    for it in range(a.steps):
        torch.cuda.synchronize()
        rec = {}
        t0 = time.perf_counter()
        marks = [("start", torch.cuda.Event(enable_timing=True))]
        marks[0][1].record()

        def mark(name):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            rec["host_" + name] = round((time.perf_counter() - t0) * 1e3, 3)
            marks.append((name, e))

        pfc.forward_backward(label, feat, opt)
        mark("fwd_bwd")
        opt.step()
        mark("opt_step")
        pfc.update()
        mark("update")
        torch.cuda.synchronize()
        rec["host_total_synced"] = round((time.perf_counter() - t0) * 1e3, 3)
        for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
            rec["gpu_" + n1] = round(e0.elapsed_time(e1), 3)
        rec["reserved_MB"] = torch.cuda.memory_reserved() >> 20
        rec["n_cudaMalloc"] = torch.cuda.memory_stats()["num_device_alloc"]
        rows.append(rec)
    # finer split inside forward_backward, one extra step: prepare (side stream) vs the rest
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tl, (wn, inv) = pfc.prepare(label, opt)
    h_prep = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    prep_total = (time.perf_counter() - t0) * 1e3
    print(json.dumps({"args": vars(a), "steps": rows, "prepare_host_ms": round(h_prep, 3), "prepare_synced_ms": round(prep_total, 3)}))


if __name__ == "__main__":
    main()
