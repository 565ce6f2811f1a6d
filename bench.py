#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native MSML hot path.

    python bench.py --gpus N --steps K --warmup W            (N=1; N>1 under torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W     CPU port of the reference step
    python bench.py --workload fusion|head ...               micro-benchmarks (BASELINE configs 2 / 4)
    python bench.py --workload aux                           FM concat, consensus loss, fused PartialFC SGD, each alone

Default workload = BASELINE.json config 3: ires50_msml + PartialFC(ArcFace s=64 m=0.5, 93,431
classes, sample_rate 1) bf16 training step, 112x112 synthetic images, batch 128 per GPU.
One step = backbone forward -> F.normalize -> PartialFC.forward_backward -> features.backward(x_grad)
-> clip_grad_norm_(5) -> SGD steps -> pfc.update()   (ref train.py:283-300, the PartialFC variant).

Prints ONE JSON line (rank 0).  `value` = imgs/s with inputs resident in HBM; `e2e` = the same step
fed from pinned host memory each step with the loss read back each step.  `roofline.achieved` is measured
live: every launch of this library is bracketed by CUDA events on its own stream (msml_profile_*), work =
algorithmic bytes / flops of that launch.  `roofline.traffic` is NOT measured in this run: it is the DRAM byte count of
that kernel from the committed `ncu --set full` capture named in `traffic_source`.  Before anything is timed, every
rank runs one head step on seeded embeddings and rank 0 checks the loss it reports against the fp64 oracle evaluated
on the gathered inputs (`head_check`): the class-sharded path is verified on the very ranks that are benchmarked.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NUM_CLASSES = 93431
BATCH = 128
S, M = 64.0, 0.5
FM_PARAMS = (3, 2, "sigmoid", "mul")
FEAT_ELEMS_PER_IMG = 376320       # 64*56^2 + 128*28^2 + 256*14^2 + 512*7^2  


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx = [], set(), None
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------ CPU port
def cpu_step_factory(batch, frb="iresnet50", threads=None):
    """The reference's step restated on CPU in torch fp32, all host threads (oracle/model_cpu.py + oracle/partial_fc_torch.py:
    the same ATen calls the reference issues, ref train.py:283-300 and headers/partial_fc.py:96-177)."""
    import torch
    from msml_b200.backbones import MSML
    from oracle import model_cpu, partial_fc_torch as pt
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(1)
    net = MSML(frb, "unet", (1, 1, 1, 1), NUM_CLASSES, header_type=None, fm_params=FM_PARAMS)   # weights only (CPU init)
    sd = model_cpu.trainable_state(net)
    params = [v for v in sd.values() if v.requires_grad]
    opt = torch.optim.SGD(params, lr=0.1 * batch / 512, momentum=0.9, weight_decay=5e-4)
    weight = torch.normal(0, 0.01, (NUM_CLASSES, 512))
    mom = torch.zeros_like(weight)
    split = {"backbone_s": 0.0, "head_s": 0.0}

    def step():
        t0 = time.perf_counter()
        img = torch.randn(batch, 3, 112, 112)
        label = torch.randint(0, NUM_CLASSES, (batch,))
        feat, _seg = model_cpu.msml_forward(sd, img, frb, training=True, fm_params=FM_PARAMS)
        featn = torch.nn.functional.normalize(feat)
        t1 = time.perf_counter()
        x_grad, w_grad, loss = pt.head_step(featn, label, weight, "arc", S, M)      # PartialFC.forward_backward
        t2 = time.perf_counter()
        featn.backward(x_grad)
        torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 5)
        opt.step()
        opt.zero_grad()
        t3 = time.perf_counter()
        pt.sgd_update(weight, mom, w_grad)                                          # opt_pfc.step() (sample_rate 1: update() is a no-op)
        t4 = time.perf_counter()
        split["backbone_s"] += (t1 - t0) + (t3 - t2)
        split["head_s"] += (t2 - t1) + (t4 - t3)
        return float(loss)
    step.split = split
    return step


def run_cpu(batch, steps, warmup):
    import torch
    cores = os.cpu_count() or 1
    step = cpu_step_factory(batch, threads=cores)
    for _ in range(warmup):
        step()
    step.split.update(backbone_s=0.0, head_s=0.0)
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return dict(value=steps * batch / dt, unit="imgs/s", cores=torch.get_num_threads(), kind="port",
                sample="%d steps of batch %d (of the %d/GPU workload), torch-CPU fp32 on all host threads: oracle/model_cpu.py "
                       "(backbone) + oracle/partial_fc_torch.py (PartialFC step, the reference's own ATen call sequence)" % (steps, batch, BATCH),
                ms_per_step=dt / steps * 1e3, head_share=round(step.split["head_s"] / max(dt, 1e-9), 4),
                backbone_share=round(step.split["backbone_s"] / max(dt, 1e-9), 4))


# ------------------------------------------------------------------------------------ GPU arm
def collect_profile(lib):
    import ctypes
    buf = ctypes.create_string_buffer(1 << 16)
    lib.msml_profile_collect(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms, work, nbytes = line.split()
        out[name] = dict(launches=int(n), total_ms=float(ms), work=float(work), min_bytes=float(nbytes))
    return out


_TRAFFIC = None
TRAFFIC_FILES = ("r02_traffic.json", "r01c_traffic.json")      # newest capture first


def ncu_traffic(name, suffix=""):
    """-> (dram__bytes_read.sum + dram__bytes_write.sum per launch, source) from the newest COMMITTED `ncu --set full`
    capture of this kernel (profiles/*_traffic.json), or (None, None) if it was never captured.  Not measured in this run:
    ncu cannot run inside a timed benchmark, so the number is a property of the kernel at the captured commit and the
    JSON line says which file it came from."""
    global _TRAFFIC
    if _TRAFFIC is None:
        _TRAFFIC = []
        for f in TRAFFIC_FILES:
            path = os.path.join(ROOT, "profiles", f)
            if os.path.exists(path):
                _TRAFFIC.append((f, json.load(open(path))))
    for sfx in (suffix if isinstance(suffix, (list, tuple)) else (suffix,)):
        for f, table in _TRAFFIC:
            rec = table.get(name + sfx)
            if rec:
                return rec["dram_bytes_per_launch"], "profiles/%s (%s)" % (f, rec.get("captured", table.get("_captured", "ncu --set full, earlier commit")))
    return None, None


def roofline_entry(name, rec, pk, sustained=True, traffic_suffix=""):
    traffic, traffic_source = ncu_traffic(name, traffic_suffix)
    avg_s = rec["total_ms"] / rec["launches"] * 1e-3
    per_launch = rec["work"] / rec["launches"]
    if name.endswith("_gemm"):
        peak = pk["tf_sustained"] if sustained else pk["tf_burst"]
        ach = per_launch / avg_s / 1e12
        bytes_pl = rec.get("min_bytes", 0.0) / rec["launches"]
        t_tensor, t_hbm = per_launch / (peak * 1e12), bytes_pl / (pk["hbm"] * 1e9)
        if t_hbm > t_tensor:   # short-M contraction: the class-centre stream binds, not the tensor pipe
            gbs = bytes_pl / avg_s / 1e9
            return dict(kernel=name, bound="hbm", achieved=round(gbs, 1), peak=pk["hbm"], unit="GB/s", frac=round(gbs / pk["hbm"], 4),
                        traffic=traffic, traffic_source=traffic_source, launches=rec["launches"], avg_us=round(avg_s * 1e6, 2), work_per_launch=bytes_pl,
                        tflops=round(ach, 2), note="HBM-bound at this M: min bytes = operands + outputs streamed once",
                        peak_source=pk["source"])
        return dict(kernel=name, bound="tensor", achieved=round(ach, 2), peak=peak, unit="TFLOP/s", frac=round(ach / peak, 4),
                    traffic=traffic, traffic_source=traffic_source, launches=rec["launches"], avg_us=round(avg_s * 1e6, 2), work_per_launch=per_launch,
                    peak_source=pk["source"] + (", sustained" if sustained else ", burst"))
    ach = per_launch / avg_s / 1e9
    return dict(kernel=name, bound="hbm", achieved=round(ach, 1), peak=pk["hbm"], unit="GB/s", frac=round(ach / pk["hbm"], 4),
                traffic=traffic, traffic_source=traffic_source, launches=rec["launches"], avg_us=round(avg_s * 1e6, 2), work_per_launch=per_launch,
                peak_source=pk["source"])


def fusion_microbench(iters=20, batch=512):
    """BASELINE config 2: 4-scale mask fusion fwd+bwd, bf16 NHWC, one batched launch each way."""
    import torch
    from msml_b200 import ops
    shapes = [(batch, 64, 56, 56), (batch, 128, 28, 28), (batch, 256, 14, 14), (batch, 512, 7, 7)]
    g = torch.Generator(device="cuda").manual_seed(1)
    mk = lambda s: torch.randn(s, device="cuda", generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    yfs, zs, ds = [mk(s) for s in shapes], [mk(s) for s in shapes], [mk(s) for s in shapes]
    for _ in range(3):
        ops.fm_gate_multi_fwd(yfs, zs); ops.fm_gate_multi_bwd(ds, yfs, zs)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * iters)]
    for i in range(iters):      # 385 MB per tensor pass >> 126 MB L2: no flush needed
        ev[3 * i].record()
        outs = ops.fm_gate_multi_fwd(yfs, zs)
        ev[3 * i + 1].record()
        dyfs, dzs = ops.fm_gate_multi_bwd(ds, yfs, zs)
        ev[3 * i + 2].record()
        del outs, dyfs, dzs
    torch.cuda.synchronize()
    fwd = sorted(ev[3 * i].elapsed_time(ev[3 * i + 1]) for i in range(iters))[iters // 2]
    bwd = sorted(ev[3 * i + 1].elapsed_time(ev[3 * i + 2]) for i in range(iters))[iters // 2]
    elems = FEAT_ELEMS_PER_IMG * batch
    return dict(batch=batch, fwd_ms=round(fwd, 4), bwd_ms=round(bwd, 4), fwd_gbs=round(3 * elems * 2 / fwd / 1e6, 1),
                bwd_gbs=round(5 * elems * 2 / bwd / 1e6, 1), fwd_bwd_gbs=round(8 * elems * 2 / (fwd + bwd) / 1e6, 1),
                algorithmic_bytes=8 * elems * 2,
                ncu_dram_bytes={"fwd": ncu_traffic("fm_gate_fwd", "@config2")[0], "bwd": ncu_traffic("fm_gate_bwd", "@config2")[0],
                                "source": ncu_traffic("fm_gate_fwd", "@config2")[1]} if batch == 512 else None)


def head_microbench(iters=5, b_tot=1024, n_s=125000, fused_projection=False):
    """BASELINE config 4 per-rank shapes on ONE GPU: the 8-GPU run gathers B_tot = 8 x 128 = 1024 embeddings against a
    125,000-class shard per rank (1M classes / 8); the head kernels see exactly these GEMM shapes, only the collectives
    are absent.  Every launch is bracketed by CUDA events (msml_profile_*); tensor-bound kernels are quoted against the
    measured cuBLAS bf16 burst peak.  fused_projection=False: forward_backward alone, sub_weight.grad is the exact gradient.
    fused_projection=True: the step as bench.py trains it — forward_backward + headers.PartialFCSGD(fuse_projection=True),
    where the backward of normalize(sub_weight) moves from the dcos / dW epilogues into the optimizer kernel (whose time is
    reported beside the GEMMs, not hidden)."""
    import torch
    from msml_b200 import _lib
    from msml_b200.headers import ArcFace, PartialFC, PartialFCSGD
    lib = _lib.load()
    pk = peaks()
    torch.manual_seed(1)
    pfc = PartialFC(0, torch.cuda.current_device(), 1, b_tot, False, ArcFace(S, M), n_s, sample_rate=1.0, embedding_size=512)
    opt = PartialFCSGD(pfc, lr=0.1, momentum=0.9, weight_decay=5e-4, fuse_projection=True) if fused_projection else None
    g = torch.Generator(device="cuda").manual_seed(1)
    feat = torch.nn.functional.normalize(torch.randn(b_tot, 512, device="cuda", generator=g))
    label = torch.randint(0, n_s, (b_tot,), device="cuda", generator=g)

    def one():
        pfc.forward_backward(label, feat, opt)
        if opt is not None:
            opt.step()
        pfc.sub_weight.grad = None
    for _ in range(3):
        one()
    torch.cuda.synchronize()
    lib.msml_profile_enable(1)
    for _ in range(iters):
        one()
    torch.cuda.synchronize()
    lib.msml_profile_enable(0)
    prof = collect_profile(lib)
    sfx = ("@config4raw", "@config4") if fused_projection else "@config4"
    rl = sorted((roofline_entry(k, v, pk, sustained=False, traffic_suffix=sfx) for k, v in prof.items()),
                key=lambda r: -r["avg_us"] * r["launches"])
    gemm_ms = sum(v["total_ms"] for k, v in prof.items() if k.endswith("_gemm")) / iters
    del pfc
    torch.cuda.empty_cache()
    return dict(b_tot=b_tot, n_s=n_s, algorithmic_flops=6.0 * b_tot * n_s * 512, fused_projection=fused_projection,
                gemm_us=round(gemm_ms * 1e3, 2),
                tflops_over_gemm_time=round(6.0 * b_tot * n_s * 512 / (gemm_ms * 1e-3) / 1e12, 1),
                frac_of_bf16_burst_peak=round(6.0 * b_tot * n_s * 512 / (gemm_ms * 1e-3) / 1e12 / pk["tf_burst"], 4), rooflines=rl)


def aux_microbench(iters=10):
    """The kernels around the two hot ops, each timed alone with CUDA events (median of `iters`, inputs > L2 or rotated) next to
    the ATen sequence it replaces on the same GPU:
      fm_cat    FM input assembly at BASELINE config-2 shapes (B = 512, 4 scales)      vs torch.cat + strided add
      consensus segmentation criterion on final_seg (B = 512, 2 x 112 x 112)           (no ATen twin here: the reference loops in Python)
      pfc_sgd   fused PartialFC optimizer, config 3 at W = 1 (93,431 x 512, full) and a config-4 sample (12,500 of 125,000 rows)
                vs gather -> torch.optim.SGD(fused) -> scatter."""
    import torch
    from msml_b200 import _lib, ops
    lib = _lib.load()
    pk = peaks()
    g = torch.Generator(device="cuda").manual_seed(1)

    def med(fn, n=iters):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return sorted(ts)[len(ts) // 2]

    def entry(name, ms, nbytes, base_ms=None):
        gbs = nbytes / ms / 1e6
        return dict(kernel=name, bound="hbm", ms=round(ms, 4), achieved=round(gbs, 1), peak=pk["hbm"], unit="GB/s", frac=round(gbs / pk["hbm"], 4),
                    algorithmic_bytes=nbytes, aten_ms=None if base_ms is None else round(base_ms, 4))

    out = []
    B = 512
    shapes = [(B, 64, 56, 56), (B, 128, 28, 28), (B, 256, 14, 14), (B, 512, 7, 7)]
    fwd_b = bwd_b = 0
    yfs, yos, dcs, dts = [], [], [], []
    for (b, c, h, w) in shapes:
        mk = lambda ch: torch.randn(b, ch, h, w, device="cuda", generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        ct = -(-(c + 18) // 8) * 8
        yfs.append(mk(c)); yos.append(mk(18)); dcs.append(mk(ct)); dts.append(mk(c))
        fwd_b += b * h * w * (c + 18 + ct) * 2
        bwd_b += b * h * w * 3 * c * 2

    def cat_fwd():
        for yf, yo in zip(yfs, yos):
            ops.fm_cat(yf, yo)

    def cat_fwd_aten():
        for yf, yo in zip(yfs, yos):
            ops.cat_channels_padded((yf, yo))

    def cat_bwd():
        st = torch.cuda.current_stream().cuda_stream
        for yf, dc, dt in zip(yfs, dcs, dts):
            b, c, h, w = yf.shape
            dyf = torch.empty_like(yf)
            _lib.check(lib.msml_fm_cat_bwd(dc.data_ptr(), dt.data_ptr(), dyf.data_ptr(), None, b * h * w, c, 18, dc.shape[1], _lib.BF16, st))

    def cat_bwd_aten():
        for yf, dc, dt in zip(yfs, dcs, dts):
            dt + dc[:, :yf.shape[1]]
    with torch.no_grad():
        out.append(entry("fm_cat_fwd@config2", med(cat_fwd), fwd_b, med(cat_fwd_aten)))
        out.append(entry("fm_cat_bwd@config2", med(cat_bwd), bwd_b, med(cat_bwd_aten)))
    del yfs, yos, dcs, dts

    N, H, W = 512, 112, 112
    z = torch.randn(N, 2, H, W, device="cuda", generator=g).requires_grad_(True)
    msk = torch.zeros(N, H, W, dtype=torch.int64, device="cuda")
    msk[:, 30:70, 20:80] = 1
    one = torch.ones((), device="cuda")
    holder = {}

    def seg_fwd():
        holder["loss"] = ops.consensus_loss(z, msk, msk)

    def seg_bwd():
        z.grad = None
        holder["loss"].backward(one, retain_graph=True)
    out.append(entry("consensus_fwd@B512", med(seg_fwd), N * H * W * (2 * 4 + 16)))
    out.append(entry("consensus_bwd@B512", med(seg_bwd), N * H * W * (2 * 2 * 4 + 8)))
    del z, msk, holder

    st = torch.cuda.current_stream().cuda_stream
    for tag, num_local, n_s in (("config3_w1_full", 93431, 93431), ("config4_sampled", 125000, 12500)):
        w = torch.randn(num_local, 512, device="cuda", generator=g) * 0.01
        mom = torch.zeros_like(w)
        dw = torch.randn(n_s, 512, device="cuda", generator=g) * 0.01
        index = None if n_s == num_local else torch.sort(torch.randperm(num_local, device="cuda", generator=g)[:n_s]).values

        def fused():
            _lib.check(lib.msml_pfc_sgd_update(w.data_ptr(), mom.data_ptr(), dw.data_ptr(), index.data_ptr() if index is not None else None,
                                               n_s, num_local, 512, None, 0.1, 0.9, 5e-4, 0.0, 0, None, None, st))
        sub = torch.nn.Parameter(w if index is None else w[index].clone())
        subm = mom if index is None else mom[index].clone()
        opt = torch.optim.SGD([sub], lr=0.1, momentum=0.9, weight_decay=5e-4, fused=True)
        opt.state[sub]["momentum_buffer"] = subm
        sub.grad = dw

        def stock():
            if index is not None:
                sub.data.copy_(w[index]); subm.copy_(mom[index])
            opt.step()
            if index is not None:
                w[index] = sub.data; mom[index] = subm
        with torch.no_grad():
            out.append(entry("pfc_sgd_update@" + tag, med(fused), n_s * 512 * 20, med(stock)))
        del w, mom, dw, sub, subm, opt
    torch.cuda.empty_cache()
    return out


def head_parity_check(pfc, rank, world, dev, batch, classes, tol=1e-3):
    """One head step on seeded L2-normalised embeddings on every rank (real NCCL ranks when world > 1); rank 0 gathers
    the inputs and the class shards and evaluates the reference's loss in fp64 (oracle/partial_fc.py: forward_loss follows
    ref headers/partial_fc.py:132-163 over all simulated ranks).  Raises if the loss the ranks report differs by more
    than `tol` relative.  Only for sample_rate 1 (the sampled path is checked index-for-index in run_head)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    feat = torch.nn.functional.normalize(torch.randn(batch, 512, device=dev, generator=g))
    label = torch.randint(0, classes, (batch,), device=dev, generator=g)
    x_grad, loss = pfc.forward_backward(label, feat, None)
    pfc.sub_weight.grad = None
    feat16 = feat.to(torch.bfloat16)                   # what the head contracts (it gathers the embeddings in bf16)
    nl_max = classes // world + 1
    wpad = torch.zeros(nl_max, 512, device=dev)
    wpad[:pfc.num_local] = pfc.weight
    if world > 1:
        feats = [torch.empty_like(feat16) for _ in range(world)]
        labels = [torch.empty_like(label) for _ in range(world)]
        ws = [torch.empty_like(wpad) for _ in range(world)]
        losses = [torch.empty_like(loss) for _ in range(world)]
        dist.all_gather(feats, feat16); dist.all_gather(labels, label); dist.all_gather(ws, wpad); dist.all_gather(losses, loss)
    else:
        feats, labels, ws, losses = [feat16], [label], [wpad], [loss]
    out = None
    if rank == 0:
        from oracle import partial_fc as opfc
        t0 = time.perf_counter()
        weights = [ws[r][:opfc.shard_geometry(classes, world, r)[0]].double().cpu().numpy() for r in range(world)]
        want, _, _ = opfc.forward_loss([f.double().cpu().numpy() for f in feats], [l.cpu().numpy() for l in labels], weights,
                                       classes, "arc", S, M)
        got = [float(l) for l in losses]
        rel = max(abs(v - want) for v in got) / abs(want)
        out = {"loss_per_rank": [round(v, 6) for v in got], "oracle_loss_fp64": round(float(want), 6), "max_rel_err": float("%.3g" % rel),
               "tolerance": tol, "ranks": world, "b_tot": batch * world, "classes": classes, "x_grad_finite": bool(torch.isfinite(x_grad).all()),
               "oracle_s": round(time.perf_counter() - t0, 1)}
        if not (rel <= tol and np.isfinite(want)):
            raise SystemExit("bench.py: head parity check FAILED on %d rank(s): %r" % (world, out))
    if world > 1:
        dist.barrier()
    return out


def run_train(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from msml_b200 import _lib, ops
    from msml_b200.backbones import MSML
    from msml_b200.engine import FlatSGD, TrainStep, broadcast_parameters
    from msml_b200.headers import ArcFace, PartialFC, PartialFCSGD

    lib = _lib.load()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.benchmark = True
    pk = peaks()

    torch.manual_seed(1)                                     # same init on every rank, then broadcast (ref train.py:133-134)
    net = MSML("iresnet50", "unet", (1, 1, 1, 1), NUM_CLASSES, fp16=True, header_type=None, fm_params=FM_PARAMS).to(dev).train()
    broadcast_parameters(net)
    pfc = PartialFC(rank, local_rank, world, BATCH, False, ArcFace(S, M), NUM_CLASSES, sample_rate=1.0, embedding_size=512)
    lr = 0.1 * BATCH * world / 512
    # the backbone optimizer (ref train.py:186-191: SGD over backbone.parameters(), momentum 0.9, wd 5e-4) as ONE kernel over
    # flat parameter / momentum / gradient buffers that also writes next step's bf16 shadow weights (engine.FlatSGD)
    bb_params = [p for p in net.parameters() if p.requires_grad]
    opt = (torch.optim.SGD(bb_params, lr=lr, momentum=0.9, weight_decay=5e-4, fused=True) if args.stock_backbone_sgd
           else FlatSGD(bb_params, lr=lr, momentum=0.9, weight_decay=5e-4))
    # the head's optimizer (ref train.py:188-191: SGD over module_partial_fc.parameters(), momentum 0.9, wd 5e-4) as ONE
    # kernel on the shard rows that also emits the next step's normalised bf16 centres
    opt_pfc = (torch.optim.SGD([{"params": pfc.parameters()}], lr=lr, momentum=0.9, weight_decay=5e-4, fused=True) if args.stock_head_sgd
               else PartialFCSGD(pfc, lr=lr, momentum=0.9, weight_decay=5e-4, emit_normalized=True, fuse_projection=not args.exact_head_grad))
    step = TrainStep(net, pfc, opt, opt_pfc, (BATCH, 3, 112, 112), world_size=world, max_norm=5.0, use_graph=not args.eager)
    head_check = head_parity_check(pfc, rank, world, dev, BATCH, NUM_CLASSES)     # before anything trains the class centres

    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    n_buf = 4
    imgs = [torch.randn(BATCH, 3, 112, 112, device=dev, generator=gen).contiguous(memory_format=torch.channels_last) for _ in range(n_buf)]
    labels = [torch.randint(0, NUM_CLASSES, (BATCH,), device=dev, generator=gen) for _ in range(n_buf)]
    imgs_h = [t.cpu().pin_memory() for t in imgs]
    labels_h = [t.cpu().pin_memory() for t in labels]

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, host_fed):
        fence()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        last = 0.0
        if host_fed:
            # every step: H2D of that step's inputs from pinned memory (issued through the engine's prefetcher, so it
            # runs under the previous step, as a data loader's prefetcher would) and a D2H read of that step's loss
            # (an async copy into pinned memory, read one step later — how a training loop logs its loss without draining the
            # GPU: step i+1 is already enqueued when the host waits for the loss of step i)
            loss_h = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
            loss_ev = [torch.cuda.Event() for _ in range(2)]
            step.prefetch(imgs_h[0], labels_h[0])
            for i in range(n):
                loss_t = step()
                loss_h[i & 1].copy_(loss_t, non_blocking=True)
                loss_ev[i & 1].record()
                if i + 1 < n:
                    step.prefetch(imgs_h[(i + 1) % n_buf], labels_h[(i + 1) % n_buf])
                if i > 0:
                    loss_ev[(i - 1) & 1].synchronize()
                    last = float(loss_h[(i - 1) & 1])
            loss_ev[(n - 1) & 1].synchronize()
            last = float(loss_h[(n - 1) & 1])
        else:
            for i in range(n):
                last = step(imgs[i % n_buf], labels[i % n_buf])
        b.record()
        fence()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, float(last)

    for i in range(args.warmup):                             # includes the 3 eager steps + graph capture
        step(imgs[i % n_buf], labels[i % n_buf])
    fence()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if os.environ.get("MSML_PROFILER_RANGE"):          # ncu --profile-from-start off: only the timed region is profiled
        torch.cuda.cudart().cudaProfilerStart()
    ms, loss = timed(args.steps, host_fed=False)
    if os.environ.get("MSML_PROFILER_RANGE"):
        torch.cuda.cudart().cudaProfilerStop()
    ms_e2e, loss_e2e = timed(args.steps, host_fed=True)
    clocks = sampler.stop() if sampler else None

    # per-kernel roofline pass: the SAME step run eagerly (a CUDA graph cannot be bracketed kernel by
    # kernel), every launch of this library between two CUDA events on its own stream
    prof_steps = 3
    # (single stream: with the side stream on, concurrent kernels would inflate each other's event-bracketed durations)
    eager = step if args.eager else TrainStep(net, pfc, opt, opt_pfc, (BATCH, 3, 112, 112), world_size=world, max_norm=5.0,
                                              use_graph=False, wgrad_side_stream=False)
    if not args.eager:
        eager.share_state_from(step)
    eager(imgs[0], labels[0])
    fence()
    ops.launch_count_reset()
    lib.msml_profile_enable(1)
    for i in range(prof_steps):
        eager(imgs[i % n_buf], labels[i % n_buf])
    torch.cuda.synchronize()
    lib.msml_profile_enable(0)
    launches_per_step = ops.launch_count() // prof_steps
    prof = collect_profile(lib)
    fence()

    if rank != 0:
        return None
    value = args.steps * BATCH * world / (ms * 1e-3)
    rl = sorted((roofline_entry(k, v, pk) for k, v in prof.items()), key=lambda r: -r["avg_us"] * r["launches"])
    mine_ms = sum(v["total_ms"] for v in prof.values()) / prof_steps
    res = {
        "metric": "train imgs/s ires50-MSML+PartialFC", "value": round(value, 1), "unit": "imgs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "ires50_msml + PartialFC(ArcFace s=64 m=0.5, 93431 classes, sample_rate 1) bf16 training step, "
                               "112x112, batch 128/GPU (BASELINE config 3)",
                   "global_batch": BATCH * world,
                   "parallelism": "dp%d backbone (flat-gradient NCCL all-reduce) + class-sharded head" % world,
                   "execution": "eager" if args.eager else "whole step captured in one CUDA graph, two streams (wgrad + OSB on a side stream) (msml_b200.engine.TrainStep)",
                   "optimizers": {"backbone": "torch.optim.SGD(fused=True)" if args.stock_backbone_sgd else "engine.FlatSGD (one kernel, emits bf16 shadow weights)",
                                  "head": "torch.optim.SGD(fused=True) + update()" if args.stock_head_sgd else "headers.PartialFCSGD"},
                   "l2": "per-step working set (activations, GBs) >> 126 MB L2; 4 rotating input batches"},
        "e2e": {"value": round(args.steps * BATCH * world / (ms_e2e * 1e-3), 1), "unit": "imgs/s",
                "h2d_bytes_per_step": imgs_h[0].numel() * 4 + labels_h[0].numel() * 8, "d2h_bytes_per_step": 4,
                "ms_per_step": round(ms_e2e / args.steps, 3),
                "how": "TrainStep.prefetch(pinned img, label) under the previous step + TrainStep(); every step's loss copied to pinned "
                       "memory and read by the host one step later"},
        "gpu_launches": int(launches_per_step * args.steps),
        "gpu_launches_per_step": int(launches_per_step),
        "own_kernel_ms_per_step": round(mine_ms, 4),
        "roofline": rl[0] if rl else None,
        "rooflines": rl,
        "roofline_pass": "%d eager single-stream replays of the same step after the timed region, CUDA events around every launch" % prof_steps,
        "clocks": clocks,
        "loss": round(loss, 4),
        "head_check": head_check,
    }
    return res


def sampled_index_check(pfc, label_all, rank, dev):
    """BASELINE config 4: the class indices this rank samples must equal, bit for bit, what ref headers/partial_fc.py:83-94
    selects from the same torch.rand draw (oracle/partial_fc.py: sample).  -> dict for the JSON line."""
    import numpy as np
    import torch
    from oracle import partial_fc as opfc
    state = torch.cuda.get_rng_state(dev)
    tl = label_all.clone()
    pfc.sample(tl)
    torch.cuda.synchronize()
    torch.cuda.set_rng_state(state, dev)
    perm = torch.rand(size=[pfc.num_local], device=dev).cpu().numpy()          # the draw sample() consumed
    want_tl, want_index = opfc.sample(label_all.cpu().numpy(), perm, pfc.class_start, pfc.num_local, pfc.num_sample, pfc.sample_rate)
    ok = bool(np.array_equal(pfc.index.cpu().numpy(), want_index) and np.array_equal(tl.cpu().numpy(), want_tl))
    return {"rank": rank, "n_index": int(pfc.index.numel()), "bit_exact": ok}


def run_head(args, rank, local_rank, world):
    """BASELINE config 4: PartialFC head-only step (forward_backward -> optimizer step -> update) on a per-rank shard of
    `classes`, B = --batch per GPU, sample_rate 0.1 / 1.0, class-sharded over `world` ranks."""
    import torch
    import torch.distributed as dist
    from msml_b200 import _lib, ops
    from msml_b200.headers import ArcFace, PartialFC, PartialFCSGD
    lib = _lib.load()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    pk = peaks()
    torch.manual_seed(1)
    BATCH = args.batch                                       # per-rank batch (shadows the module constant on purpose)
    sampled = int(args.sample_rate) != 1
    pfc = PartialFC(rank, local_rank, world, BATCH, False, ArcFace(S, M), args.classes, sample_rate=args.sample_rate, embedding_size=512)
    hp = dict(lr=0.1, momentum=0.9, weight_decay=5e-4)
    opt = (PartialFCSGD(pfc, fuse_projection=not args.exact_head_grad, emit_normalized=True, **hp) if args.fused_sgd
           else torch.optim.SGD([{"params": pfc.parameters()}], **hp))
    check = None
    if not sampled and not args.no_head_check:
        check = head_parity_check(pfc, rank, world, dev, BATCH, args.classes)
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    feat = torch.nn.functional.normalize(torch.randn(BATCH, 512, device=dev, generator=g))
    label = torch.randint(0, args.classes, (BATCH,), device=dev, generator=g)
    if sampled:                                              # bit-exact sampled indices on every rank, before timing
        label_all = torch.empty(BATCH * world, dtype=torch.int64, device=dev)
        pfc.comm.all_gather(label_all, label)
        rec = sampled_index_check(pfc, label_all, rank, dev)
        recs = [None] * world
        if world > 1:
            dist.all_gather_object(recs, rec)
        else:
            recs = [rec]
        check = {"sampled_index_bit_exact_per_rank": [r["bit_exact"] for r in recs], "n_index_per_rank": [r["n_index"] for r in recs]}
        if rank == 0 and not all(check["sampled_index_bit_exact_per_rank"]):
            raise SystemExit("bench.py: sampled class indices differ from the oracle: %r" % check)

    def one():
        pfc.forward_backward(label, feat, opt); opt.step(); pfc.update()
    # warm-up: the caching allocator needs a few steps to settle (the first two sampled steps cudaMalloc ~0.4 GB of gather
    # buffers each: tens of ms, once), so at least 5 untimed steps whatever --warmup says
    for _ in range(max(args.warmup, 5)):
        one()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    lib.msml_profile_enable(1)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        one()
        ev[i + 1].record()
    torch.cuda.synchronize()
    lib.msml_profile_enable(0)
    prof = collect_profile(lib)
    per_step = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps))
    ms = ev[0].elapsed_time(ev[-1])
    if world > 1:
        t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    if rank != 0:
        return None
    n_s = pfc.sub_weight.shape[0]
    flops = 6.0 * BATCH * world * n_s * 512
    rl = sorted((roofline_entry(k, v, pk, sustained=False) for k, v in prof.items()), key=lambda r: -r["avg_us"] * r["launches"])
    gemm_ms = sum(v["total_ms"] for k, v in prof.items() if k.endswith("_gemm")) / args.steps
    own_ms = sum(v["total_ms"] for v in prof.values()) / args.steps
    return {"metric": "PartialFC head step", "value": round(args.steps * BATCH * world / (ms * 1e-3), 1), "unit": "imgs/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 5), "ms_per_step": round(ms / args.steps, 4),
            "ms_per_step_median": round(per_step[len(per_step) // 2], 4), "ms_per_step_max": round(per_step[-1], 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "PartialFC head only (forward_backward + SGD + update), %d classes, sample_rate %g, B=%d/GPU, %d rank(s) (BASELINE config 4)"
                                   % (args.classes, args.sample_rate, BATCH, world),
                       "n_s_per_rank": n_s, "optimizer": ("headers.PartialFCSGD (fused%s)" % ("" if args.exact_head_grad else ", normalise-backward projection in the optimizer kernel")) if args.fused_sgd else "torch.optim.SGD + update()"},
            "head_algorithmic_tflops_over_gemm_time": round(flops / (gemm_ms * 1e-3) / 1e12, 2),
            "own_kernel_ms_per_step": round(own_ms, 4), "gemm_ms_per_step": round(gemm_ms, 4),
            "gpu_launches": int(ops.launch_count()), "head_check": check,
            "rooflines": rl, "roofline": rl[0] if rl else None}


def run_infer(args, rank, local_rank, world):
    """BASELINE config 5: test.py-style occluded-face verification embedding extraction (LFW shape), ires50_msml, batch
    1024, one B200.  A step = one batch of 1024 uint8 images: H2D, normalise, f(x) + f(flip x), L2-normalise, D2H."""
    import numpy as np
    import torch
    from msml_b200.backbones import MSML
    from msml_b200.datasets.augment.rand_occ import random_block_batch
    from msml_b200.eval import extract_embeddings
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(1)
    net = MSML("iresnet50", "unet", (1, 1, 1, 1), NUM_CLASSES, fp16=True, header_type=None, fm_params=FM_PARAMS).to(dev)
    net = net.to(memory_format=torch.channels_last).eval()
    bs = 1024
    n = bs * max(1, min(args.steps, 12))                    # LFW: 12,000 images
    # LFW-shape synthetic set, occluded on the host exactly as ref eval/qeval_mxnet.py:528-547 does: RandomBlock(40, 41, 'black')
    # (msml_b200.datasets.augment mirrors ref datasets/augment/rand_occ.py:25-72 pixel for pixel for a given numpy seed)
    np.random.seed(1)
    imgs = np.random.randint(0, 256, (n, 3, 112, 112), dtype=np.uint8)
    imgs = torch.from_numpy(random_block_batch(imgs, 40, 41, 'black')).pin_memory()
    for _ in range(max(args.warmup, 3)):
        extract_embeddings([imgs[:bs]], net, bs)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    emb, _ = extract_embeddings([imgs], net, bs)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    assert np.isfinite(emb).all() and abs(np.linalg.norm(emb[0]) - 1.0) < 1e-6
    return {"metric": "occluded verification embedding extraction imgs/s (BASELINE config 5)", "value": round(n / (ms * 1e-3), 1),
            "unit": "imgs/s", "n_gpus": 1, "steps": n // bs, "warmup": max(args.warmup, 3), "ms_per_step": round(ms / (n // bs), 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "ires50_msml eval, %d uint8 112x112 images, RandomBlock(40, 41, 'black') (40 %% of the area, ref rand_occ.py:25-72), batch 1024, f(x)+f(flip x), "
                                   "host-fed (pinned uint8) and embeddings read back: end to end" % n}}


def main():
    # the contract is ONE JSON line on stdout: native libraries (NCCL's version banner, cuDNN warnings) print to fd 1,
    # so fd 1 is pointed at stderr for the whole run and the JSON goes to the saved descriptor
    real_stdout = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    sys.stdout = os.fdopen(os.dup(2), "w")

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "fusion", "head", "infer", "aux"])
    ap.add_argument("--classes", type=int, default=1_000_000)
    ap.add_argument("--sample-rate", type=float, default=1.0)
    ap.add_argument("--batch", type=int, default=BATCH, help="head workload only: per-rank batch (1024 on one GPU reproduces the "
                    "per-rank GEMM shapes of the 8-GPU config-4 run: B_tot=1024 rows against a 125,000-class shard)")
    ap.add_argument("--cpu-batch", type=int, default=16, help="images per step of the CPU arm (a bounded sample of the 128/GPU workload)")
    ap.add_argument("--stock-backbone-sgd", action="store_true", help="train workload: torch.optim.SGD(fused=True) for the backbone instead "
                    "of engine.FlatSGD (one flat kernel + emitted bf16 shadow weights)")
    ap.add_argument("--stock-head-sgd", action="store_true", help="train workload: torch.optim.SGD(fused=True) + update() for the class centres "
                    "instead of headers.PartialFCSGD")
    ap.add_argument("--exact-head-grad", action="store_true", help="with headers.PartialFCSGD: keep the normalise-backward projection in the "
                    "GEMM epilogues (sub_weight.grad is the exact gradient) instead of fusing it into the optimizer kernel")
    ap.add_argument("--fused-sgd", action="store_true", help="head workload: headers.PartialFCSGD instead of torch.optim.SGD + update()")
    ap.add_argument("--no-head-check", action="store_true", help="skip the fp64-oracle loss check before the timed region")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="run the step eagerly instead of replaying the captured CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = min(args.steps, 8)
        cb = run_cpu(args.cpu_batch, steps, min(args.warmup, 1))
        emit(({
            "impl": "reference", "metric": "train imgs/s ires50-MSML+PartialFC", "value": round(cb["value"], 3), "unit": "imgs/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": round(cb["ms_per_step"], 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ires50_msml + PartialFC(ArcFace s=64 m=0.5, 93431 classes, sample_rate 1) training step on host CPU "
                                   "cores, 112x112 (BASELINE config 3 shapes; bounded sample: batch %d per step)" % args.cpu_batch},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "head_share", "backbone_share")},
            "e2e": {"value": round(cb["value"], 3), "unit": "imgs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if args.workload == "fusion":
        torch.cuda.set_device(local_rank)
        fm = fusion_microbench(iters=max(args.steps, 5))
        pk = peaks()
        if rank == 0:
            emit(({"metric": "mask-fusion fwd+bwd GB/s (BASELINE config 2)", "value": fm["fwd_bwd_gbs"], "unit": "GB/s",
                              "n_gpus": 1, "steps": args.steps, "warmup": 3, "ms_per_step": round(fm["fwd_ms"] + fm["bwd_ms"], 4),
                              "higher_is_better": True, "dtype": "bf16", "data": "synthetic", "vs_baseline": None, "scaling": "weak",
                              "config": {"workload": "4-scale mask fusion, batch 512, bf16 NHWC, fwd+bwd, one launch each way"},
                              "roofline": {"bound": "hbm", "achieved": fm["fwd_bwd_gbs"], "peak": pk["hbm"], "unit": "GB/s",
                                           "frac": round(fm["fwd_bwd_gbs"] / pk["hbm"], 4), "traffic": None}, "detail": fm}))
        return 0

    if args.workload == "aux":
        torch.cuda.set_device(local_rank)
        if rank == 0:
            rl = aux_microbench(iters=max(args.steps, 5))
            emit({"metric": "auxiliary kernels GB/s (FM concat, consensus loss, fused PartialFC SGD)", "value": rl[0]["achieved"], "unit": "GB/s",
                  "n_gpus": 1, "steps": args.steps, "warmup": 3, "ms_per_step": rl[0]["ms"], "higher_is_better": True, "dtype": "bf16",
                  "data": "synthetic", "vs_baseline": None, "scaling": "weak",
                  "config": {"workload": "each kernel alone, median of CUDA-event timings; inputs larger than L2"}, "roofline": rl[0], "rooflines": rl})
        return 0

    if args.workload == "infer":
        if rank == 0:
            emit(run_infer(args, rank, local_rank, world))
        return 0
    res = run_head(args, rank, local_rank, world) if args.workload == "head" else run_train(args, rank, local_rank, world)
    if rank == 0 and args.workload == "train":
        res["fusion_microbench"] = fusion_microbench()
        res["head_microbench"] = head_microbench()
        res["head_microbench_fused_projection"] = head_microbench(fused_projection=True)
        if world == 1 and not args.no_cpu_baseline:
            cb = run_cpu(args.cpu_batch, 4, 1)
            res["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "head_share", "backbone_share")}
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(res)
    return 0


if __name__ == "__main__":
    sys.exit(main())
