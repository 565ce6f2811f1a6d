#!/usr/bin/env python
"""The reference's PartialFC training recipe (ref train.py:188-197 optimizers / schedulers, :283-300 "op2. partial fc"
step, :133-137 broadcast, :367 scheduler step) on the B200-native modules.  Synthetic 112x112 data by default (the
reference's mxnet record reader is out of scope); swap ``batches()`` for a real loader of (uint8 or float images, labels).

    python examples/train_partial_fc.py --steps 200                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_partial_fc.py --steps 200

Differences from the reference loop, all inside ``TrainStep`` and equivalent in result: no GradScaler (bf16 autocast
needs no loss scaling), the whole step is one CUDA-graph replay, gradients are all-reduced from one flat buffer while
the backward pass runs, and the next batch is copied to the device under the current step.
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from msml_b200.backbones import MSML  # noqa: E402
from msml_b200.engine import FlatSGD, TrainStep, broadcast_parameters  # noqa: E402
from msml_b200.headers import ArcFace, PartialFC, PartialFCSGD  # noqa: E402


def batches(batch, num_classes, device_generator, steps):
    """Synthetic stand-in for the reference's DataLoaderX: pinned host batches."""
    for _ in range(steps):
        img = torch.randn(batch, 3, 112, 112, generator=device_generator).contiguous(memory_format=torch.channels_last).pin_memory()
        label = torch.randint(0, num_classes, (batch,), generator=device_generator).pin_memory()
        yield img, label


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--batch", type=int, default=128, help="per GPU (ref config.py: batch_size)")
    ap.add_argument("--classes", type=int, default=93431, help="ref config.py: num_classes (MS1M-RetinaFace)")
    ap.add_argument("--frb", default="iresnet50")
    ap.add_argument("--lr", type=float, default=0.1)
    ap.add_argument("--sample-rate", type=float, default=1.0)
    ap.add_argument("--out", default="./", help="prefix for PartialFC.save_params()")
    ap.add_argument("--stock-backbone-sgd", action="store_true", help="torch.optim.SGD(fused=True) for the backbone, as ref train.py:186-187, "
                    "instead of engine.FlatSGD")
    ap.add_argument("--stock-head-sgd", action="store_true", help="torch.optim.SGD + update() for the class centres, exactly as ref "
                    "train.py:188-191,299-300, instead of the fused headers.PartialFCSGD")
    args = ap.parse_args()

    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(1)                                                   # ref train.py:34-38

    backbone = MSML(args.frb, "unet", (1, 1, 1, 1), args.classes, fp16=True, header_type=None,
                    fm_params=(3, 2, "sigmoid", "mul")).to(dev).train()
    broadcast_parameters(backbone)                                         # ref :133-134
    pfc = PartialFC(rank, local_rank, world, args.batch, False, ArcFace(64.0, 0.5), args.classes,
                    sample_rate=args.sample_rate, embedding_size=512, prefix=args.out)
    lr = args.lr / 512 * args.batch * world                               # ref :176, :190
    bb_params = [p for p in backbone.parameters() if p.requires_grad]
    if args.stock_backbone_sgd:
        opt_backbone = torch.optim.SGD(bb_params, lr=lr, momentum=0.9, weight_decay=5e-4, fused=True)
    else:   # the same optimizer (a torch.optim.SGD subclass) as one kernel over flat parameter / momentum / gradient buffers
        opt_backbone = FlatSGD(bb_params, lr=lr, momentum=0.9, weight_decay=5e-4)
    if args.stock_head_sgd:
        opt_pfc = torch.optim.SGD([{"params": pfc.parameters()}], lr=lr, momentum=0.9, weight_decay=5e-4, fused=True)
    else:   # same hyper-parameters and param_groups (LR schedulers work); one kernel on the shard rows, update() has nothing to scatter
        opt_pfc = PartialFCSGD(pfc, lr=lr, momentum=0.9, weight_decay=5e-4, emit_normalized=True, fuse_projection=True)
    warmup, total = max(1, args.steps // 20), args.steps

    def lr_func(step):                                                     # ref config.py: lr_step_func (warm-up, then decay)
        return (step + 1) / warmup if step < warmup else 0.1 ** sum(step >= total * f for f in (0.5, 0.75, 0.9))
    sched_backbone = torch.optim.lr_scheduler.LambdaLR(opt_backbone, lr_func)
    sched_pfc = torch.optim.lr_scheduler.LambdaLR(opt_pfc, lr_func)

    # sampled heads are captured too (fixed-capacity index / gather buffers) as long as the gathered batch fits in the sample
    step = TrainStep(backbone, pfc, opt_backbone, opt_pfc, (args.batch, 3, 112, 112), world_size=world, max_norm=5.0,
                     use_graph=args.sample_rate == 1.0 or args.batch * world <= pfc.num_sample)
    gen = torch.Generator().manual_seed(1 + rank)
    it = batches(args.batch, args.classes, gen, args.steps)
    step.prefetch(*next(it))
    for i in range(args.steps):
        loss = step()                                                      # forward, PartialFC, backward, clip, both SGD steps, update()
        nxt = next(it, None)
        if nxt is not None:
            step.prefetch(*nxt)                                            # H2D of the next batch under this step
        sched_backbone.step()
        sched_pfc.step()
        if rank == 0 and (i % 20 == 0 or i == args.steps - 1):
            print("step %5d  loss %.4f  lr %.5f" % (i, float(loss), float(opt_backbone.param_groups[0]["lr"])), flush=True)
    pfc.save_params()                                                      # ref partial_fc.py:73-75
    if rank == 0:
        torch.save(backbone.state_dict(), os.path.join(args.out, "backbone.pth"))   # same keys as the reference's checkpoints
    if world > 1:
        dist.barrier()
        del step                                                           # release the captured graph before tearing NCCL down
        torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
