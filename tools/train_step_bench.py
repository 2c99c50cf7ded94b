"""BASELINE.json config 4 (GPU): one TRAINING step of spiking_fullsubnet-<size> (forward + backward through the
surrogate gradient + AdamW), utterance batch sharded over the ranks, gradients all-reduced by NCCL through
torch DDP (exactly what `accelerator.prepare(model)` sets up in the reference, recipes/.../run.py:39).

    python tools/train_step_bench.py [--size L] [--batch 32] [--seconds 6] [--steps 3] [--cpu-sample]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_step_bench.py ...

Loss: the recipe's own (recipes/.../trainer.py:33-37): freq_MAE + mag_MAE + 0.001 * (100 - SI-SNR) from
spiking_fullsubnet_b200.losses (row f3; each 2048-point STFT computed once).  `--loss standin` keeps the earlier
waveform-L1 + magnitude-L1 stand-in (the rows of profiles/r01_training_step.md before the last one used it).
Prints ONE JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402  (synthetic weights / inputs only)
from spiking_fullsubnet_b200 import SpikingFullSubNet, losses  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="L")
ap.add_argument("--batch", type=int, default=32, help="clips per GPU")
ap.add_argument("--seconds", type=float, default=6.0, help="the recipe's training crop (dataloader.py:13)")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--cpu-sample", action="store_true", help="also time a bounded CPU sample of the same step")
ap.add_argument("--loss", default="recipe", choices=["recipe", "standin"])
args = ap.parse_args()

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cfg = synth.CONFIGS[args.size]
L = int(args.seconds * 16000)
T = 1 + L // cfg["hop_length"]
model = SpikingFullSubNet(**cfg)
model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.make_params(cfg, 5).items()})
model = model.to(dev).train()
net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
wave = torch.from_numpy(synth.make_wave(args.batch, L, 31 + rank)).to(dev)
clean = torch.from_numpy(synth.make_wave(args.batch, L, 41 + rank)).to(dev)
clean_mag = torch.stft(clean, 512, 128, 512, window=torch.hann_window(512, device=dev), return_complex=True,
                       pad_mode="constant").abs()


def step():
    opt.zero_grad(set_to_none=True)
    enh_y, enh_mag, *_ = net(wave)
    if args.loss == "recipe":
        loss = losses.ndns_training_loss(enh_y, clean)["loss"]
    else:
        loss = (enh_y - clean).abs().mean() + (enh_mag - clean_mag).abs().mean()
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    loss = step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    line = {"metric": "training frames/sec", "value": world * args.batch * T / (float(ms) * 1e-3), "unit": "frames/s",
            "n_gpus": world, "ms_per_step": float(ms), "loss": float(loss.detach()),
            "config": {"workload": f"spiking_fullsubnet-{args.size} training step (fwd + BPTT + AdamW), batch "
                                   f"{args.batch} x {args.seconds:g} s per GPU (T={T}), DDP/NCCL gradient all-reduce",
                       "global_batch": world * args.batch, "loss": args.loss},
            "grad_bytes": int(sum(p.numel() for p in model.parameters()) * 4)}
    if args.cpu_sample:
        from oracle import gsn_oracle_torch as OT
        torch.set_num_threads(os.cpu_count() or 1)
        bs, ts = 4, 126  # bounded sample: 4 clips x 1 s
        mag = torch.from_numpy(synth.make_mag(bs, 257, ts, 3))
        params = OT.to_torch(synth.make_params(cfg, 5))
        OT.spiking_fullsubnet_train_step(mag[:, :, :16], params, cfg)
        t0 = time.perf_counter()
        OT.spiking_fullsubnet_train_step(mag, params, cfg)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": bs * ts / dt, "unit": "frames/s", "cores": torch.get_num_threads(),
                                "kind": "port", "sample": f"{bs} clips x {ts} frames, forward (train-mode BN) + "
                                                          f"backward of the path, torch-CPU port"}
    print(json.dumps(line))
if world > 1:
    dist.destroy_process_group()
