#!/usr/bin/env python
"""bench.py -- frames/s of the GSN hot path (BASELINE.json metric) on N B200s, plus the CPU reference arm.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--size S|M|L|XL]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (STFT magnitude in -> deep-filter coefficients out, SURVEY.md 8a
a9) over one batch of synthetic spectrograms.  Default workload = BASELINE.json configs[1]:
spiking_fullsubnet-S inference, batch 32 x 4 s clips (T = 501 frames of a 512-pt / 257-bin STFT) per GPU.
Weak scaling: every rank processes its own batch of 32 clips (utterance-batch sharding, no collective
on the data path).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", default="S", choices=["S", "M", "L", "XL"])
    ap.add_argument("--batch", type=int, default=32, help="clips per GPU")
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--backend", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chunks", type=int, default=12, help="frame chunks of the wavefront schedule (graph mode)")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel from Python instead of "
                    "replaying the captured CUDA graph")
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="infer: BASELINE configs[1] (default).  train: configs[3], one training step (forward + BPTT "
                         "through the surrogate gradient + AdamW) per GPU, gradients all-reduced by NCCL (DDP); use with "
                         "--size L --batch 32 --seconds 6")
    ap.add_argument("--schedule", default="auto", choices=["auto", "stream", "wavefront"],
                    help="auto/stream: frame-granular streaming pipeline of persistent kernels where it is co-resident "
                         "(else the frame-chunked wavefront); wavefront: round-1 schedule")
    return ap.parse_args()


def workload(args):
    from oracle import synth  # synthetic weights/inputs only (test infrastructure, not on the timed path)
    cfg = synth.CONFIGS[args.size]
    L = int(round(args.seconds * 16000))
    T = 1 + L // cfg["hop_length"]
    return synth, cfg, L, T


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops")), d.get("hbm_gbs"), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.01)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def time_cpu_port(synth, cfg, batch, T, steps, warmup, seed=11):
    """The reference's CPU implementation of the path (torch-CPU port, oracle/gsn_oracle_torch.py)."""
    from oracle import gsn_oracle_torch as OT
    torch.set_num_threads(os.cpu_count() or 1)
    params = OT.to_torch(synth.make_params(cfg, 5))
    mag = torch.from_numpy(synth.make_mag(batch, cfg["n_fft"] // 2 + 1, T, seed))
    for _ in range(warmup):
        OT.spiking_fullsubnet_network(mag[:, :, : max(8, T // 8)], params, cfg)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        OT.spiking_fullsubnet_network(mag, params, cfg)
        times.append(time.perf_counter() - t0)
    return times, torch.get_num_threads()


def time_cpu_reference(synth, cfg, batch, T, steps, warmup, seed=11):
    """The UNMODIFIED reference (vendored under baseline/_ref by __graft_entry__.build(), see oracle/run_reference.py)
    through its own modules' stock forward code path for the hot path: SpikingFullSubNet.fb_model / .sb_model driven
    exactly as SpikingFullSubNet.forward drives them (modeling_spiking_fullsubnet.py:434-447), eval mode, no_grad, all
    host threads.  Returns (per-step seconds, threads) or None when the vendored copy is not there."""
    from oracle import run_reference as RR
    if not RR.available():
        return None
    torch.set_num_threads(os.cpu_count() or 1)
    model = RR.build_surface_a(cfg, synth.make_params(cfg, 5))
    mag = torch.from_numpy(synth.make_mag(batch, cfg["n_fft"] // 2 + 1, T, seed))
    for _ in range(warmup):
        RR.network_a(model, mag[:, :, : max(8, T // 8)], cfg)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        RR.network_a(model, mag, cfg)
        times.append(time.perf_counter() - t0)
    return times, torch.get_num_threads()


def cpu_leg(synth, cfg, B, T, steps, warmup):
    """CPU arm shared by `--impl reference` and the `cpu_baseline` key: mean over `steps` full-batch passes after
    `warmup` short ones.  kind "reference" = the vendored reference itself; the torch port of it (bit-identical
    results, tests/test_oracle_golden.py; no per-frame weight.repeat, so faster) is reported beside it."""
    ref = time_cpu_reference(synth, cfg, B, T, steps, warmup)
    port_times, cores = time_cpu_port(synth, cfg, B, T, max(1, min(steps, 3)), 1)
    port = B * T / float(np.mean(port_times))
    if ref is None:
        sec = float(np.mean(port_times))
        return sec, {"value": B * T / sec, "unit": "frames/s", "cores": cores, "kind": "port",
                     "sample": f"full batch {B} x T={T} per step, mean of {len(port_times)}; torch-CPU port of the "
                               f"reference path (oracle/gsn_oracle_torch.py): baseline/_ref is not present"}
    times, cores = ref
    sec = float(np.mean(times))
    return sec, {"value": B * T / sec, "unit": "frames/s", "cores": cores, "kind": "reference",
                 "sample": f"full batch {B} x T={T} per step, mean of {len(times)} after {warmup} warm-up; the "
                           f"unmodified reference modules (baseline/_ref) on torch {torch.__version__} CPU, "
                           f"network part (magnitude in -> coefficients out)",
                 "port_value": port, "port_note": "oracle/gsn_oracle_torch.py (bit-identical port without the "
                                                  "per-frame weight.repeat), same batch, mean of "
                                                  f"{len(port_times)}"}


def flops_per_frame(cfg):
    """Dense 2*MAC count of the forward path per frame per utterance: input-to-hidden, hidden-to-hidden
    and proj products of the full-band model and of every sub-band unit."""
    g = 1 if cfg.get("shared_weights", False) else 2
    S = cfg.get("num_spks", 1)
    shapes = [(1, cfg["fb_input_size"], cfg["fb_hidden_size"], cfg["fb_proj_size"], cfg["fb_num_layers"])]
    for i, (ctr, nbr, df) in enumerate(zip(cfg["center_freq_sizes"], cfg["neighbor_freq_sizes"], cfg["df_orders"])):
        n = (cfg["freq_cutoffs"][i + 1] - cfg["freq_cutoffs"][i]) // ctr
        shapes.append((n, 2 * ctr + 2 * nbr, cfg["sb_hidden_size"], 2 * ctr * df * S, cfg["sb_num_layers"]))
    return sum(r * (g * 2 * H * K + (L - 1) * g * 2 * H * H + L * g * 2 * H * H + 2 * H * P) for r, K, H, P, L in shapes)


def train_mode(args, synth, cfg, L, T, rank, world, local, dev):
    """BASELINE configs[3]: one training step per GPU on its own shard of the utterance batch -- forward (train-mode
    BatchNorm) + BPTT through the Triangle surrogate + the recipe's loss (recipes/.../trainer.py:33-37) + AdamW, the
    gradients all-reduced by NCCL through DDP exactly as `accelerator.prepare(model)` sets it up in the reference
    (recipes/intel_ndns/spiking_fullsubnet/run.py:39).  Weak scaling: `--batch` clips per GPU."""
    import torch.distributed as dist
    from spiking_fullsubnet_b200 import SpikingFullSubNet, losses, ops
    B = args.batch
    model = SpikingFullSubNet(**cfg)
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.make_params(cfg, 5).items()})
    model = model.to(dev).train()
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    wave_host = torch.from_numpy(synth.make_wave(B, L, 31 + rank)).pin_memory()
    clean = torch.from_numpy(synth.make_wave(B, L, 41 + rank)).to(dev)
    wave = wave_host.to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)

    def step(x):
        opt.zero_grad(set_to_none=True)
        enh_y = net(x)[0]
        loss = losses.ndns_training_loss(enh_y, clean)["loss"]
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(wave)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ops.LAUNCHES[0] = 0
    evs = []
    for _ in range(args.steps):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(wave)
        e1.record()
        evs.append((e0, e1))
    barrier()
    launches = ops.LAUNCHES[0]
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    # end to end: the step's waveforms come from pinned host memory and the loss value goes back to the host
    t0 = time.perf_counter()
    for _ in range(args.steps):
        x = wave_host.to(dev, non_blocking=True)
        loss_host = float(step(x).detach())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join()
    t = torch.tensor([dev_ms, e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s = float(t[0]), float(t[1])
    if rank == 0:
        tf_peak, _, peak_src = peaks()
        flops = 3.0 * flops_per_frame(cfg) * B * T  # forward + two backward contractions per product
        ms = dev_ms / args.steps
        line = {"metric": "training frames/sec", "value": world * B * T * args.steps / (dev_ms * 1e-3), "unit": "frames/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (training forward on tcgen05 bf16x3 planes; BPTT on fp32 CUDA cores)", "data": "synthetic",
                "config": {"workload": f"intel_ndns spiking_fullsubnet-{args.size} training step (forward + BPTT through the "
                                       f"surrogate gradient + recipe loss + AdamW), batch {B} x {args.seconds:g} s per GPU "
                                       f"(T={T}), DDP / NCCL gradient all-reduce", "size": args.size, "batch_per_gpu": B,
                           "global_batch": world * B, "frames_per_clip": T, "sharding": f"utterance batch x{world}",
                           "l2": "flushed between timed iterations (256 MiB write)",
                           "grad_bytes_allreduced": int(sum(p.numel() for p in model.parameters()) * 4) if world > 1 else 0},
                "clocks": sampler.summary(),
                "e2e": {"value": world * B * T * args.steps / e2e_s, "unit": "frames/s",
                        "h2d_bytes_per_step": int(wave_host.numel() * 4), "d2h_bytes_per_step": 4,
                        "what": "pinned host waveform -> H2D -> training step -> loss value back on the host, every step",
                        "last_loss": loss_host},
                "gpu_launches": launches,
                "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": tf_peak, "unit": "TFLOP/s",
                             "frac": flops / (ms * 1e-3) / 1e12 / tf_peak if tf_peak else None, "traffic": None,
                             "kernel": "whole training step (3 x the forward's algorithmic FLOPs)", "peak_source": peak_src}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    synth, cfg, L, T = workload(args)
    B = args.batch
    wl = {"workload": f"intel_ndns spiking_fullsubnet-{args.size} inference (surface A args), batch {B} x "
                      f"{args.seconds:g} s synthetic clips per GPU, T={T} frames, 257-bin STFT; hot path = "
                      f"magnitude in -> deep-filter coefficients out",
          "size": args.size, "batch_per_gpu": B, "frames_per_clip": T, "sharding": f"utterance batch x{world}",
          "weights": "random-init (synthetic, seed 5), eval-mode BatchNorm"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        sec, cb = cpu_leg(synth, cfg, B, T, args.steps, args.warmup)
        val = B * T / sec
        line = {"impl": "reference", "metric": "frames/sec", "value": val, "unit": "frames/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": wl,
                "cpu_baseline": cb,
                "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from spiking_fullsubnet_b200 import SpikingFullSubNet, ops
    if args.mode == "train":
        return train_mode(args, synth, cfg, L, T, rank, world, local, dev)

    params = synth.make_params(cfg, 5)
    model = SpikingFullSubNet(**cfg)
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in params.items()}, strict=True)
    model = model.eval().to(dev).set_backend(args.backend)
    streaming = False
    if args.schedule != "wavefront" and args.backend in ("auto", "tcgen05"):
        model.enable_streaming(True)
        # co-resident as a whole, or in waves of utterances that are (on request, or when the model's own estimate says
        # waves beat the wavefront / band-stream schedules)
        waves = model._stream_plan(B) is None
        if waves and args.schedule == "stream" and model._stream_wave_size(B) is not None:
            os.environ.setdefault("GSN_STREAM_WAVES", "1")
        streaming = not waves or model._waves_pay(B, T)
        if not streaming:
            if args.schedule == "stream":
                raise SystemExit(f"--schedule stream: size {args.size} batch {B} is not co-resident on this device")
            model.enable_streaming(False)
    mag = torch.from_numpy(synth.make_mag(B, 257, T, 11 + rank)).to(dev)
    wave_host = torch.from_numpy(synth.make_wave(B, L, 21 + rank)).pin_memory()
    out_host = torch.empty((B, L), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        with torch.no_grad():
            return model.network(mag)

    ops.LAUNCHES[0] = 0
    step()  # eager pass: counts the kernels one step launches
    launches_per_step = ops.LAUNCHES[0]
    model.enable_cuda_graph(not args.no_graph, frame_chunks=args.chunks)
    if not args.no_graph:
        step()  # warm-up + capture of the replayed schedule
        launches_per_step = model.graph_launches  # kernels of this library inside one replay of the graph
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    evs = []
    for _ in range(args.steps):
        flush.fill_(1.0)  # L2 flush between timed iterations (untimed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        evs.append((e0, e1))
    barrier()
    launches = launches_per_step * args.steps
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)

    # end to end through the public API: pinned host waveform -> H2D -> forward() -> D2H enhanced waveform.
    # Every step copies its own input in and its own result out; as a serving loop would, the copies of step
    # k+1 / k-1 run on a copy stream while step k computes (double-buffered device + pinned buffers).
    # host->device and device->host copies on their OWN streams (the link is full duplex).  (At 8 GPUs per host the e2e
    # number per GPU is ~25 % below the 1-GPU one either way -- 15.5 - 16.6 against 21.2 M frames/s, r02 -- while the
    # device-timed value scales at 0.99: the host side, not these copies' ordering.)
    copy_s = torch.cuda.Stream(device=dev)
    out_s = torch.cuda.Stream(device=dev)
    comp_s = torch.cuda.current_stream(dev)
    win = [torch.empty((B, L), device=dev) for _ in range(2)]
    wout = [torch.empty((B, L), device=dev) for _ in range(2)]
    hout = [torch.empty((B, L), dtype=torch.float32).pin_memory() for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]

    def e2e_run(n):
        for k in range(n):
            b = k & 1
            if k >= 2:
                with torch.cuda.stream(out_s):
                    out_s.wait_event(ev_done[b])       # step k-2 has produced wout[b] ...
                    hout[b].copy_(wout[b], non_blocking=True)   # ... whose result goes back to the host
                    ev_out[b].record(out_s)
            with torch.cuda.stream(copy_s):
                if k >= 2:
                    copy_s.wait_event(ev_done[b])      # step k-2 has consumed win[b]
                win[b].copy_(wave_host, non_blocking=True)      # this step's input
                ev_in[b].record(copy_s)
            comp_s.wait_event(ev_in[b])
            if k >= 2:
                comp_s.wait_event(ev_out[b])           # wout[b] of step k-2 has left for the host
            with torch.no_grad():
                y = model(win[b])[0]
            wout[b].copy_(y)
            ev_done[b].record(comp_s)
        with torch.cuda.stream(out_s):                 # drain: results of the last two steps
            for k in range(max(0, n - 2), n):
                out_s.wait_event(ev_done[k & 1])
                hout[k & 1].copy_(wout[k & 1], non_blocking=True)
        torch.cuda.synchronize(dev)

    e2e_run(4)
    import gc
    gc.collect()
    gc.disable()  # the wall-clock window is ~15 ms at the default 20 steps and the slowest of N ranks counts
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    e2e_s = time.perf_counter() - t0
    gc.enable()
    sampler.stop_flag = True
    sampler.join()

    # per-kernel timing of the dominant kernel (the recurrence), live, with CUDA events on its stream, every launch
    # timed ALONE (`share_of_step` is the kernel's share of the serial step, i.e. of the sum of all launches run one
    # after the other)
    model.enable_cuda_graph(False)
    per_launch = {}
    if streaming:
        # streaming schedule: the persistent recurrence kernels normally overlap for the whole step; here each one is
        # relaunched without counters on inputs the step has left complete
        model.record_stream_launches = True
        step()
        torch.cuda.synchronize()
        model.record_stream_launches = False
        recs = model.stream_launches
        model.enable_streaming(False)
        model.sb_model.concurrent_bands = False
        step()
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(3):
            step()
        s1.record()
        torch.cuda.synchronize()
        model.sb_model.concurrent_bands = True
        serial_ms = s0.elapsed_time(s1) / 3.0
        rec_ms = rec_flops = 0.0
        for r in recs:
            def one(r=r):
                # (out_cnt: the ring-buffered operand images of a fused layer 0 require the back-pressure counters)
                # (a layer fed by the ring of spike operand images finds only the last frames there when it runs alone:
                # same work, but its output goes to a scratch trace)
                ops.recurrence_stream(r["w_hh"], r["bias"], r["a"], r["b"],
                                      out_bits=torch.empty_like(r["out_bits"]) if r.get("scratch_out") else r["out_bits"],
                                      sm_budget=r["budget"], out_cnt=r["out_cnt"], **r["ins"], **r.get("img", {}))
            one()
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record()
                one()
                b_.record()
                torch.cuda.synchronize()
                ts.append(a_.elapsed_time(b_))
            ms = float(np.mean(ts))
            rec_ms += ms
            rec_flops += r["flops"]
            per_launch.setdefault((r["T"], r["R"], r["H"], r["K_in"], r["layer"] > 0), []).append(ms * 1e3 / r["T"])
        serial_ms = max(serial_ms, rec_ms)
        model.enable_streaming(True)
    else:
        model.sb_model.concurrent_bands = False
        step()
        torch.cuda.synchronize()
        ops.PROFILE = []
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(3):
            step()
        s1.record()
        torch.cuda.synchronize()
        rec = ops.PROFILE
        ops.PROFILE = None
        model.sb_model.concurrent_bands = True
        serial_ms = s0.elapsed_time(s1) / 3.0
        rec_ms = sum(a.elapsed_time(b) for (_, a, b, _) in rec) / 3.0
        rec_flops = sum(f for (f, _, _, _) in rec) / 3.0
        for (_, a, b, (t_, r_, h_)) in rec:
            per_launch.setdefault((t_, r_, h_, 0, False), []).append(a.elapsed_time(b) * 1e3 / t_)
    # latency model of the serial frame chain (SURVEY 8d "Bound"): per launch, measured us per frame against the
    # tensor-pipe time of one frame = 3 planes x ceil(H/16) recurrent tcgen05.mma (+ the fused input product: 3 planes
    # for spike inputs, 8 / 6 plane pairs for the real-valued layer-0 input) at ~17 cycles per 128x16x16 instruction
    # with the A operand in tensor memory (measured inside the kernel: 30 MMAs complete 520 cycles after issue)

    t = torch.tensor([dev_ms, e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    frames = world * B * T * args.steps
    value = frames / (dev_ms * 1e-3)
    tf_peak, hbm_peak, peak_src = peaks()
    achieved = rec_flops / (rec_ms * 1e-3) / 1e12 if rec_ms > 0 else 0.0
    backends = sorted({ops.pick_backend(r, h, cfg["shared_weights"]) if args.backend == "auto" else args.backend
                       for (r, h) in [(B, cfg["fb_hidden_size"])] +
                       [(B * ((cfg["freq_cutoffs"][i + 1] - cfg["freq_cutoffs"][i]) // c), cfg["sb_hidden_size"])
                        for i, c in enumerate(cfg["center_freq_sizes"])]})
    line = {
        "metric": "frames/sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (recurrent weights as exact bf16x3 planes on tcgen05 where that backend runs)",
        "data": "synthetic",
        "config": dict(wl, l2="flushed between timed iterations (256 MiB write)",
                       recurrence_backends=backends,
                       launch=("eager enqueue from Python" if args.no_graph else "CUDA graph replay of the step") +
                              (", streaming pipeline: every (model, layer) recurrence and helper stage is ONE persistent "
                               "kernel for all frames, chained through per-frame counters; layer-0 and layer >= 1 input "
                               "products fused into the recurrences" +
                               (f"; {getattr(model, 'stream_waves', (B, 1))[1]} waves of "
                                f"{getattr(model, 'stream_waves', (B, 1))[0]} utterances, each a co-resident pipeline"
                                if getattr(model, "stream_waves", (B, 1))[1] > 1 else "") if streaming else
                               f", frame-chunked wavefront ({args.chunks} chunks, one stream per model x layer)")),
        "clocks": sampler.summary(),
        "e2e": {"value": world * B * T * args.steps / e2e_s, "unit": "frames/s",
                "h2d_bytes_per_step": int(wave_host.numel() * 4), "d2h_bytes_per_step": int(out_host.numel() * 4),
                "what": "model.forward(wave): pinned host waveform -> H2D -> STFT -> network -> deep filter -> iSTFT "
                        "-> D2H, every step; copies double-buffered on a copy stream"},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s",
                     "frac": achieved / tf_peak if tf_peak else None, "traffic": None,
                     "kernel": "GSN recurrence (all layers of all sequence models of one step" + (", input products fused)" if streaming else ")"),
                     "algorithmic_flops_per_step": rec_flops, "kernel_ms_per_step": rec_ms,
                     "share_of_step": rec_ms / serial_ms, "serial_step_ms": serial_ms,
                     "peak_source": peak_src},
    }
    # dram__bytes_read.sum + dram__bytes_write.sum of the recurrence launches from one `ncu --set full` capture of THIS
    # schedule (tools/ncu_traffic.py writes the file with the schedule / shape it was taken on; anything else -> null)
    traffic_file = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if os.path.exists(traffic_file):
        tr = json.load(open(traffic_file))
        if (tr.get("size"), tr.get("batch"), tr.get("frames"), tr.get("schedule")) == \
                (args.size, B, T, "stream" if streaming else "wavefront"):
            line["roofline"]["traffic"] = tr["dram_bytes_per_step"]
            line["roofline"]["traffic_source"] = tr["source"]
    mhz = (line["clocks"]["sm_mhz"] or 1965.0)

    def mma_per_frame(h_, k_in, spikes_in):
        n = 3 * ((h_ + 15) // 16)
        if k_in:
            ks = (k_in + 15) // 16
            n += (3 if spikes_in else (6 if ks >= 8 else 8)) * ks
        return n

    line["roofline"]["latency_model"] = [
        {"frames": t_, "rows": r_, "hidden": h_, "fused_input": k_, "us_per_frame": float(np.mean(v)),
         "mma_per_frame": mma_per_frame(h_, k_, sp_),
         "mma_us_per_frame": mma_per_frame(h_, k_, sp_) * 17.0 / mhz,
         "tensor_pipe_share_of_frame": mma_per_frame(h_, k_, sp_) * 17.0 / mhz / float(np.mean(v))}
        for (t_, r_, h_, k_, sp_), v in sorted(per_launch.items())]
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_leg(synth, cfg, B, T, 3, 1)[1]
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
