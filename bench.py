#!/usr/bin/env python
"""Benchmark of the FISRnet hot path on B200: 1080p -> 4K tiled inference (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port), rank 0 only

A *step* is one pass of the hot path over one batch of synthetic input: N windows (one per GPU) of three 1080x1920 YUV
frames + 4 flows + 4 warped frames -> N x three 2048x3840 output frames, through the reference's (2,2) tile grid with
its 32-px halo (FISRnet.py:994-1065).  Consecutive windows of a clip share one frame, so a window contributes TWO new
4K frames (2*num_fr - 3 outputs for num_fr inputs, FISRnet.py:1066-1077): value = 2 * windows / second.

  value : inputs already resident in HBM; (window, tile) units sharded tile-major over the ranks, one NCCL all-gather
          of the uint8 tiles per step, frames assembled on every rank.  Weak scaling: 4 units per rank per step.
  e2e   : the same metric through the host-buffer entry point (fisr_window_host <- FISRnet.FISR_for_video): every step
          copies its window from pinned host memory, runs the tiles and copies the uint8 canvas back.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H_IN, W_IN = 1080, 1920          # main.py:100-101 default FISR_input_size
GRID = (2, 2)                    # main.py:102-103 default FISR_test_patch
METRIC = "4K FISR output frames/sec (1080p->4K tiled inference, unique frames: 2 per 3-frame window)"
WORKLOAD = "configs[3]: 1080p->4K tiled inference (FISR_for_video path), (2,2) tiles of 544x992 with 32-px halo"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops_burst": p["bf16_tflops"], "tflops_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (B200_PROFILING.md clocks line)."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def synthetic_windows(n_windows: int, seed: int = 3):
    """uint8 YUV frames (low-passed noise), flow N(0, 4 px), warp = neighbouring frame + noise, like SURVEY 8d config 4."""
    import numpy as np
    rng = np.random.default_rng(seed)
    frames = np.empty((n_windows, H_IN, W_IN, 9), np.uint8)
    for w in range(n_windows):
        small = rng.integers(16, 236, size=(H_IN // 8 + 1, W_IN // 8 + 1, 9)).astype(np.float32)
        up = np.repeat(np.repeat(small, 8, axis=0), 8, axis=1)[:H_IN, :W_IN]
        frames[w] = np.clip(up + rng.normal(0, 6, up.shape), 0, 255).astype(np.uint8)
    flow = (rng.standard_normal((n_windows, H_IN, W_IN, 8)) * 4).astype(np.float32)
    idx = [3, 4, 5, 0, 1, 2, 6, 7, 8, 3, 4, 5]
    warp = (frames[..., idx].astype(np.float32) / 255. + rng.normal(0, 0.02, (n_windows, H_IN, W_IN, 12))).astype(np.float32)
    return frames, flow, warp


# ------------------------------------------------------------------------------------------- CPU baseline (oracle)
def cpu_tile_seconds(budget_s: float, reps: int = 1):
    """Times the oracle (torch-CPU fp32 restatement of FISRnet.model) on one whole 544x992 tile when that fits ~2.5x budget_s
    of CPU work, else on a crop scaled by pixels.  Returns (seconds per full 544x992 tile, description, threads)."""
    import torch
    from oracle import fisrnet_oracle as O          # the checker, here as the reported CPU baseline
    torch.set_num_threads(os.cpu_count() or 1)
    params = O.init_params(seed=0)
    O.model(params, O.synthetic_input(1, 64, 96, 1))                 # warm-up
    t0 = time.perf_counter()
    O.model(params, O.synthetic_input(1, 128, 192, 1))               # calibration
    px_rate = 128 * 192 / (time.perf_counter() - t0)
    full = 544 * 992
    h, w = 544, 992
    while h * w > 64 * 96 and h * w / px_rate > budget_s * 2.5 and h > 64:
        h, w = max(64, (h // 2) // 32 * 32), max(96, (w // 2) // 32 * 32)
    x = O.synthetic_input(1, h, w, 2)
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        O.model(params, x)
        ts.append(time.perf_counter() - t)
    t_crop = sorted(ts)[len(ts) // 2]
    what = "one whole 544x992 tile" if (h, w) == (544, 992) else f"a {h}x{w} crop (scaled by pixels to the 544x992 tile)"
    return t_crop * full / (h * w), f"{reps} x FISRnet.model on {what}, fp32 oneDNN", torch.get_num_threads()


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path.  TensorFlow 1.13 is not installable in
    this image, so this is the oracle port of the identical graph on all host cores.  Every step is ONE WHOLE
    544x992 tile of the (2,2) grid (a quarter of a window; the four tiles of a window are equal-sized and independent,
    FISRnet.py:1028-1057), at every N, so the arm measures the product arm's config and not a pixel-scaled crop."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    import torch
    from oracle import fisrnet_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)                       # torchrun exports OMP_NUM_THREADS=1: set the pool explicitly
    params = O.init_params(seed=0)
    h, w = 544, 992
    x = O.synthetic_input(1, h, w, 2)
    t0 = time.perf_counter()
    O.model(params, x)                                   # calibration pass (also the first warm-up)
    t_cal = time.perf_counter() - t0
    sample = f"one whole {h}x{w} tile per step (1/4 window)"
    if t_cal * (steps + warm) > 1500.0:                  # a very slow host: keep the run bounded, say so
        h, w = 272, 992
        x = O.synthetic_input(1, h, w, 2)
        sample = f"one {h}x{w} half tile per step, scaled by pixels (host too slow for whole tiles: {t_cal:.1f} s each)"
    for _ in range(max(0, warm - 1)):
        O.model(params, x)
    t = time.perf_counter()
    for _ in range(steps):
        O.model(params, x)
    per_step = (time.perf_counter() - t) / steps
    t_window = per_step * (4 * 544 * 992) / (h * w)               # 4 tiles of 544x992 per window
    value = 2.0 / t_window
    line = {"metric": METRIC, "value": value, "unit": "frames/s", "impl": "reference", "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": sample, "same_config": h == 544,
                       "note": "CPU arm: 1 process on all host cores whatever N (the reference has no multi-GPU path)"},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{steps} x FISRnet.model (oracle, torch-CPU fp32 oneDNN), {sample}"},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    args.emit(json.dumps(line))


# ------------------------------------------------------------------------------------------- this repo's arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import fisr_b200
    from fisr_b200 import sharding
    from fisr_b200.init import xavier_params

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    eng = fisr_b200.Engine(local, precision=args.precision)
    eng.set_params(xavier_params(seed=0, bias_std=0.01))           # random-init weights of the reference architecture

    T = GRID[0] * GRID[1]
    B = world                                                       # windows per step: one per GPU (weak scaling)
    frames_h, flow_h, warp_h = synthetic_windows(B)
    frames, flow, warp = (torch.from_numpy(a).to(dev) for a in (frames_h, flow_h, warp_h))
    my_units = sharding.rank_units(rank, world, B, T)
    oh, ow, _ = eng.canvas_shape(H_IN, W_IN, GRID)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    class Exchange:
        """One step = this rank's tile units computed straight into its own frames [B,2048,3840,9] + the exchange that completes
        the frames on every rank.  "p2p" (default): finished tiles are pushed into every peer's frames by the copy engines over
        NVLink peer memory (sharding.PeerFrames), overlapping the next step's kernels with no SM-resident collective.  "nccl":
        unit-major send buffer, one ncclAllGather per step (async, second buffer set), frames re-assembled by a strided copy."""

        def __init__(self, n_windows, units, mode):
            self.B, self.units, self.k = n_windows, units, 0
            self.mode = mode if world > 1 else "single"
            self.peer = None
            if self.mode == "p2p":
                try:
                    self.peer = sharding.PeerFrames(eng, rank, world, n_windows, oh, ow)
                except Exception as e:                     # CUDA IPC unavailable in this container: fall back to NCCL
                    print(f"[bench] peer-memory exchange unavailable ({e}); using the NCCL all-gather", file=sys.stderr)
                    self.mode = "nccl"
                ok = torch.tensor([1 if self.mode == "p2p" else 0], device=dev)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if int(ok.item()) == 0 and self.mode == "p2p":
                    self.peer.close()
                    self.peer, self.mode = None, "nccl"
            if self.mode == "single":
                self.out = [torch.zeros((n_windows, oh, ow, 9), dtype=torch.uint8, device=dev)]
            elif self.mode == "nccl":
                shp = (oh // GRID[0], ow // GRID[1], 9)
                self.local_out = [torch.zeros((len(units),) + shp, dtype=torch.uint8, device=dev) for _ in range(2)]
                self.gathered = [torch.zeros((n_windows * T,) + shp, dtype=torch.uint8, device=dev) for _ in range(2)]
                self.pending = [None, None]
                self.last = None

        def step(self, fr, fl, wp):
            slot = self.k & 1
            self.k += 1
            if self.mode == "single":
                eng.units(fr, fl, wp, self.units, GRID, layout="frames", out=self.out[0])
            elif self.mode == "p2p":
                eng.units(fr, fl, wp, self.units, GRID, layout="frames", out=self.peer.local(slot))
                self.peer.publish(slot, self.units, GRID)
            else:
                eng.units(fr, fl, wp, self.units, GRID, layout="units", out=self.local_out[slot])
                self.pending[slot] = sharding.gather_units(self.local_out[slot], world, out=self.gathered[slot], async_op=True)
                if self.pending[slot ^ 1] is not None:
                    self._finish(slot ^ 1)

        def _finish(self, slot):
            g, work = self.pending[slot]
            self.pending[slot] = None
            if work is not None:
                work.wait()
            self.last = sharding.assemble_frames(g, self.B, GRID)

        def drain(self):
            """Frames of the last step, complete on this rank once every rank has drained (the caller's barrier)."""
            if self.mode == "single":
                return self.out[0]
            if self.mode == "p2p":
                self.peer.drain()
                return self.peer.local((self.k - 1) & 1)
            for slot in ((self.k & 1), (self.k & 1) ^ 1):                # older step first
                if self.pending[slot] is not None:
                    self._finish(slot)
            return self.last

        def close(self):
            if self.peer is not None:
                self.peer.close()

    ex = Exchange(B, my_units, args.exchange)

    def step():
        ex.step(frames, flow, warp)

    def drain():
        return ex.drain()

    warm, steps = max(3, args.warmup), max(1, args.steps)
    for _ in range(warm):
        step()
    out = drain()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    out = drain()
    e1.record()
    barrier()
    launches = eng.launch_count - launches0
    t_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    t_s = float(t_ms.item()) / 1e3
    clocks = sampler.stop() if sampler else None
    value = 2.0 * B * steps / t_s
    checksum = int(out.sum().item())                                # forces / proves a real result (complete frames: after the barrier)
    exchange_mode = ex.mode

    # ---- strong scaling / latency configuration: the tiles of ONE window spread over the ranks (2 windows at N = 8)
    strong = None
    if world > 1:
        B2 = max(1, world // T)
        units2 = sharding.rank_units(rank, world, B2, T)
        ex2 = Exchange(B2, units2, ex.mode)
        fr2, fl2, wp2 = frames[:B2].contiguous(), flow[:B2].contiguous(), warp[:B2].contiguous()
        for _ in range(3):
            ex2.step(fr2, fl2, wp2)
        ex2.drain()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n2 = max(4, min(steps, 10))
        s0.record()
        for _ in range(n2):
            ex2.step(fr2, fl2, wp2)
        out2 = ex2.drain()
        s1.record()
        barrier()
        t2 = torch.tensor([s0.elapsed_time(s1)], device=dev)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ms2 = float(t2.item()) / n2
        strong = {"scaling": "strong", "windows_per_step": B2, "tiles_per_rank_per_step": len(units2), "ms_per_step": ms2,
                  "value": 2.0 * B2 / (ms2 * 1e-3), "unit": "frames/s", "output_checksum": int(out2.sum().item()),
                  "note": "latency configuration: every window's 4 tiles on 4 different ranks; ms_per_step is the time from a "
                          "window's inputs in HBM to its complete frames on every rank (pipelined over steps)"}
        ex2.close()

    # ---- e2e: host buffers -> fisr_window_host -> host canvas, each rank its own window (window-level sharding)
    pin = [torch.from_numpy(a[rank % B]).pin_memory() for a in (frames_h, flow_h, warp_h)]
    canvases = [torch.empty((oh, ow, 9), dtype=torch.uint8).pin_memory() for _ in range(2)]
    pin_np = [p.numpy() for p in pin]
    e2e_steps = max(2, min(steps, 10))
    for k in range(4):       # warm-up through the same pipelined entry points (slot staging buffers, copy streams, events)
        eng.window_submit(k & 1, pin_np[0], pin_np[1], pin_np[2], GRID, out=canvases[k & 1].numpy())
        if k > 0:
            eng.window_wait((k - 1) & 1)
    eng.window_wait(1)
    barrier()
    # two windows in flight (what FISRnet.FISR_for_video does): every window still pays its H2D and D2H inside the timed
    # region, on copy streams that overlap the previous / next window's kernels
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        eng.window_submit(k & 1, pin_np[0], pin_np[1], pin_np[2], GRID, out=canvases[k & 1].numpy())
        if k > 0:
            eng.window_wait((k - 1) & 1)
    eng.window_wait((e2e_steps - 1) & 1)
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = 2.0 * world * e2e_steps / float(t_e2e.item())
    h2d = int(sum(p.numel() * p.element_size() for p in pin))
    d2h = int(canvases[0].numel())
    e2e_checksum = int(canvases[(e2e_steps - 1) & 1].sum().item())

    line = None
    if rank == 0:
        peaks = _peaks()
        # ---- roofline of the dominant kernel (conv3x3_umma_kernel): algorithmic conv FLOPs of the launches of one
        # batched 4-tile forward / their summed CUDA-event durations, measured live, launch by launch
        ops = eng.profile_ops(T, 544, 992, reps=2)
        conv = [o for o in ops if o["kind"] == "conv"]
        conv_ms = sum(o["ms"] for o in conv)
        all_ms = sum(o["ms"] for o in ops)
        conv_flops = sum(o["flops"] for o in conv)
        achieved = conv_flops / (conv_ms * 1e-3) / 1e12
        mma_factor = {"f16x3": 3, "f16f8": 2, "f16": 1}[eng.precision]      # tensor-pipe time per K slice in fp16-rate MMA units
        top = max(conv, key=lambda o: o["ms"])
        # DRAM bytes of the same 138 launches from the committed ncu capture (profiles/r01_traffic.json), next to the
        # algorithmic bytes (every input plane, weight and output once per launch)
        traffic, traffic_src = None, None
        try:
            import glob
            cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
            with open(cands[-1]) as f:
                t = json.load(f).get(eng.precision)
            if t:
                traffic, traffic_src = t["dram_bytes_read"] + t["dram_bytes_write"], os.path.relpath(cands[-1], ROOT)
        except (OSError, ValueError, KeyError, IndexError):
            pass
        alg_bytes = sum(o["bytes"] for o in conv)
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                    "frac": achieved / peaks["tflops_sustained"], "traffic": traffic, "algorithmic_bytes": alg_bytes,
                    "traffic_note": f"DRAM read+write bytes of the conv launches of one 4-tile forward, summed, from the committed ncu capture {traffic_src} (ncu cannot run inside the timed bench)",
                    "kernel": f"conv3x3_umma_kernel ({len(conv)} launches per forward, summed)",
                    "peak_source": peaks["source"] + ", bf16 sustained (fp16 operands run at the bf16 rate)",
                    "issued_tflops": achieved * mma_factor, "issued_frac": achieved * mma_factor / peaks["tflops_sustained"],
                    "conv_share_of_forward": conv_ms / all_ms, "forward_ms_launch_by_launch": all_ms,
                    "slowest_launch": {"name": top["name"], "ms": top["ms"], "tflops": top["flops"] / top["ms"] / 1e9}}
        # ---- CPU baseline beside it (bounded sample of the same workload on this box's host cores)
        info = eng.plan_info(T, 544, 992)
        cpu_baseline = None
        extra = None
        if world == 1:          # rank 0 at N = 1 only: at N > 1 the other ranks spin in the barrier on the same host cores
            t_tile, sample, cores = cpu_tile_seconds(budget_s=args.cpu_budget)
            cpu_baseline = {"value": 2.0 / (4 * t_tile), "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample}
            if not args.no_extras:
                extra = measure_extras(eng, dev, peaks, args.precision)
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": warm,
                "ms_per_step": t_s / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"f16x3": "f32 via fp16 (hi,lo) split operands, fp32 accumulate", "f16f8": "f32 via fp16 main term + fp8 cross terms (e5m2 activations x e4m3 / e5m2 weights), fp32 accumulate", "f16": "f16 operands, f32 accumulate"}[eng.precision],
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "windows_per_step": B, "units_per_rank_per_step": len(my_units),
                           "input": f"{B} x (frames u8 [1080,1920,9] + flow f32 [..,8] + warp f32 [..,12])",
                           "output": f"{B} x uint8 [2048,3840,9]", "precision": eng.precision, "weights": "random-init (Xavier)",
                           "sharding": {"single": "single GPU, 4 tiles batched",
                                        "p2p": "tile-major (window,tile) units computed into frame layout; finished tiles pushed to every peer's frames by the copy engines over NVLink peer memory (all-gather without an SM-resident collective)",
                                        "nccl": "tile-major (window,tile) units, one NCCL all-gather of uint8 tiles per step + strided re-assembly"}[exchange_mode],
                           "l2": "inputs (185 MB/window) and the 35 GB activation workspace exceed the 126 MB L2; no flush needed",
                           "windows_per_s": B * steps / t_s, "conv_gflop_per_window": info["flops"] / 1e9,
                           "mma_row_efficiency": info["mma_row_efficiency"], "output_checksum": checksum},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps, "api": "Engine.window_submit/window_wait -> fisr_window_submit/_wait (pinned host buffers, 2 windows in flight, as FISRnet.FISR_for_video)",
                        "output_checksum": e2e_checksum},
                "gpu_launches": int(launches),
                "roofline": roofline,
                "cpu_baseline": cpu_baseline, "extra": extra, "strong_scaling": strong}
    barrier()
    ex.close()
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    if line is not None:
        args.emit(json.dumps(line))


def measure_extras(eng, dev, peaks, bench_precision):
    """The other BASELINE.json configs and the second precision mode, N = 1 only: config 2 forward,
    config 3 training step INCLUDING Adam, the fp32-class mode on the bench workload, and the flow-warp kernel against the
    HBM roof.  Device-timed with CUDA events on the library's stream unless the call is synchronous by contract."""
    import torch
    out = {}
    peak_tf, peak_gb = peaks["tflops_sustained"], peaks["hbm_gbs"]

    def timed(fn, warm, reps):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # ---- flow warp (FISR_for_video_warp_img_with_flo.py:61-67,112-128): the 16 warps of a 9-frame 1080p clip in one launch
    # (fisr_warp_batch_device): 763 MB of traffic per launch, far beyond the 126 MB L2
    g = torch.Generator(device="cpu").manual_seed(5)
    nfr = 9
    jobs = 2 * (nfr - 1)
    yuv = torch.randint(0, 256, (nfr, H_IN, W_IN, 3), dtype=torch.uint8, generator=g).to(dev)
    flo = (torch.randn(jobs, H_IN, W_IN, 2, generator=g) * 4).to(dev)
    src = [fr + 1 - (j & 1) for fr in range(nfr - 1) for j in range(2)]
    warped = torch.empty((jobs, H_IN, W_IN, 3), dtype=torch.float32, device=dev)      # output allocated once: no allocator work in the timed loop
    ms = timed(lambda: eng.warp_batch(yuv, flo, src, 0.5, 1.0 / 255.0, out=warped), 3, 10)
    del warped
    px = jobs * H_IN * W_IN
    by = px * (3 + 8 + 12)                      # u8 YUV source + f32 flow + f32 output, each once
    out["warp_yuv_1080p"] = {"ms_per_launch": ms, "warps_per_launch": jobs, "us_per_1080p_warp": ms * 1e3 / jobs,
                             "algorithmic_bytes": by, "bytes_per_px": 23, "gbs": by / ms / 1e6,
                             "frac_of_hbm_peak": by / ms / 1e6 / peak_gb,
                             "gbs_at_reference_layout_32B_per_px": px * 32 / ms / 1e6,
                             "note": "Engine.warp_batch: 16 warps (9-frame clip) per launch, 10 launches, CUDA events; 23 B/px = u8 source + f32 flow + f32 output"}
    del yuv, flo

    # ---- PWC-Net (the flow step of BASELINE configs[4]): both directions of one 1080p frame pair after the
    # reference's x2 pre-upscale, padded to multiples of 64 (..pwcnet_predict_from_img_test.py:126-131, model_pwcnet.py:371-409)
    try:
        import numpy as np
        from fisr_b200.pwcnet import PWCNet, param_inventory as pwc_inventory
        rng = np.random.default_rng(7)
        net = PWCNet(dev.index)
        net.set_params({n: ((rng.standard_normal(sh) * (2.0 / max(1, int(np.prod(sh[:3])))) ** 0.5) if len(sh) == 4 else np.zeros(sh)).astype(np.float32)
                        for n, sh in pwc_inventory().items()})
        a = torch.rand(2, 2176, 3840, 3, generator=g).to(dev)
        b = torch.rand(2, 2176, 3840, 3, generator=g).to(dev)
        n0 = net.launch_count
        flow_out = torch.empty((2, 2176, 3840, 2), dtype=torch.float32, device=dev)
        ms = timed(lambda: net.forward(a, b, out=flow_out), 2, 5)
        del flow_out
        out["pwcnet_1080p_pair_x2"] = {"ms": ms, "launches_per_forward": (net.launch_count - n0) // 7, "input": "2 x [2176,3840,3] (both directions of one pair)",
                                       "note": "PWC-Net-large (6 levels, flow at level 2, dense + residual connections): stride-1 3x3 convs on the tcgen05 split-mode (fp16 hi/lo, fp32-class) kernel, the rest on CUDA cores; random weights"}
        # the whole per-pair job of FISR_for_video_Compute_Flow from host frames: upload of two uint8 YUV frames, pre-processing, network,
        # post-processing (all on the device), download of the [2,1080,1920,2] flow
        y1 = rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
        y2 = rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
        net.flow_pair_yuv(y1, y2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            net.flow_pair_yuv(y1, y2)
        out["pwcnet_1080p_pair_x2"]["flow_pair_host_to_host_ms"] = (time.perf_counter() - t0) / 3 * 1e3
        # the same as a pipeline over a 7-frame clip: every frame uploaded once, the download of pair k under the kernels of pair k + 1
        clip = [y1, y2, y1, y2, y1, y2, y1]
        for _ in net.flow_sequence_yuv(iter(clip[:3])):
            pass
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in net.flow_sequence_yuv(iter(clip)):
            pass
        out["pwcnet_1080p_pair_x2"]["flow_sequence_ms_per_pair"] = (time.perf_counter() - t0) / 6 * 1e3
        net.close()
        del a, b
    except Exception as e:                       # a next-row component must not take the headline line down with it
        out["pwcnet_1080p_pair_x2"] = {"error": str(e)[:200]}

    # ---- config 2: img [8,192,192,29] forward (FISRnet.py:747-748), both precision modes
    x = torch.rand(8, 192, 192, 29, generator=g).to(dev)
    for prec in ("f16f8", "f16x3"):
        eng.set_precision(prec)
        info = eng.plan_info(8, 192, 192)
        ms = timed(lambda: eng.forward(x), 3, 20)
        tf = info["flops"] / ms / 1e9
        out[f"config2_forward_{prec}"] = {"ms": ms, "gflop": info["flops"] / 1e9, "tflops": tf, "frac": tf / peak_tf,
                                          "workspace_gb": info["workspace_bytes"] / 1e9,
                                          "note": f"Engine.forward on a device tensor: pack + {info['launches']} launches (CUDA graph) + 3 output copies; activation workspace > L2"}

    # ---- config 3: one training step, forward + loss + backward + Adam: B = 16 at LR 192x192 (4 weight-shared passes = 64 images), and the
    # reference's own default patch (LR 96x96 from 192x192 HR labels, B = 8: main.py:48-51, FISRnet.py:187)
    eng.set_precision("f16x3")
    for key, B, hh in (("config3_train_step_f16x3", 16, 192), ("config3_native_patch96_B8_train_step_f16x3", 8, 96)):
        batch = [torch.rand(B, hh, hh, 15, generator=g), (torch.randn(B, hh, hh, 16, generator=g) * 4 / 96 / 2).clamp(-1, 1),
                 (torch.randn(B, hh, hh, 8, generator=g) * 8 / 96 / 2).clamp(-1, 1), torch.rand(B, hh, hh, 24, generator=g),
                 torch.rand(B, hh, hh, 12, generator=g), torch.rand(B, 2 * hh, 2 * hh, 21, generator=g)]
        batch = [t.to(dev) for t in batch]
        eng.adam_reset(0)
        n0 = eng.launch_count
        for _ in range(2):
            eng.train_step(*batch, lr=1e-6)
        torch.cuda.synchronize()
        n1 = eng.launch_count
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            s_ = eng.train_step(*batch, lr=1e-6)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / reps * 1e3
        fwd = eng.plan_info(4 * B, hh, hh)["flops"]
        out[key] = {"ms": ms, "batch": B, "lr_patch": hh, "algorithmic_tflop": 3 * fwd / 1e12, "tflops": 3 * fwd / ms / 1e9,
                    "frac": 3 * fwd / ms / 1e9 / peak_tf, "launches_per_step": (n1 - n0) // 2, "total_loss": s_["total_loss"],
                    "note": "fisr_train_step: window assembly, 4B-image forward, multi-scale temporal loss, dgrad + wgrad, "
                            "multi-tensor TF-1.13 Adam + operand re-pack; host-timed around the synchronous calls (the 11 loss scalars are read back every step, like sess.run)"}
        eng.adam_reset(0)
        del batch

    # ---- the other precision mode on the bench workload (4 tiles of 544x992, inputs in HBM, and through host buffers)
    other = "f16x3" if bench_precision == "f16f8" else "f16f8"
    eng.set_precision("f16")          # drops the f16x3 plans (training workspace) before the 35 GB tile plan is built
    eng.set_precision(other)
    frames_h, flow_h, warp_h = synthetic_windows(1)
    frames, flow, warp = (torch.from_numpy(a[0]).to(dev) for a in (frames_h, flow_h, warp_h))
    canvas = torch.zeros(eng.canvas_shape(H_IN, W_IN, GRID), dtype=torch.uint8, device=dev)
    ms = timed(lambda: eng.window(frames, flow, warp, GRID, out=canvas), 3, 8)
    pin = [torch.from_numpy(a[0]).pin_memory().numpy() for a in (frames_h, flow_h, warp_h)]
    outs = [torch.empty(tuple(canvas.shape), dtype=torch.uint8).pin_memory().numpy() for _ in range(2)]
    for kk in range(3):
        eng.window_submit(kk & 1, pin[0], pin[1], pin[2], GRID, out=outs[kk & 1])
        if kk > 0:
            eng.window_wait((kk - 1) & 1)
    eng.window_wait(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n_e2e = 6
    for kk in range(n_e2e):
        eng.window_submit(kk & 1, pin[0], pin[1], pin[2], GRID, out=outs[kk & 1])
        if kk > 0:
            eng.window_wait((kk - 1) & 1)
    eng.window_wait((n_e2e - 1) & 1)
    torch.cuda.synchronize()
    e2e = 2.0 * n_e2e / (time.perf_counter() - t0)
    info = eng.plan_info(4, 544, 992)
    out[f"config4_{other}"] = {"value": 2.0 / (ms * 1e-3), "unit": "frames/s", "ms_per_window": ms, "e2e": e2e,
                               "tflops": info["flops"] / ms / 1e9, "frac": info["flops"] / ms / 1e9 / peak_tf,
                               "note": "same workload and units as the headline line, in the other precision mode"}
    eng.set_precision(bench_precision)
    return out


def _json_only_stdout():
    """Keeps stdout for the ONE JSON line: everything else that writes to file descriptor 1 (NCCL prints its version banner
    there at communicator creation) is sent to stderr.  Returns the function that emits the line."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)

    def emit(text: str) -> None:
        sys.stdout.flush()
        os.write(keep, (text + "\n").encode())

    return emit


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="f16f8", choices=["f16x3", "f16", "f16f8"],
                    help="f16f8 (default) = fp16 main term + fp8 cross terms, 3-6e-5 max-abs vs the fp64 oracle (north-star bar 1e-3); "
                         "f16x3 = fp32-class mode (what training uses); f16 = fast mode, outside the bar")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: how the frames become complete on every rank -- p2p (copy engines over NVLink peer memory, default) or one NCCL all-gather per step")
    ap.add_argument("--no-extras", action="store_true", help="skip the `extra` block (configs 2 and 3, warp kernel, other precision mode)")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for the cpu_baseline sample")
    args = ap.parse_args()
    args.emit = _json_only_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
