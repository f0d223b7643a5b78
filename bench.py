#!/usr/bin/env python
"""bench.py -- images/s of BiRefNet Swin-L `forward_logits` @1024^2 on N B200s (BASELINE.json metric).

A "step" is one pass of the hot path (BiRefNet::forward_logits, src/birefnet.rs:412-461) over one batch of 16
synthetic 1024x1024 images per GPU (BASELINE.json configs[2]: "Full BiRefNet Swin-L 1024^2 batch 16 bf16 on 1 B200").
Weights are the seeded random-init Swin-L set (weight-set B: random deformable offsets); there is no network.

  value     images/s, whole job, inputs already resident in HBM (device pointers, CUDA events on the launch stream,
            max over ranks).  Inputs rotate over 3 batches of 201 MB each (> 126 MB L2), activations are GBs.
  e2e       the same metric through the C ABI with HOST buffers (pinned): H2D of the batch and D2H of the masks are
            inside the timed region.
  roofline  dominant kernel class (tcgen05 implicit GEMM): algorithmic FLOPs / summed device time of its launches,
            measured live with CUDA events around every launch (brn_profile_enable(m, 2)) in a separate,
            untimed step; peak = MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step).
  cpu_baseline  the oracle (PyTorch CPU restatement of the reference, kind "port": candle cannot be built here) timed
            on this box's host cores on a bounded sample (1 image of the same 1024^2 workload), rank 0, N=1 only.

`--impl reference` times only that CPU restatement (all host threads) and prints the same JSON line.
Multi-GPU: one process per GPU (torchrun), images sharded, no data-path collective ("scaling": "weak").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

GFLOP_PER_IMAGE_1024 = 2534.9   # SURVEY.md Appendix B (MAC x2, 1024^2)


def load_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        p = json.loads(f.read_text())
        return dict(hbm_gbs=p["hbm_gbs"], tflops_burst=p["bf16_tflops"], tflops_sustained=p["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # keep the samples taken under load (upper half of the clock range)
        load = [s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)] if pw else sm
        return dict(sm_mhz=statistics.median(load) if load else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(pw) if pw else None, samples=len(sm), reasons=sorted(reasons))


def oracle_images_per_s(H: int, W: int, steps: int, warmup: int, budget_s: float = 240.0):
    """Times the CPU restatement of the reference (deform_mode=cpu_fallback: what candle computes on Device::Cpu,
    src/aspp.rs:183-185) on one image per step.  Bounded: stops early when the budget is spent."""
    import torch
    from oracle import birefnet_ref as R
    from oracle.make_weights import as_torch, make_input, make_weights
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = R.Config.swin_l()
    w = as_torch(make_weights(cfg, seed=0, weight_set="A"))
    x = torch.from_numpy(make_input(1, H, W, seed=1234))
    t_start = time.time()
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.time()
            R.forward_logits(x, w, cfg, "cpu_fallback")
            dt = time.time() - t0
            if i >= warmup:
                times.append(dt)
            elapsed = time.time() - t_start
            if times and elapsed + dt > budget_s:
                break
            if not times and i + 1 >= warmup:
                continue
    if not times:   # budget exhausted inside warm-up: the last warm-up run is the sample
        times = [dt]
    med = statistics.median(times)
    return 1.0 / med, cores, len(times), med


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    H = W = args.size
    warm = min(args.warmup, 1)
    ips, cores, nsteps, med = oracle_images_per_s(H, W, args.steps, warm)
    line = {
        "impl": "reference", "metric": "images/s BiRefNet Swin-L @1024^2", "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": nsteps, "warmup": warm, "ms_per_step": med * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BiRefNet swin_l forward_logits {H}x{W}, batch 16 per GPU (BASELINE.json configs[2])",
                   "implementation": "reference CPU path: PyTorch-CPU restatement of candle's CPU forward (candle itself "
                                     "cannot be built here: no Rust toolchain)",
                   "sample": "1 image per step (1/16 of a batch)", "deform_mode": "cpu_fallback"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"1 image {H}x{W} per step, {nsteps} timed step(s)"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="images per GPU per step")
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--deform-mode", default="deformable", choices=["deformable", "cpu_fallback"])
    ap.add_argument("--e2e-threads", type=int, default=2, help="host threads issuing e2e calls on the one handle")
    ap.add_argument("--dev-streams", type=int, default=1, help="streams the device-resident steps alternate over")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--kernel-log", default="", help="write a per-launch CSV (class, ms, gflop, desc) of one step")
    ap.add_argument("--model", default="swin_l", choices=["swin_l", "mini"], help="mini is for smoke runs only")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import candle_birefnet_b200 as cb
    from candle_birefnet_b200.synth import synthetic_input as make_input

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    warmup = max(args.warmup, 3)
    H = W = args.size
    B = args.batch

    swin = cb.SwinConfig.swin_l() if args.model == "swin_l" else cb.SwinConfig(embed_dim=64, depths=(2, 2, 2, 2),
                                                                               num_heads=(2, 4, 8, 16))
    pcfg = cb.BiRefNetConfig(swin=swin, precision=args.precision, deform_mode=args.deform_mode)
    model = cb.BiRefNet.new_synthetic(pcfg, seed=0, weight_set="B", offset_sigma=2.0, device=local)

    # ---- inputs: 3 rotating batches, device-resident for `value`, pinned host copies for `e2e` ----
    nrot = 3
    host_in = [torch.from_numpy(make_input(B, H, W, seed=1234 + 7 * rank + i)).pin_memory() for i in range(nrot)]
    dev_in = [h.cuda(non_blocking=True) for h in host_in]
    dev_out = torch.empty((B, 1, H, W), dtype=torch.float32, device="cuda")
    host_out = torch.empty((B, 1, H, W), dtype=torch.float32).pin_memory()
    stream = torch.cuda.Stream()          # a real (non-default) stream: kernels and the timing events share it
    torch.cuda.set_stream(stream)

    nds = max(1, args.dev_streams)
    dstreams = [stream] + [torch.cuda.Stream() for _ in range(nds - 1)]
    dev_outs = [dev_out] + [torch.empty_like(dev_out) for _ in range(nds - 1)]

    def step_dev(i):
        model.forward_logits(dev_in[i % nrot], out=dev_outs[i % nds], stream=dstreams[i % nds].cuda_stream)

    # e2e: `e2e_threads` host threads share the handle; each call is synchronous (returns with the masks in its
    # pinned host buffer), the handle keeps two calls in flight so the copies of one overlap the kernels of the other
    nthr = max(1, args.e2e_threads)
    host_outs = [host_out] + [torch.empty((B, 1, H, W), dtype=torch.float32).pin_memory() for _ in range(nthr - 1)]

    def step_e2e(i, t=0):
        import ctypes as C
        h = host_in[i % nrot]
        cb._lib.check(cb.lib().brn_forward_logits(model._h, C.c_void_p(h.data_ptr()), B, H, W, 0,
                                                  C.c_void_p(host_outs[t].data_ptr()), 0, None))

    def run_e2e(steps):
        """`steps` calls spread over the host threads (shared counter)."""
        nxt = iter(range(steps))
        lock, errs = threading.Lock(), []

        def work(t):
            try:
                while True:
                    with lock:
                        i = next(nxt, None)
                    if i is None:
                        return
                    step_e2e(i, t)
            except Exception as e:     # noqa: BLE001 - re-raised below
                errs.append(e)
        ts = [threading.Thread(target=work, args=(t,)) for t in range(nthr)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        if errs:
            raise errs[0]

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """(device ms, host wall ms) of `steps` calls, both max over ranks; barrier + synchronize on both sides, the
        host clock runs strictly inside the barriers."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record(stream)
        for ds in dstreams[1:]:
            ds.wait_stream(stream)
        for i in range(steps):
            fn(i)
        for ds in dstreams[1:]:
            stream.wait_stream(ds)
        e1.record(stream)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - w0) * 1e3
        ms = e0.elapsed_time(e1)
        if dist:
            t = torch.tensor([ms, wall_ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall_ms = float(t[0].item()), float(t[1].item())
        barrier()
        return ms, wall_ms

    # every rotating input buffer is seen twice before timing: the second sighting of a (buffers, shape) key is where
    # the library captures its CUDA graph; capture must not land inside the timed region
    for i in range(max(warmup, 2 * nrot * nds)):
        step_dev(i)
    torch.cuda.synchronize()
    model.reset_launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total, _ = timed(step_dev, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = model.launch_count()
    value = world * B * args.steps / (ms_total / 1e3)

    # ---- e2e through the C ABI with host buffers ----
    run_e2e(4 * nthr)                      # each lane's staging buffers are seen twice: CUDA-graph capture happens here
    run_e2e(2 * nthr)
    # host-pointer calls synchronise inside the call: the host wall clock (max over ranks) is the figure
    barrier()
    w0 = time.perf_counter()
    run_e2e(args.steps)
    torch.cuda.synchronize()
    e2e_wall = (time.perf_counter() - w0) * 1e3
    if dist:
        t = torch.tensor([e2e_wall], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_wall = float(t[0].item())
    barrier()
    e2e_s = e2e_wall / 1e3
    e2e_value = world * B * args.steps / e2e_s

    # ---- roofline of the dominant kernel class, live CUDA events around every launch (untimed step) ----
    peaks = load_peaks()
    roof, classes = None, None
    if rank == 0:
        model.profile(2)
        step_dev(0)                        # first pass creates the event pool
        torch.cuda.synchronize()
        # three profiled passes, keep the one with the smallest kernel-time total (a pass now and then contains one
        # multi-ms outlier on an arbitrary launch: a host-side hiccup between two event records, not kernel time)
        best = None
        for rep in range(3):
            if args.kernel_log:
                os.environ["BRN_KERNEL_LOG"] = args.kernel_log + (".tmp%d" % rep)
            step_dev(0)
            torch.cuda.synchronize()
            os.environ.pop("BRN_KERNEL_LOG", None)
            c = model.kernel_class_times()
            t = sum(v["ms"] for v in c.values())
            if best is None or t < best[0]:
                best = (t, c, model.profile_get(), rep)
        if args.kernel_log:
            for rep in range(3):
                tmp = args.kernel_log + (".tmp%d" % rep)
                if os.path.exists(tmp):
                    if rep == best[3]:
                        os.replace(tmp, args.kernel_log)
                    else:
                        os.remove(tmp)
        classes, stages = best[1], best[2]
        model.profile(0)
        g = classes["gemm_tcgen05"] if classes["gemm_tcgen05"]["launches"] else classes["gemm_simt"]
        tot_ms = sum(c["ms"] for c in classes.values())
        achieved = g["flops"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
        # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (profiles/), if present
        traffic, traffic_note = None, None
        tf = ROOT / "profiles" / "traffic.json"
        if tf.exists():
            tj = json.loads(tf.read_text())
            traffic, traffic_note = tj.get("dram_bytes_per_launch"), tj.get("note")
        roof = {"bound": "tensor", "kernel": "tc_gemm_kernel (tcgen05 implicit GEMM)", "achieved": achieved,
                "peak": peaks["tflops_sustained"], "peak_source": peaks["source"] + " bf16_tflops_sustained",
                "unit": "TFLOP/s", "frac": achieved / peaks["tflops_sustained"], "traffic": traffic,
                "traffic_note": traffic_note,
                "launches_per_step": g["launches"], "avg_launch_ms": g["ms"] / max(g["launches"], 1),
                "share_of_kernel_time": g["ms"] / tot_ms if tot_ms else None,
                "classes_ms": {k: round(v["ms"], 3) for k, v in classes.items()},
                "classes_tflops": {k: (round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1) if v["ms"] > 0 and v["flops"] else None)
                                   for k, v in classes.items()},
                "classes_gbs": {k: (round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 0) if v["ms"] > 0 and v.get("bytes") else None)
                                for k, v in classes.items()},
                "hbm_peak_gbs": peaks["hbm_gbs"],
                "stages_ms": {n: round(ms, 3) for n, ms in stages}}

    # ---- p50 batch-1 latency (device-resident input) ----
    lat = None
    if rank == 0 and not args.no_latency:
        x1 = dev_in[0][:1].contiguous()
        o1 = torch.empty((1, 1, H, W), dtype=torch.float32, device="cuda")
        for _ in range(3):
            model.forward_logits(x1, out=o1, stream=stream.cuda_stream)
        ts = []
        for _ in range(15):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            model.forward_logits(x1, out=o1, stream=stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        lat = statistics.median(ts)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.model == "swin_l":
        ips, cores, n, med = oracle_images_per_s(H, W, 5, 1, budget_s=45.0)
        cpu_base = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                    "sample": f"{n} x 1 image {H}x{W} (each 1/{B} of a step) after 1 warm-up, median {med:.1f} s per image; "
                              f"PyTorch-CPU restatement of the reference's CPU forward (deform_mode=cpu_fallback)"}

    if rank == 0:
        gflop = GFLOP_PER_IMAGE_1024 * (H * W) / (1024.0 * 1024.0)
        line = {
            "metric": "images/s BiRefNet Swin-L @1024^2", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp16": "fp16", "bf16": "bf16", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": {"workload": f"BiRefNet {args.model} forward_logits {H}x{W}, batch {B} per GPU "
                                   f"(BASELINE.json configs[2])",
                       "global_batch": B * world, "parallelism": f"image-sharded x{world}, no collective",
                       "precision": args.precision, "deform_mode": args.deform_mode,
                       "weights": "seeded random-init, weight-set B", "l2": "inputs rotate over 3 batches of "
                       f"{B * 3 * H * W * 4 / 1e6:.0f} MB (> 126 MB L2)"},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * 3 * H * W * 4,
                    "d2h_bytes_per_step": B * H * W * 4, "ms_per_step": e2e_s * 1e3 / args.steps,
                    "host_threads": nthr},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "model_tensor_frac": (value / world) * gflop / 1e3 / peaks["tflops_sustained"],
            "latency_ms_p50_b1": lat,
            "cpu_baseline": cpu_base,
        }
        print(json.dumps(line))
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    model.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
