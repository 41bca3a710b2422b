#!/usr/bin/env python
"""bench.py -- throughput of the MergeRec hot paths on B200, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the hot path over one batch of synthetic input (random-init weights of the named
architecture; there is no network for checkpoints).  `value` is the whole-job throughput with inputs already
in HBM; `e2e` is the same metric through the public API with HOST buffers (pinned) and the host<->device
copies inside the timed region.  `roofline` times the dominant kernel alone with CUDA events;
`cpu_baseline` times the CPU oracle (a port of the reference's algorithm, oracle/) on the box's host cores.
`--impl reference` runs only that CPU arm.  See DESIGN.md section "Measurement".
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

import numpy as np  # noqa: E402
import torch  # noqa: E402


# ------------------------------------------------------------------------------------------------ utilities
_JSON_FD = None


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=float(p["hbm_gbs"]), bf16_tflops=float(p["bf16_tflops"]),
                    bf16_tflops_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        threading.Thread(target=self._pump, daemon=True).start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))


def use_all_host_threads() -> int:
    """The CPU arms use every host core: torchrun exports OMP_NUM_THREADS=1 to its workers, which would otherwise
    serialise the OpenMP oracle and numpy's BLAS.  Returns the thread count in effect."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    os.environ["OMP_NUM_THREADS"] = str(n)
    from oracle import oracle as orc
    orc.set_threads(n)
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=n)
    except Exception:  # noqa: BLE001
        pass
    torch.set_num_threads(n)
    return n


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def event_time_ms(fn, iters: int) -> float:
    """Average device time of fn() over `iters` calls (after one untimed call), CUDA events on the current stream."""
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / iters


# ------------------------------------------------------------------------------------------------ workloads
from bench_workloads import WORKLOADS, DEFAULT_WORKLOAD  # noqa: E402


def run_ours(args):
    rank, local_rank, world = dist_env()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wl = WORKLOADS[args.workload](rank=rank, world=world, device=torch.device("cuda", local_rank))
    wl.setup()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks are sampled from the first warm-up step to the end of the last measured loop (the timed region
    # itself can be a few ms, shorter than one nvidia-smi sample period)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        wl.step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        wl.step()
    ev1.record()
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    if hasattr(wl, "finish"):
        wl.finish()   # deferred host-side checks of the timed steps (e.g. the TIES select status word)

    # end to end through the public API: pinned host inputs -> H2D -> kernels -> D2H result, every step
    wl.setup_e2e()
    for _ in range(min(args.warmup, 2)):
        wl.step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, wl.e2e_steps_cap))
    ev0.record()
    for _ in range(e2e_steps):
        wl.step_e2e()
    ev1.record()
    barrier()
    e2e_wall = time.perf_counter() - t0
    ms2 = torch.tensor([max(ev0.elapsed_time(ev1), e2e_wall * 1e3)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms = float(ms2.item()) / e2e_steps

    roof = wl.roofline(measured_peaks()) if rank == 0 else None
    clocks = sampler.stop() if rank == 0 else None
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        use_all_host_threads()
        cpu = wl.cpu_baseline()

    if rank == 0:
        ms_per_step = total_ms / args.steps
        units = wl.units_per_step_all_ranks()
        line = {
            "metric": wl.metric, "value": units / (ms_per_step * 1e-3), "unit": wl.unit, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": wl.scaling, "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
            "config": wl.config(), "clocks": clocks,
            "e2e": {"value": units / (e2e_ms * 1e-3), "unit": wl.unit, "h2d_bytes_per_step": wl.h2d_bytes,
                    "d2h_bytes_per_step": wl.d2h_bytes, "ms_per_step": e2e_ms, "steps": e2e_steps},
            "gpu_launches": wl.launches_per_step * args.steps,
            "roofline": roof,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        line.update(wl.extra())
        if hasattr(wl, "Q"):
            line["eval_seqs_per_s"] = wl.Q / (ms_per_step * 1e-3)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """The reference's own CPU implementation of the path (the oracle port of it: the reference is pure
    Python/torch and does not travel to the GPU box), all host threads, same metric/config."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    use_all_host_threads()
    wl = WORKLOADS[args.workload](rank=0, world=1, device=None)
    res = wl.reference_arm(steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": wl.metric, "value": res["value"], "unit": wl.unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": wl.scaling, "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic", "config": wl.config(),
        "cpu_baseline": {"value": res["value"], "unit": wl.unit, "cores": res["cores"], "kind": "port",
                         "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main():
    # Only the JSON line may reach stdout: libraries (NCCL with NCCL_DEBUG set, ...) print there from C code, so
    # file descriptor 1 is pointed at stderr for the whole run and the line is written to the saved descriptor.
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("note: fewer than 3 warm-up steps requested; the timing rules ask for >= 3")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
