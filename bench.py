#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric) on N GPUs of one node.

A "step" = one pass of the hot path over one batch of synthetic OCP instances: cold Controller::step for every
instance of BASELINE config[1] (batch=1024 per GPU, unicycle quadratic-form, N=50, 5 circular obstacles).

  python bench.py --gpus 1 --steps K --warmup W                       (single GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (one rank per GPU)
  python bench.py --impl reference ...                                (CPU arm: the oracle port on the host cores)

Rank 0 prints ONE JSON line.  `value` = whole-job converged solves/s with the inputs already resident in HBM;
`e2e` = the same metric through the C-ABI call with pinned HOST buffers (H2D + D2H inside the timed region);
`roofline` = the Riccati KKT kernel's algorithmic bytes / its CUDA-event time, against the measured HBM peak;
`cpu_baseline` = the CPU oracle (port of the algorithm, test infrastructure) on a bounded sample of the workload.
"""
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

METRIC = "converged MPC solves/sec (N=50 unicycle quadratic-form OCP, 5 obstacles, batched)"
UNIT = "solves/s"
CONFIG_ID = 2
BATCH_PER_GPU = 1024
WORKLOAD = ("BASELINE configs[1]: batch=1024 per GPU, unicycle quadratic_form, N=50, fixed dt=0.3, 5 circular "
            "obstacles, rate limits 0.2, tol 1e-6, max_iter 100, cold start")


def kkt_bytes_per_instance(cfg):
    """w * ((46 + 4*[rate limits] + 3*[dt free]) * (N-1) + 12) bytes, fp64."""
    rate = any(cfg.du_ub[i] < 1e29 or cfg.du_lb[i] > -1e29 for i in range(2))
    words = 46 + (4 if rate else 0) + (3 if cfg.variable_dt else 0)
    return 8 * (words * (cfg.n - 1) + 12)


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = False
        self.rows = []

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = [("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)]
        while not self.stop_flag:
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            try:
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
            except Exception:
                pw = 0.0
            self.rows.append([str(sm), str(mx), str(pw)] + ["Active" if (r & b) else "Not Active" for _, b in bits])
            time.sleep(0.01)

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                o = subprocess.check_output(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                             "-i", str(self.index)], timeout=5).decode().strip()
                self.rows.append([c.strip() for c in o.split(",")])
            except Exception:
                pass
            time.sleep(0.03)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_arm(cfg, data_fn, seconds_target, threads):
    """The CPU oracle (port) on a bounded sample of the same workload, all host threads."""
    from oracle import oracle_py as orc
    orc.build()
    n = max(threads * 8, 32)
    data = data_fn(n)
    t = time.time()
    out = orc.step_batch(cfg, data, n_threads=threads)
    el = time.time() - t
    # grow the sample towards the time target (bounded)
    if el < seconds_target / 4:
        n2 = int(min(n * (seconds_target / 2) / max(el, 1e-3), 32768))
        data = data_fn(n2)
        t = time.time()
        out = orc.step_batch(cfg, data, n_threads=threads)
        el = time.time() - t
        n = n2
    conv = int((out["status"] == 0).sum())
    return conv / el, n, conv, el


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    from mpc_local_planner_b200 import capi, configs
    cfg = configs.config_for(CONFIG_ID, tol=1e-6)
    B = args.batch
    threads = os.cpu_count() or 1

    # ------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        from oracle import oracle_py as orc
        orc.build()
        sample = max(threads * 32, 256)
        data = configs.generate(CONFIG_ID, sample)
        for _ in range(max(args.warmup, 0)):
            orc.step_batch(cfg, data, n_threads=threads)
        t0 = time.time()
        conv = 0
        for _ in range(args.steps):
            out = orc.step_batch(cfg, data, n_threads=threads)
            conv += int((out["status"] == 0).sum())
        el = time.time() - t0
        val = conv / el
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "reference arm = CPU oracle port of the same algorithm (the reference's "
                       "control_box_rst + Ipopt stack cannot be built here: no ROS/Eigen/Ipopt, see DESIGN.md)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{sample} instances of the workload per step, {args.steps} steps"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return 0

    # ------------------------------------------------------------------------------------------ our arm
    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the solver has no CPU fallback"}))
        return 1
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)

    # weak scaling: rank r solves instances [r*B, (r+1)*B) of the same seeded stream
    data = configs.generate(CONFIG_ID, B, first=rank * B)
    solver = capi.BatchSolver(cfg, B, device=dev)
    N = cfg.n

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # one CUDA stream carries the solver kernels, the L2 flush, the export of u* and the NCCL all-gather, so that
    # CUDA events recorded on it bracket the timed region on the device
    stream = torch.cuda.Stream(device=dev)
    solver.set_stream(stream.cuda_stream)

    # all-gather buffers for u* (SURVEY 8e): every rank ends up with all optimal controls
    send = torch.empty(B * (N - 1) * 2, dtype=torch.float64, device=f"cuda:{dev}")
    recv = torch.empty(world * B * (N - 1) * 2, dtype=torch.float64, device=f"cuda:{dev}") if world > 1 else None

    # The all-gather of step i runs beside the solve of step i+1 (async NCCL work; the work stream waits for it only before the
    # send buffer is written again and at the end of the timed region), so the ranks are not re-synchronised on every step.
    gather = {"work": None}

    def gather_wait():
        if gather["work"] is not None:
            gather["work"].wait()   # device-side: the work stream waits for the collective
            gather["work"] = None

    def gather_controls():
        if world > 1:
            gather_wait()
            solver.export_controls(send.data_ptr())
            gather["work"] = dist.all_gather_into_tensor(recv, send, async_op=True)

    def resident_step():
        solver.flush_l2()  # working set (68 MB) < L2 (126 MB): evict between steps
        solver.solve_resident(cold=True)
        gather_controls()

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize; returns the device time between CUDA events on the work stream
        (the last all-gather included)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                last = fn()
            gather_wait()
            e1.record(stream)
        barrier()
        return e0.elapsed_time(e1) * 1e-3, last

    solver.upload(data["x0"], data["xf"], data["u_prev"], data["u_prev_dt"], data["obstacles"], data["viapoints"])
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            resident_step()
    torch.cuda.synchronize()
    solver.stats_reset()
    sampler = ClockSampler(dev)
    sampler.start()
    el, _ = timed(resident_step, args.steps)
    sampler.stop_flag = True
    st = solver.stats()
    res = solver.fetch()
    conv_local = int((res["status"] == 0).sum())
    iters_mean = float(res["iters"].mean())
    # per-phase device times: one extra, untimed step with every phase bracketed by events (the brackets cost stream time,
    # so the timed region above only brackets the KKT phase)
    solver.set_timing(0x1f)
    solver.stats_reset()
    with torch.cuda.stream(stream):
        resident_step()
    torch.cuda.synchronize()
    st_all = solver.stats()
    solver.set_timing(1 << capi.PHASE_KKT)

    # ---- end-to-end through the C ABI with pinned host buffers ----
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy(), t
    keep = []
    hx0, k = pinned(data["x0"]); keep.append(k)
    hxf, k = pinned(data["xf"]); keep.append(k)
    hup, k = pinned(data["u_prev"]); keep.append(k)
    oc, k = pinned(data["obstacles"][0]); keep.append(k)
    ot, k = pinned(data["obstacles"][1]); keep.append(k)
    op, k = pinned(data["obstacles"][2]); keep.append(k)

    def e2e_step():
        solver.reset()
        solver.flush_l2()
        out = solver.step(hx0, hxf, hup, data["u_prev_dt"], (oc, ot, op), None)  # H2D inputs, solve, D2H results
        gather_controls()
        return out
    with torch.cuda.stream(stream):
        for _ in range(2):
            e2e_step()
    torch.cuda.synchronize()
    st0 = solver.stats()
    e2e_steps = max(3, args.steps // 2)
    el_e2e, out = timed(e2e_step, e2e_steps)
    st1 = solver.stats()
    conv_e2e = int((out["status"] == 0).sum())

    # ---- max over ranks / totals ----
    if dist is not None:
        t = torch.tensor([el, el_e2e], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        el, el_e2e = float(t[0]), float(t[1])
        c = torch.tensor([conv_local, conv_e2e], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        conv_total, conv_e2e_total = int(c[0]), int(c[1])
    else:
        conv_total, conv_e2e_total = conv_local, conv_e2e
    # ---- continuous batching: the K steps' instances as ONE queue through the same 1024-slot pool (mpcb200_solve_stream),
    #      host buffers in, host buffers out.  Reported next to the per-batch numbers, not instead of them; a failure here
    #      must never take the contract line down. ----
    reps = args.steps
    stream_err, el_stream, conv_stream = None, 1.0, 0
    try:
        tile = lambda a: np.ascontiguousarray(np.concatenate([a] * reps))
        q = dict(x0=tile(data["x0"]), xf=tile(data["xf"]), u_prev=tile(data["u_prev"]), obstacles=tuple(tile(a) for a in data["obstacles"]))
        with torch.cuda.stream(stream):  # warm-up with the full queue: the job-sized device arrays are allocated here
            solver.solve_stream(q["x0"], q["xf"], q["u_prev"], data["u_prev_dt"], q["obstacles"], None)
        torch.cuda.synchronize()
        el_stream, sout = timed(lambda: solver.solve_stream(q["x0"], q["xf"], q["u_prev"], data["u_prev_dt"], q["obstacles"], None), 1)
        conv_stream = int((sout["status"] == 0).sum())
        if dist is not None:
            t = torch.tensor([el_stream], dtype=torch.float64, device=f"cuda:{dev}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            c = torch.tensor([conv_stream], dtype=torch.float64, device=f"cuda:{dev}")
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
            el_stream, conv_stream = float(t[0]), int(c[0])
    except Exception as e:
        stream_err = repr(e)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    value = conv_total * args.steps / el
    e2e_value = conv_e2e_total * e2e_steps / el_e2e
    # ---- roofline of the dominant kernel (Riccati KKT) from the CUDA-event times of the timed region ----
    peak, peak_src = hbm_peak()
    bpi = kkt_bytes_per_instance(cfg)
    kkt_ms = st["ms"][capi.PHASE_KKT]
    kkt_launches = max(st["launches"][capi.PHASE_KKT], 1)
    units = st["kkt_instances"]
    achieved = (units * bpi) / (kkt_ms * 1e-3) / 1e9 if kkt_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "kkt_traffic.json")
    if os.path.exists(tp):
        try:
            ratio = json.load(open(tp)).get("traffic_over_algorithmic")
            traffic = int(ratio * units * bpi / kkt_launches) if ratio else None  # DRAM bytes of the average launch (ncu ratio x algorithmic)
        except Exception:
            traffic = None
    total_ms = el / args.steps * 1e3 * args.steps
    roofline = {"bound": "hbm", "kernel": "kkt_kernel (Riccati factorisation + solve)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_instance": bpi, "instances_per_launch_avg": units / kkt_launches,
                "avg_launch_ms": kkt_ms / kkt_launches, "share_of_step": kkt_ms / total_ms if total_ms > 0 else None,
                "sweeps_per_instance": st["kkt_sweeps"] / max(units, 1)}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world, "horizon_n": N,
                   "parallelism": f"instances sharded over {world} GPU(s), NCCL all-gather of u* (step i) beside the solve of step i+1" if world > 1 else "1 GPU",
                   "l2": "flushed between steps (working set 68 MB < 126 MB L2)",
                   "converged_fraction": conv_total / float(B * world), "mean_ipm_iterations": iters_mean},
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": (st1["h2d_bytes"] - st0["h2d_bytes"]) // e2e_steps,
                "d2h_bytes_per_step": (st1["d2h_bytes"] - st0["d2h_bytes"]) // e2e_steps},
        "streaming": {"error": stream_err} if stream_err else {"value": conv_stream / el_stream, "unit": UNIT, "queue_per_gpu": reps * B, "pool_slots_per_gpu": B,
                      "ms_per_1024": el_stream / reps * 1e3,
                      "what": "the same K x 1024 instances per GPU as ONE queue through the 1024-slot pool (mpcb200_solve_stream, "
                              "continuous batching: a finished slot takes the next instance), host buffers in and out; per-instance "
                              "results are bit-identical to the batch solves"},
        "gpu_launches": int(st["launches_total"]),
        "kernel_ms": dict(zip(["init", "associate", "eval", "kkt", "linesearch"], st_all["ms"])),
        "timing": "CUDA events on the work stream around the K steps (max over ranks); kernel_ms from one extra step with all phases bracketed",
        "clocks": sampler.summary(),
    }
    if not args.no_cpu_baseline and world == 1:
        v, n, conv, secs = cpu_arm(cfg, lambda n: configs.generate(CONFIG_ID, n), 20.0, threads)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{n} instances of the same workload ({conv} converged) in {secs:.1f} s, "
                                          "CPU oracle (same algorithm, plain C, one instance per thread)"}
    elif world > 1:
        line["cpu_baseline"] = None
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
