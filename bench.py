#!/usr/bin/env python
"""bench.py -- forward+likelihood evaluations per second on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload target|c2|c3|c3_buried|c4|c5|sample] [--chains C]
    python bench.py --impl reference ...      # the CPU restatement of the reference algorithm on the host cores

A step = one pass of the hot path (format_model -> propagator -> filter/FFT -> misfit -> correlated-noise
likelihood) over one batch of `--chains` models per GPU.  `value` is measured with the models resident in HBM
(rfinv_eval_batch_device); `e2e` goes through rfinv_eval_batch with host buffers, H2D/D2H copies in the timed region.
One process per GPU (torchrun); chains shard across ranks with no data-path collective (weak scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from rf_inv_b200 import workloads  # noqa: E402

METRIC = "forward+likelihood evals/sec"
UNIT = "evals/s"
DEFAULT_CHAINS = {"target": 16384, "c2": 4096, "c3": 8192, "c3_buried": 8192, "c4": 8192, "c5": 4096, "sample": 4096}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="target")
    ap.add_argument("--chains", type=int, default=0, help="models per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=0, help="models in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pt-iters", type=int, default=20, help="timed PT-MCMC iterations (0 = skip the PT leg)")
    return ap.parse_args()


def workload_desc(cfg, name, chains, k_mean):
    return {"workload": name, "nfft": cfg.nfft, "k_max": cfg.k_max, "ntrc": cfg.ntrc, "nsmp": cfg.nsmp,
            "rays": "common" if cfg.is_ray_common else "distinct", "sea_layer": cfg.sdep > 0,
            "chains_per_gpu": chains, "k_mean": round(float(k_mean), 2),
            "models": "k uniform on [k_min,k_max), interfaces uniform, dVs ~ N(0,0.3): as init_model draws them",
            "l2": "flushed between timed steps (256 MiB write)"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, reasons, mx = [], set(), None
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        hi = [s for s in sm if s > 0.5 * (max(sm) if sm else 1)]
        return {"sm_mhz": float(np.median(hi)) if hi else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(cfg, chains, seed):
    return workloads.draw_models(cfg, chains, seed=seed, dvs_scale=0.3)


def attach_obs_via_cuda(cfg, Evaluator):
    """Observed traces = CUDA forward of the data-generating model + noise, rounded to float32 like a SAC file."""
    tm = workloads.true_model(cfg)
    cfg.obs = np.zeros((cfg.ntrc, cfg.nsmp))
    cfg.r_inv = workloads.lapack_r_inv(cfg)
    with Evaluator(cfg) as ev:
        _, rft, _ = ev.calc_likelihood(tm["k"], tm["z"], tm["dvp"], tm["dvs"], tm["sig"], want_rft=True)
    obs = rft[0, :, :cfg.nsmp] + np.random.default_rng(7).normal(0.0, 0.01, (cfg.ntrc, cfg.nsmp))
    cfg.obs = obs.astype(np.float32).astype(np.float64)
    return cfg


def host_cores():
    """Host threads the CPU arm uses: every core this process may run on (torchrun exports OMP_NUM_THREADS=1 to its
    workers, which must not throttle the CPU baseline)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_sample_size(cfg, models, target_s, cap):
    """Number of models that keep one CPU pass near `target_s` seconds (calibrated on a 64-model probe)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_c
    cores = host_cores()
    probe = min(max(4 * cores, 64), cap)
    sub = {k: v[:probe] for k, v in models.items()}
    oracle_c.eval_batch(cfg, sub["k"], sub["z"], sub["dvp"], sub["dvs"], sub["sig"], want_rft=False, nthreads=cores)
    t0 = time.perf_counter()
    oracle_c.eval_batch(cfg, sub["k"], sub["z"], sub["dvp"], sub["dvs"], sub["sig"], want_rft=False, nthreads=cores)
    rate = probe / max(time.perf_counter() - t0, 1e-6)
    return int(min(cap, max(probe, rate * target_s)))


def cpu_arm(cfg, models, n_sample, steps, warmup):
    """Times the C restatement of the reference algorithm (oracle/) on all host cores: evals/s."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_c
    cores = host_cores()
    sub = {k: v[:n_sample] for k, v in models.items()}
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        oracle_c.eval_batch(cfg, sub["k"], sub["z"], sub["dvp"], sub["dvs"], sub["sig"], want_rft=False, nthreads=cores)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return n_sample / float(np.mean(times)), cores, float(np.mean(times))


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (C restatement; the Fortran/MPI/FFTW/LAPACK original cannot
    be compiled in this image) on the box's host cores, same config/metric.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = workloads.make_config(args.workload)
    chains = args.chains or DEFAULT_CHAINS.get(args.workload, 4096)
    sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    cfg = helpers.attach_obs_and_rinv(cfg, noise=0.01)
    import oracle_c
    cores = host_cores()
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    models = make_inputs(cfg, min(chains, 16384), seed=100)
    # bounded sample: about 20 s of CPU work over the whole --steps/--warmup run
    n_sample = args.cpu_sample or cpu_sample_size(cfg, models, 20.0 / (steps + warmup), len(models["k"]))
    models = {k: v[:n_sample] for k, v in models.items()}
    k_mean = float(np.mean(models["k"]))
    val, cores, sec = cpu_arm(cfg, models, n_sample, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_desc(cfg, args.workload, chains, k_mean),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n_sample} models of the workload per step, OpenMP over models"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    from rf_inv_b200 import capi
    from rf_inv_b200.evaluator import Evaluator
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; rf_inv_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL announces its version on stdout when the communicator is created (NCCL_DEBUG=VERSION in this image); rank 0's
        # stdout must carry the one JSON line and nothing else, so stdout points at stderr until the communicator exists
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    dev = torch.device("cuda", local_rank)

    cfg = workloads.make_config(args.workload)
    chains = args.chains or DEFAULT_CHAINS.get(args.workload, 4096)
    cfg = attach_obs_via_cuda(cfg, lambda c: Evaluator(c, device=local_rank))
    models = make_inputs(cfg, chains, seed=100 + rank)
    k_mean = float(np.mean(models["k"]))
    soa = workloads.to_soa(models)
    d = {k: torch.from_numpy(v).to(dev) for k, v in soa.items()}
    logl = torch.empty(chains, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # pinned host copies for the end-to-end arm
    pin = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in models.items()}
    pin_np = {k: v.numpy() for k, v in pin.items()}
    # bytes rfinv_eval_batch copies to the device per call: dVp stays on the host when vp_mode = 0 (format_model ignores it)
    h2d = sum(v.numel() * v.element_size() for k, v in pin.items() if not (k == "dvp" and cfg.vp_mode == 0))
    d2h = 8 * chains

    ev = Evaluator(cfg, device=local_rank)
    stream = torch.cuda.current_stream()
    ev.set_stream(stream.cuda_stream)
    lib = capi.load()
    capi.check(lib.rfinv_set_timing(ev.handle, 1))

    def step_device():
        ev.calc_likelihood_device(chains, d["k"].data_ptr(), d["z"].data_ptr(), d["dvp"].data_ptr(), d["dvs"].data_ptr(),
                                  d["sig"].data_ptr(), logl.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm (value) ----
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    step_ms, kern_ms = [], []
    launches = 0
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)                      # evict L2 between timed steps (outside the per-step events)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step_device()
        e1.record(stream)
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        t3 = (C.c_double * 3)()
        capi.check(lib.rfinv_get_timing(ev.handle, t3))
        kern_ms.append(list(t3))
        launches += ev.last_launch_count
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    total_ms = float(np.sum(step_ms))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * chains * args.steps / (total_ms * 1e-3)

    # ---- end-to-end arm: host buffers through the C-ABI call a Fortran host would make ----
    for _ in range(2):
        ev.calc_likelihood(pin_np["k"], pin_np["z"], pin_np["dvp"], pin_np["dvs"], pin_np["sig"])
    barrier()
    e2e_t = []
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ll_host, _, _ = ev.calc_likelihood(pin_np["k"], pin_np["z"], pin_np["dvp"], pin_np["dvs"], pin_np["sig"])
        e2e_t.append(time.perf_counter() - t0)
    e2e_total = float(np.sum(e2e_t))
    if world > 1:
        t = torch.tensor([e2e_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_total = float(t.item())
    e2e_value = world * chains * args.steps / e2e_total
    assert np.array_equal(ll_host, logl.cpu().numpy()), "host and device entry points disagree"

    # ---- PT-MCMC leg: iterations/s of pt_control with every chain on device (second half of BASELINE.json's metric) ----
    pt_info = None
    if args.pt_iters > 0:
        from rf_inv_b200.pt import ParallelTempering
        nch = cfg.nchains
        nproc_total = world * max(1, chains // nch)
        pt = ParallelTempering(cfg, nproc_total, device=local_rank, world=world, rank=rank)
        pt.ev.set_stream(stream.cuda_stream)

        def pt_iters(n):
            if world == 1:
                pt.run(n)
            else:
                pt.run_distributed(n, dist, torch)

        pt_iters(3)
        barrier()
        n_eval0 = pt.counters()["n_eval"]
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        pt_iters(args.pt_iters)
        e1.record(stream)
        e1.synchronize()
        pt_ms = e0.elapsed_time(e1)
        n_eval = pt.counters()["n_eval"] - n_eval0
        k_pt = float(np.mean(pt.state()["k"]))
        tt = torch.tensor([pt_ms, float(n_eval)], dtype=torch.float64, device=dev)
        if world > 1:
            mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            pt_ms, n_eval = float(mx[0].item()), float(sm[1].item())
        pt_info = {"iters_per_s": args.pt_iters / (pt_ms * 1e-3), "chain_steps_per_s": world * pt.n_local * args.pt_iters / (pt_ms * 1e-3),
                   "forward_evals_per_s": n_eval / (pt_ms * 1e-3), "iters": args.pt_iters, "chains_total": world * pt.n_local,
                   "virtual_ranks": nproc_total, "chains_per_rank": nch, "k_mean_after": round(k_pt, 2),
                   "exchange": "none (single process)" if world == 1 else "one NCCL all-gather of (T, logL, next-uniform) tables per iteration"}
        pt.close()

    if rank == 0:
        # ---- roofline of the dominant kernel (forward_kernel): algorithmic fp64 flop / CUDA-event time ----
        dfma, dmma = C.c_double(0), C.c_double(0)
        capi.check(lib.rfinv_measure_fp64_peak(local_rank, C.byref(dfma), C.byref(dmma)))
        peak = max(dfma.value, dmma.value)
        fl = workloads.flops_per_eval(cfg, k_mean)
        # the quadratic form as the library evaluates it (chosen at rfinv_create): S^2 + 3S dense, 2 S r + 2 r factor form,
        # S r + S + 2 r split form (sums / differences of mirrored samples against the two half-length factors)
        form = ev.quadform_form()
        S = cfg.nsmp
        fl["quadform_dense_form"] = fl["quadform"]
        fl["quadform"] = float(sum((S * r + S + 2 * r) if sp else ((2 * S * r + 2 * r) if r > 0 else (S * S + 3 * S)) for r, rs, sp in form))
        fl["quadform_form"] = [{"rank": r, "rank_sym": rs, "split": bool(sp)} for r, rs, sp in form]
        fl["total"] = fl["propagator"] + fl["fft"] + fl["quadform"]
        km = np.mean(np.array(kern_ms), axis=0)
        fwd_flop = (fl["propagator"] + fl["fft"]) * chains
        achieved = fwd_flop / (km[0] * 1e-3) * 1e-12
        qf_achieved = fl["quadform"] * chains / (km[1] * 1e-3) * 1e-12
        whole = fl["total"] * chains * args.steps / (float(np.sum(step_ms)) * 1e-3) * 1e-12
        # share of the FP64 pipe's issue slots the forward path fills (an FP64 instruction holds the pipe for the time of
        # one FMA whether or not it is one): instructions x 2 flop-slots / time / peak
        slot_frac = fl["forward_fp64_instructions"] * 2.0 * chains / (km[0] * 1e-3) * 1e-12 / peak
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "forward_kernel_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("workload") == args.workload and tj.get("chains") == chains:
                traffic = tj.get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_desc(cfg, args.workload, chains, k_mean),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "fp64", "kernel": "forward_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": traffic, "fp64_issue_slot_frac": slot_frac,
                         "note": "FP64-pipe bound (arithmetic intensity > 100 flop/B, DRAM < 2% of peak): `achieved` = algorithmic "
                                 "flop of prep_kernel + forward_kernel / their CUDA-event time; `traffic` = ncu dram bytes of one "
                                 "forward_kernel launch at this shape (profiles/forward_kernel_traffic.json)",
                         "peak_source": "measured live: rfinv_measure_fp64_peak (max of DFMA %.1f / DMMA %.1f TFLOP/s); "
                                        "MEASURED_PEAKS.json has no FP64 figure" % (dfma.value, dmma.value),
                         "flop_per_eval": fl, "kernel_ms": {"forward": km[0], "quadform": km[1], "loglik": km[2]},
                         "quadform_tflops": qf_achieved, "whole_step_tflops": whole, "whole_step_frac": whole / peak},
            "wall_s_timed_region": t_wall,
            "pt": pt_info,
        }
        if world == 1 and not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            cores = host_cores()
            n_sample = args.cpu_sample or cpu_sample_size(cfg, models, 5.0, chains)   # ~15 s of CPU work in 3 passes
            val, cores, sec = cpu_arm(cfg, models, min(n_sample, chains), 2, 1)
            line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {min(n_sample, chains)} models of the same batch, {sec:.2f} s per pass, "
                                              "C restatement of the reference algorithm (OpenMP over models)"}
        print(json.dumps(line), flush=True)
    ev.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
