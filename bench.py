#!/usr/bin/env python
"""bench.py — fn evals/sec of a compiled `pytensor.function(..., mode="CUDA")` on B200, with roofline + CPU baseline.

Contract:
  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg4]
  * N=1 workload = BASELINE.json configs[1]: the 32-scalar-op fused Elemwise + CAReduce graph over fp32 (4096,4096).
    A "step" is one evaluation of the compiled function.
  * `value`  : evals/s, inputs resident in HBM, device outputs (CUDA events, K steps, max over ranks).
  * `e2e`    : evals/s through `pytensor.function(..., mode="CUDA")` with pinned HOST inputs and NumPy outputs
               (H2D + D2H inside the timed region).
  * `roofline`: achieved HBM GB/s of the dominant kernel (algorithmic bytes / its CUDA-event duration) over the
               measured peak in MEASURED_PEAKS.json.
  * `cpu_baseline`: the reference's own C linker (mode="CVM") on this box's host cores, same arrays.
  * N>1: every rank evaluates its own (4096,4096) row-shard of an (N*4096,4096) problem — the graph is row-independent,
    so there is no data-path collective ("weak"); the batch-sharded logp+grad graph with its NCCL all-reduce is reported
    beside it under "sharded_logp".
  * `--impl reference` times the reference C linker alone (rank 0 only under torchrun).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402


def _peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "bf16_tflops": d.get("bf16_tflops", 1590.0),
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1400.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def _ncu_traffic(kernel_prefix):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full`
    capture summary of the same command (profiles/*_prof_ew_ncu_summary.csv); None when no capture is committed."""
    import csv
    import glob

    for path in sorted(glob.glob(os.path.join(REPO, "profiles", "*prof_ew*_ncu_summary.csv")), reverse=True):
        try:
            rows = list(csv.reader(open(path)))
            hdr, units = rows[0], rows[1]
            kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            vals = [float(r[rd]) * scale.get(units[rd], 1.0) + float(r[wr]) * scale.get(units[wr], 1.0)
                    for r in rows[2:] if r[kn].startswith(kernel_prefix)]
            if vals:
                return sum(vals) / len(vals)
        except Exception:  # noqa: BLE001
            continue
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self.first_done = threading.Event()  # set when the first nvidia-smi call has returned (or failed)
        self._t = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self.first_done.set()
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = max(float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[3 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(self.samples)}


def pinned_like(arr):
    """NumPy array backed by page-locked host memory (cudaHostAlloc through the C-ABI)."""
    import ctypes

    from pytensor_b200.runtime import lib as L

    p = ctypes.c_void_p()
    L.check(L.lib().ptk_host_alloc_pinned(ctypes.byref(p), arr.nbytes), "pinned alloc")
    buf = (ctypes.c_char * arr.nbytes).from_address(p.value)
    out = np.frombuffer(buf, dtype=arr.dtype).reshape(arr.shape)
    out[...] = arr
    return out


def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path: mode="CVM" on the host cores."""
    if rank != 0:
        return
    from oracle import cvm
    from pytensor_b200 import workloads as W

    pytensor = cvm.configure("float32")
    ins, outs, make_args, meta = W.cfg2_fused_elemwise(args.n)
    f = pytensor.function(ins, outs, mode="CVM", trust_input=True)
    a = make_args()
    for _ in range(max(1, args.warmup)):
        f(*a)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        f(*a)
    dt = time.perf_counter() - t0
    v = args.steps / dt
    env = cvm.describe()
    print(json.dumps({
        "impl": "reference", "metric": "fn evals/sec", "value": v, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"cfg2: 32-op fused Elemwise+CAReduce, fp32 ({args.n},{args.n}), reference C linker (CVM)"},
        "cpu_baseline": {"value": v, "unit": "evals/s", "cores": 1, "kind": "reference",
                         "sample": f"{args.steps} full evaluations of the workload", "env": env},
        "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def _time_dev(f, dev_args, torch, steps, warm):
    for _ in range(warm):
        f(*dev_args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        f(*dev_args)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def extra_workloads(pytensor, W, cuda_mode, dev, torch, peaks):
    """Secondary configs of BASELINE.json (reported under "others"; parity for each is in tests/)."""
    out = {}
    try:  # configs[2]: 3-layer MLP 4096^3, bf16 tcgen05 tensor cores
        ins, outs, make_args, meta = W.cfg3_mlp(4096)
        f = pytensor.function(ins, outs, mode=cuda_mode(device_outputs=True, borrow_outputs=True, gemm_precision="bf16"),
                              trust_input=True)
        a = [dev.to_device(x) for x in make_args()]
        ms = _time_dev(f, a, torch, 10, 4)
        tf = meta["flops"] / (ms * 1e-3) / 1e12
        out["cfg3_mlp_bf16"] = {"evals_per_s": 1e3 / ms, "ms": ms, "tflops": tf, "frac_of_bf16_peak": tf / peaks["bf16_tflops"],
                                "note": "includes fp32->bf16 operand staging and the bias+tanh Elemwise per layer"}
        f32 = pytensor.function(ins, outs, mode=cuda_mode(device_outputs=True, borrow_outputs=True), trust_input=True)
        ms32 = _time_dev(f32, a, torch, 3, 2)
        out["cfg3_mlp_fp32_native"] = {"evals_per_s": 1e3 / ms32, "ms": ms32, "tflops": meta["flops"] / (ms32 * 1e-3) / 1e12}
    except Exception as e:  # noqa: BLE001
        out["cfg3_mlp_bf16"] = {"error": repr(e)[:300]}
    try:  # configs[3]: Scan, 1000 steps, carried state (8192,512) fp32, persistent fused kernel
        ins, outs, make_args, meta = W.cfg4_scan(8192, 512, 1000)
        f = pytensor.function(ins, outs, mode=cuda_mode(device_outputs=True, borrow_outputs=True), trust_input=True)
        a = [dev.to_device(x) for x in make_args()]
        ms = _time_dev(f, a, torch, 5, 3)
        out["cfg4_scan_persistent"] = {"evals_per_s": 1e3 / ms, "ms": ms, "state_bytes_per_step_over_time_GBs":
                                       2 * meta["state_bytes"] * meta["n_steps"] / (ms * 1e-3) / 1e9,
                                       "hbm_algorithmic_GBs": meta["bytes"] / (ms * 1e-3) / 1e9}
        fe = pytensor.function(ins, outs, mode=cuda_mode(device_outputs=True, borrow_outputs=True, fuse=False),
                               trust_input=True)
        mse = _time_dev(fe, a, torch, 2, 3)
        out["cfg4_scan_general_loop"] = {"evals_per_s": 1e3 / mse, "ms": mse}
    except Exception as e:  # noqa: BLE001
        out["cfg4_scan_persistent"] = {"error": repr(e)[:300]}
    try:  # configs[3], full-trace variant: every step's state is an output -> 1000 x 16 MiB of HBM writes
        ins, outs, make_args, meta = W.cfg4_scan(8192, 512, 1000, full_trace=True)
        f = pytensor.function(ins, outs, mode=cuda_mode(device_outputs=True, borrow_outputs=True), trust_input=True)
        a = [dev.to_device(x) for x in make_args()]
        ms = _time_dev(f, a, torch, 4, 3)
        out["cfg4_scan_full_trace"] = {"evals_per_s": 1e3 / ms, "ms": ms, "trace_bytes": meta["bytes"],
                                       "hbm_write_GBs": meta["bytes"] / (ms * 1e-3) / 1e9,
                                       "frac_of_hbm_peak": meta["bytes"] / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
        del f, a
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        out["cfg4_scan_full_trace"] = {"error": repr(e)[:300]}
    try:  # the metric graph of BASELINE.json: 265 compiled nodes (84 x Dot22+tanh(+bias), 16-step Scan, Sum)
        for n, kw, steps in ((64, {}, 50), (1024, {"gemm_precision": "bf16"}, 10)):
            ins, outs, make_args, meta = W.metric_graph(n=n)
            f = pytensor.function(ins, outs, mode=cuda_mode(device_outputs=True, borrow_outputs=True, **kw),
                                  trust_input=True)
            a = [dev.to_device(x) for x in make_args()]
            ms = _time_dev(f, a, torch, steps, 4)
            nn = len(f.maker.fgraph.toposort())
            out[f"metric_graph_n{n}"] = {"evals_per_s": 1e3 / ms, "ms": ms, "compiled_nodes": nn,
                                         "launches_per_eval_after_fusion": len(f.vm.executor.program.steps),
                                         "us_per_compiled_node": 1e3 * ms / nn,
                                         "cuda_graph_replay": bool(f.vm.executor.last_from_graph)}
            if n == 64:
                f_ref = pytensor.function(ins, outs, mode="CVM", trust_input=True)
                from oracle import cvm as _cvm

                evs, k = _cvm.time_function(f_ref, make_args(), min_seconds=2.0, min_calls=5, max_calls=2000)
                out["metric_graph_n64"]["cpu_reference_evals_per_s"] = evs
    except Exception as e:  # noqa: BLE001
        out["metric_graph"] = {"error": repr(e)[:300]}
    return out


def sharded_logp(pytensor, W, cuda_mode, dev, torch, dist, world, rank, steps=10, B_local=1 << 17, collective="nccl"):
    """BASELINE.json configs[4]: hierarchical logp+grad, batch = world x 2^17 independent parameter vectors (2^20 at 8
    GPUs), sharded along the batch axis, ONE packed NCCL all-reduce of [logp, grads] (75 floats) per evaluation."""
    from pytensor_b200.sharded import ShardedSum

    ins, outs, make_args, meta = W.cfg5_logp_grad(B=B_local * world, n=1024, J=64, K=8, dtype="float32", packed=True)
    f = pytensor.function(ins, outs, mode=cuda_mode(device_outputs=True, borrow_outputs=True), trust_input=True)
    local = [dev.to_device(a) for a in make_args(seed=20 + rank, B_local=B_local)]
    sh = ShardedSum(f, batch_arg_idx=[0, 1, 2, 3], collective=collective)
    for _ in range(4):
        res = sh(*local, presharded=True)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        res = sh(*local, presharded=True)
    e1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    logp = float(dev.to_host(res[0]).reshape(-1)[0])
    return {"evals_per_s": steps / (ms * 1e-3), "ms_per_eval": ms / steps, "chains_per_s": B_local * world * steps / (ms * 1e-3),
            "global_batch": B_local * world, "per_gpu_batch": B_local, "n_rows": 1024, "allreduce_floats": 1 + meta["P"],
            "scaling": "weak (2^17 chains per GPU; 2^20 at 8 GPUs)", "nodes": len(f.maker.fgraph.toposort()),
            "logp_sum": logp, "graph_replay": bool(f.vm.executor.last_from_graph), "collective": collective if world > 1 else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sharded", action="store_true", help="skip the batch-sharded logp+grad graph (configs[4])")
    ap.add_argument("--extra", action="store_true", help="also time cfg3/cfg4 and report them under 'others'")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from oracle import cvm  # configures PYTENSOR_FLAGS for the host framework (compile dir, BLAS)

    pytensor = cvm.configure("float32")
    import pytensor_b200  # noqa: F401
    from pytensor_b200 import workloads as W
    from pytensor_b200.link.cuda import cuda_mode
    from pytensor_b200.runtime import device as dev
    from pytensor_b200.runtime import jit

    dev.device()
    ins, outs, make_args, meta = W.cfg2_fused_elemwise(args.n)
    f_dev = pytensor.function(ins, outs, mode=cuda_mode(device_outputs=True, borrow_outputs=True), trust_input=True)
    f_host = pytensor.function(ins, outs, mode="CUDA", trust_input=True)
    host_args = [make_args(1 + 10 * rank), make_args(101 + 10 * rank)]
    dev_args = [[dev.to_device(a) for a in s] for s in host_args]  # two input sets: 2 x 128 MiB > the 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM ------------------------------------------------------------------------
    # (first call of a signature runs eagerly and measures, the second captures the CUDA graph, later calls replay it;
    #  two alternating input sets = two signatures, so at least 6 warm-up calls)
    warm = max(args.warmup, 6)
    for i in range(warm):
        f_dev(*dev_args[i % 2])
    barrier()
    l0 = jit.stats["launches"]
    ex = f_dev.vm.executor
    n_kernels = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        # Spawning nvidia-smi forks this (large) process while holding the GIL for milliseconds: if that lands inside the
        # ~1 ms timed region the launch loop stalls and the device idles.  So keep running the same (untimed) work until
        # the sampler's first call has returned — that sample sees this workload's clocks — and time the K steps in the
        # 200 ms gap before its next call.
        t_wait = time.perf_counter() + 3.0
        while not clocks.first_done.is_set() and time.perf_counter() < t_wait:
            for i in range(16):
                f_dev(*dev_args[i % 2])
            torch.cuda.synchronize()
        barrier()
        e0.record()
        for i in range(args.steps):
            f_dev(*dev_args[i % 2])
        e1.record()
        barrier()
        # the timed region lasts ~1 ms: keep the SAME work running for ~1.2 s more (untimed) so that nvidia-smi (200 ms
        # period) sees the clocks / throttle reasons of this workload under sustained load
        t_end = time.perf_counter() + 1.2
        while time.perf_counter() < t_end:
            for i in range(64):
                f_dev(*dev_args[i % 2])
            torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    replayed = bool(ex.last_from_graph)
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * args.steps / (ms / 1e3)

    # ---- per-node device time: eager pass behind a queue filler so that host launch gaps do not pollute the events ---
    f_prof = pytensor.function(ins, outs, mode=cuda_mode(device_outputs=True, use_graph=False), trust_input=True)
    for i in range(3):
        f_prof(*dev_args[i % 2])
    torch.cuda.synchronize()
    exp = f_prof.vm.executor
    filler_src = dev.empty((1 << 28,), "float32")  # 1 GiB
    filler_dst = dev.empty((1 << 28,), "float32")
    for _ in range(6):
        dev.copy_strided(filler_dst, filler_src)   # ~0.35 ms each of queued device work
    l_before = jit.stats["launches"]
    exp.event_log = []
    prof_steps = 12
    for i in range(prof_steps):
        f_prof(*dev_args[i % 2])
    torch.cuda.synchronize()
    launches_per_step = (jit.stats["launches"] - l_before) // prof_steps
    log, exp.event_log = exp.event_log, None
    del filler_src, filler_dst
    per_step = {}
    for i, a, b in log:
        per_step.setdefault(i, []).append(a.elapsed_time(b))
    step_ms = {i: float(np.median(v)) for i, v in per_step.items()}
    dom = max(step_ms, key=step_ms.get)
    dom_name = repr(exp.program.steps[dom].impl)
    launches = launches_per_step * args.steps

    # ---- e2e: host buffers through the public API -------------------------------------------------------------------
    pin_args = [[pinned_like(a) for a in s] for s in host_args]
    for i in range(max(4, args.warmup)):
        res = f_host(*pin_args[i % 2])  # (keeps the previous result alive like the timed loop: warms the pinned pool)
    barrier()
    e2e_steps = max(3, min(args.steps, 20))
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        res = f_host(*pin_args[i % 2])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e = world * e2e_steps / dt
    h2d = int(sum(a.nbytes for a in host_args[0]))
    d2h = int(sum(np.asarray(r).nbytes for r in res))

    # ---- roofline of the dominant kernel ---------------------------------------------------------------------------
    peaks = _peaks()
    fused = "fused" in dom_name
    alg_bytes = meta["bytes"] if fused else 3 * 4 * args.n * args.n
    achieved = alg_bytes / (step_ms[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": _ncu_traffic("ptk_ew_red_row" if fused else "ptk_ew_vec"),
                "kernel": dom_name,
                "kernel_ms": step_ms[dom], "algorithmic_bytes": alg_bytes, "peak_source": peaks["source"],
                "how": "median CUDA-event duration of the node's launch over 12 eager evaluations queued behind device "
                       "work (no host gaps); whole_graph = all bytes / graph-replayed step time of the timed region",
                "step_ms_by_node": {repr(exp.program.steps[i].impl): v for i, v in step_ms.items()},
                "whole_graph": {"bytes": meta["bytes"], "gbs": meta["bytes"] * args.steps / (ms * 1e-3) / 1e9,
                                "frac": meta["bytes"] * args.steps / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                "cuda_graph_replay": replayed}}

    # ---- CPU baseline: the reference C linker on the host cores (rank 0, N=1 only) -------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        f_ref = pytensor.function(ins, outs, mode="CVM", trust_input=True)
        evs, n = cvm.time_function(f_ref, host_args[0], min_seconds=8.0, min_calls=3, max_calls=50)
        cpu = {"value": evs, "unit": "evals/s", "cores": 1, "kind": "reference",
               "sample": f"{n} full evaluations of the same workload (Elemwise/CAReduce C loops are single-threaded: "
                         "config.openmp=False)", "env": cvm.describe()}

    others = None
    if args.extra and rank == 0 and world == 1:
        others = extra_workloads(pytensor, W, cuda_mode, dev, torch, peaks)
    sharded = None
    if not args.no_sharded:
        try:
            sharded = sharded_logp(pytensor, W, cuda_mode, dev, torch, dist, world, rank)
        except Exception as e:  # noqa: BLE001
            sharded = {"error": repr(e)[:400]}
        if world > 1 and "error" not in sharded:
            try:  # the same evaluation with the hand-written one-shot NVLink all-reduce instead of NCCL
                alt = sharded_logp(pytensor, W, cuda_mode, dev, torch, dist, world, rank, collective="oneshot")
                sharded["oneshot_nvlink"] = {k: alt[k] for k in ("evals_per_s", "ms_per_eval", "logp_sum")}
            except Exception as e:  # noqa: BLE001
                sharded["oneshot_nvlink"] = {"error": repr(e)[:300]}

    line = {
        "metric": "fn evals/sec", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"cfg2 (BASELINE.json configs[1]): 32-scalar-op fused Elemwise+CAReduce graph, fp32 "
                               f"({args.n},{args.n}) -> e ({args.n},{args.n}) f32, r=e.sum(1) f32 (acc f64)",
                   "l2": "two alternating input sets, 384 MiB working set > 126 MB L2", "parallelism":
                   f"{world} independent row shards (no collective)"},
        "e2e": {"value": e2e, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps,
                "pipeline": ("chunked H2D|kernel|D2H x%d" % len(f_host.vm.executor._chunk_plan["bounds"]))
                if f_host.vm.executor.chunked_calls else "one-shot"},
        "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks.summary(),
    }
    if others is not None:
        line["others"] = others
    if sharded is not None:
        line["sharded_logp"] = sharded
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
