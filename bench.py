#!/usr/bin/env python
"""bench.py -- NPDR hot-path benchmark (BASELINE.json metric: attr x neighbor-pair GLM evals/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4|cfg5|small] [--impl ours|reference]

One "step" = one full pass of the hot path (distances -> multiSURF neighbour pairs -> per-attribute binomial
IRLS) over one synthetic batch.  Workloads (BASELINE.json configs; synthetic stand-ins, seed 1618):
  cfg4 (default)  4,000 instances x 20,000 attributes per GPU, binomial, multisurf, manhattan, q = 0.
                  With --gpus N the attribute count grows to 20,000 x N (weak scaling: per-GPU work fixed).
  cfg5            20,000 x 50,000, binomial, multisurf, manhattan, 2 covariates (strong scaling: total fixed).
  small           1,024 x 2,048 (development).
Sharding (N > 1, one process per GPU under torchrun): row blocks for distance+neighbours (symmetric halves of the
distance matrix swapped in one NCCL all-to-all), one all-gather of the pair segments, attribute columns for the
regression, statistics gathered to rank 0.

`value` is p*M / step time with inputs resident in HBM; `e2e` is the same metric through the host-buffer API
(H2D of X, y inside the timed region, D2H of the statistics).  `roofline` describes the dominant kernel
(the binomial IRLS pass kernel, FP64-issue bound) with CUDA-event launch durations; `cpu_baseline` /
`--impl reference` time the CPU oracle (a restatement of the reference R algorithm; R itself cannot be installed).
"""
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

WORKLOADS = {
    "cfg4": dict(N=4000, p_per_gpu=20000, p_total=None, q=0, pct_signals=0.10, scaling="weak",
                 desc="synthetic createSimulation2-style interaction data 4,000 x 20,000 per GPU: binomial, multisurf, manhattan"),
    "cfg5": dict(N=20000, p_per_gpu=None, p_total=50000, q=2, pct_signals=0.01, scaling="strong",
                 desc="synthetic 20,000 x 50,000: binomial, multisurf, manhattan, 2 covariates"),
    "small": dict(N=1024, p_per_gpu=2048, p_total=None, q=0, pct_signals=0.05, scaling="weak",
                  desc="development size 1,024 x 2,048 per GPU"),
}
SEED = 1618      # the package's convention (data-raw/fetch-data.R:10)


def f_bin(q):
    """Canonical FP64 flop count per attribute x pair x IRLS iteration: 68 for q=0, 92 for q=2."""
    k = 2 + q
    return 2 * ((k - 1) + k * (k + 1) // 2 + k) + (k - 1) + 7 + 48


def fp64_instr(q):
    """FP64 instructions the v2 IRLS pass kernel EXECUTES per attribute x record evaluation (SASS count):
    11 (diff 1, range reduction 3, degree-2 exp polynomial 2, 1+t 1, reciprocal 2, weight 1, deviance product 1) +
    2(k-1) (eta, w*x) + k(k+1)/2 (X'WX) + k (score), k = 2 + q: 18 at q = 0, 31 at q = 2."""
    k = 2 + q
    return 11 + 2 * (k - 1) + k * (k + 1) // 2 + k


def gen_columns(gen, N, ncols, col0, p_total, n_func, f, g, sgn, device, dtype):
    """Columns [col0, col0+ncols) of the stand-in data: x = sqrt(.8 - l^2) z + sqrt(.2) f + l * s * g  (l = 0 for
    background attributes; for the n_func functional attributes the sign s flips between cases and controls on
    half of them -> differential correlation, no main effect; createSimulation2 interactionErdos in spirit)."""
    import torch
    z = torch.randn((ncols, N), generator=gen, device=device, dtype=dtype)
    idx = torch.arange(col0, col0 + ncols, device=device)
    lam = torch.where(idx < n_func, 0.6, 0.0).to(dtype).unsqueeze(1)
    flip = (idx >= n_func // 2) & (idx < n_func)
    s = torch.where(flip.unsqueeze(1), sgn.unsqueeze(0), torch.ones_like(sgn).unsqueeze(0))
    return torch.sqrt(0.8 - lam * lam) * z + np.sqrt(0.2) * f.unsqueeze(0) + lam * s * g.unsqueeze(0)


def make_data(wl, world, device):
    """-> dict with device tensors dX [p, ldx] (column-major N x p, ldx = N rounded up to 16), dy, dC."""
    import torch
    N, q = wl["N"], wl["q"]
    p = wl["p_total"] or wl["p_per_gpu"] * world
    ldx = (N + 15) // 16 * 16
    gen = torch.Generator(device=device)
    gen.manual_seed(SEED)
    dt = torch.float64
    perm = torch.randperm(N, generator=gen, device=device)
    y = (perm < N // 2).to(dt)                                  # balanced classes
    f = torch.randn(N, generator=gen, device=device, dtype=dt)
    g = torch.randn(N, generator=gen, device=device, dtype=dt)
    sgn = torch.where(y > 0, 1.0, -1.0).to(dt)
    n_func = int(round(wl["pct_signals"] * p))
    dX = torch.zeros((p, ldx), device=device, dtype=dt)
    step = 2048
    for c0 in range(0, p, step):
        nc = min(step, p - c0)
        dX[c0:c0 + nc, :N] = gen_columns(gen, N, nc, c0, p, n_func, f, g, sgn, device, dt)
    dC = None
    if q:
        sex = (torch.rand(N, generator=gen, device=device, dtype=dt) < 0.5).to(dt)
        age = torch.round(40.0 + 12.0 * torch.randn(N, generator=gen, device=device, dtype=dt))
        dC = torch.stack([sex, age] + [torch.randn(N, generator=gen, device=device, dtype=dt) for _ in range(q - 2)])[:q].contiguous()
    return dict(N=N, p=p, q=q, ldx=ldx, dX=dX, dy=y.contiguous(), dC=dC, n_func=n_func,
                covar_diff_type=("match-mismatch", "numeric-abs")[:q])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_oracle_sample(wl_name, N_s, p_s, q):
    """One pass of the hot path through the CPU oracle on a bounded sample (first N_s instances x p_s attributes of
    the same synthetic recipe, numpy generator).  -> (evals, seconds, threads, M)."""
    from oracle import oracle as O
    O.build()
    rng = np.random.default_rng(SEED)
    f = rng.normal(size=(N_s, 1)); g = rng.normal(size=(N_s, 1))
    y = (rng.permutation(N_s) < N_s // 2).astype(float)
    sgn = np.where(y > 0, 1.0, -1.0)[:, None]
    n_func = max(2, int(round(WORKLOADS[wl_name]["pct_signals"] * p_s)))
    lam = np.where(np.arange(p_s) < n_func, 0.6, 0.0)[None, :]
    flip = (np.arange(p_s) >= n_func // 2) & (np.arange(p_s) < n_func)
    s = np.where(flip[None, :], sgn, 1.0)
    X = np.asfortranarray(np.sqrt(0.8 - lam * lam) * rng.normal(size=(N_s, p_s)) + np.sqrt(0.2) * f + lam * s * g)
    C = None
    if q:
        C = np.column_stack([(rng.random(N_s) < 0.5).astype(float), np.round(40 + 12 * rng.normal(size=N_s))])[:, :q]
    t0 = time.perf_counter()
    ri, nn, D, r = O.nearest_neighbors(X, "multisurf", "manhattan", 0.5)
    yd = O.pair_diffs(y, ri, nn, "match-mismatch")
    cd = None
    if q:
        types = ("match-mismatch", "numeric-abs")
        cd = np.column_stack([O.pair_diffs(C[:, c], ri, nn, types[c]) for c in range(q)])
    O.regress_attrs(X, ri, nn, yd, cd, "numeric-abs", "binomial")
    dt = time.perf_counter() - t0
    return float(p_s) * ri.size, dt, O.num_threads(), int(ri.size)


def run_reference(args, wl_name, wl):
    """--impl reference: the CPU restatement of the reference path on the host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N_s, p_s = (1000, 256) if wl_name != "small" else (512, 128)
    vals, secs, M = [], [], 0
    for it in range(args.warmup + args.steps):
        ev, dt, thr, M = cpu_oracle_sample(wl_name, N_s, p_s, wl["q"])
        if it >= args.warmup:
            vals.append(ev / dt); secs.append(dt)
    value = float(np.sum([v * s for v, s in zip(vals, secs)]) / np.sum(secs))
    sample = (f"first {N_s} instances x {p_s} attributes of the {wl_name} recipe (M={M} pairs), all stages, "
              f"OpenMP over attributes (the reference's dopar.reg analogue)")
    line = {"impl": "reference", "metric": "npdr_attr_x_pair_glm_evals_per_s", "value": value, "unit": "evals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)),
            "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl_name, "description": wl["desc"], "regression": "binomial", "nbd": "multisurf",
                       "metric": "manhattan", "covariates": wl["q"]},
            "cpu_baseline": {"value": value, "unit": "evals/s", "cores": thr, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "CPU oracle = C restatement of the reference R algorithm (R cannot be installed here; the reference "
                    "itself cannot allocate its dense M x p matrix at this workload, R/npdr.R:313)"}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("NPDRB_BENCH_WORKLOAD", "cfg4"), choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    args = ap.parse_args()
    wl_name, wl = args.workload, WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl_name, wl)
        return

    import torch
    import torch.distributed as dist
    import npdr_b200 as nb
    from npdr_b200.distributed import GpuEngine, npdr_sharded

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (npdr_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    ctx = nb.context(local_rank)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)      # library work and torch events share one stream

    data = make_data(wl, world, device)
    N, p, q = data["N"], data["p"], data["q"]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step_resident():
        eng = GpuEngine(device_index=local_rank, dX=data["dX"], ldx=data["ldx"], dy=data["dy"], dC=data["dC"], N=N, p=p, q=q)
        stats, info = npdr_sharded(eng, N, p, "binomial", "numeric-abs", "multisurf", "manhattan", 0, 0.5, data["covar_diff_type"])
        km = eng.session.kernel_ms(); sm = eng.session.stage_ms(); passes = eng.session.irls_passes
        eng.close()
        return stats, info, km, sm, passes

    # host copies for the end-to-end arm (pinned).  One GPU: the whole matrix goes through the host-buffer C ABI.  Several
    # GPUs: every rank holds the pinned host copy of ITS attribute columns, uploads them over its own PCIe link and the
    # shards are all-gathered over NVLink (a single host thread would feed the G devices the same way); y / covariates
    # come from rank 0.  When p does not divide evenly rank 0 uploads everything and broadcasts.
    hX = hy = hC = hXt = None
    e2e_dev = None
    shard_upload = world > 1 and p % world == 0
    if not args.no_e2e and (rank == 0 or shard_upload):
        a0, na = (0, p) if not shard_upload else (rank * (p // world), p // world)
        hXt = torch.empty((na, N), dtype=torch.float64, pin_memory=True)
        hXt.copy_(data["dX"][a0:a0 + na, :N])
        if world == 1:
            hX = hXt.numpy().T                                   # (N, p) Fortran-ordered view of pinned memory
        if rank == 0:
            hy = data["dy"].cpu().numpy()
            hC = np.asfortranarray(data["dC"].cpu().numpy().T) if q else None
    if not args.no_e2e and world > 1:
        e2e_dev = dict(dX=torch.zeros((p, data["ldx"]), dtype=torch.float64, device=device),
                       dy=torch.zeros(N, dtype=torch.float64, device=device),
                       dC=torch.zeros((q, N), dtype=torch.float64, device=device) if q else None)
        hy_t = torch.from_numpy(hy).pin_memory() if rank == 0 else None
        hC_t = torch.from_numpy(np.ascontiguousarray(hC.T)).pin_memory() if (rank == 0 and q) else None

    def one_step_e2e():
        if world == 1:
            eng = GpuEngine(X=hX, y=hy, C=hC, device_index=local_rank)   # host-buffer C ABI: H2D upload inside
        else:
            if shard_upload:                                     # own columns over the own PCIe link, then NVLink all-gather
                na = p // world
                mine = e2e_dev["dX"][rank * na:(rank + 1) * na]
                mine[:, :N].copy_(hXt, non_blocking=True)
                dist.all_gather_into_tensor(e2e_dev["dX"], mine)
            elif rank == 0:
                e2e_dev["dX"][:, :N].copy_(hXt, non_blocking=True)
            if rank == 0:
                e2e_dev["dy"].copy_(hy_t, non_blocking=True)
                if q:
                    e2e_dev["dC"].copy_(hC_t, non_blocking=True)
            if not shard_upload:
                dist.broadcast(e2e_dev["dX"], 0)
            dist.broadcast(e2e_dev["dy"], 0)
            if q:
                dist.broadcast(e2e_dev["dC"], 0)
            eng = GpuEngine(device_index=local_rank, dX=e2e_dev["dX"], ldx=data["ldx"], dy=e2e_dev["dy"], dC=e2e_dev["dC"],
                            N=N, p=p, q=q)
        stats, info = npdr_sharded(eng, N, p, "binomial", "numeric-abs", "multisurf", "manhattan", 0, 0.5, data["covar_diff_type"])
        eng.close()
        return stats, info                                              # stats are host numpy on rank 0: D2H done

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        l0 = ctx.launch_count()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        launches = torch.tensor([ctx.launch_count() - l0], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(launches, op=dist.ReduceOp.SUM)
        return float(ms.item()), int(launches.item()), out

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_total, launches, (stats, info, km, sm, passes) = timed(one_step_resident, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    M = info["n_pairs"]
    ms_step = ms_total / args.steps
    value = float(p) * M / (ms_step * 1e-3)

    # the same step with the two exact work-saving devices switched off: every reference pair evaluated
    # (no mutual-pair folding) and every IRLS iteration confirmed by its own pass (no guarded speculative convergence)
    variants = {}
    if not args.no_variants:
        for name, env in (("no_speculation", {"NPDRB_NO_SPECULATE": "1"}),
                          ("no_speculation_no_folding", {"NPDRB_NO_SPECULATE": "1", "NPDRB_NO_FOLD": "1"})):
            os.environ.update(env)
            try:
                ms_v, _, (_, info_v, km_v, _, passes_v) = timed(one_step_resident, 2, 1)
            finally:
                for k in env:
                    os.environ.pop(k, None)
            variants[name] = {"value": float(p) * M / (ms_v / 2 * 1e-3), "ms_per_step": ms_v / 2, "irls_full_passes": passes_v,
                              "records_per_pair": km_v["records"] / max(info_v["n_pairs_used"], 1)}

    e2e = None
    if not args.no_e2e:
        e_steps = max(1, min(args.steps, 3))
        ms_e, _, _ = timed(one_step_e2e, e_steps, 1)
        ms_e /= e_steps
        e2e = {"value": float(p) * M / (ms_e * 1e-3), "unit": "evals/s", "ms_per_step": ms_e,
               "h2d_bytes_per_step": int(N * p * 8 + N * 8 + N * q * 8), "d2h_bytes_per_step": int(p * 8 * 8),
               "note": ("host-buffer C-ABI path: pinned X/y(/C) uploaded and statistics read back every step" if world == 1 else
                        ("every rank uploads its pinned attribute columns over its own PCIe link, NCCL all-gather over NVLink; y(/C) from "
                         "rank 0; statistics read back on rank 0" if shard_upload else
                         "rank 0 uploads pinned X/y(/C) and broadcasts over NVLink (NCCL) every step; statistics read back on rank 0"))}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the dominant kernel (IRLS pass kernel), rank 0's launches of the last step
    peak = ctx.fp64_peak()
    n_full = max(int(km["n_pass_full"]), 1)
    # issue-slot roofline: every FP64 instruction (DADD, DMUL or DFMA) takes one slot of the pipe whose peak the DFMA
    # micro-benchmark measures as 2 flops per slot, so achieved = evaluations x executed FP64 instructions x 2 / time
    flops = km["evals_pass_full"] * fp64_instr(q) * 2.0
    achieved = flops / (km["pass_full_ms"] * 1e-3) / 1e12 if km["pass_full_ms"] > 0 else 0.0
    canon = km["evals_pass_full"] * f_bin(q) / (km["pass_full_ms"] * 1e-3) / 1e12 if km["pass_full_ms"] > 0 else 0.0
    hbm_peak = 6553.6
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    p_local = info["attrs"][1]
    alg_bytes = n_full * (N * p_local * 8 + km["records"] * (4 + 8 * q) + p_local * 8 * 16)   # per pass: X slab once, record stream once, partials
    traffic = traffic_src = None                  # DRAM bytes per launch of this kernel from the committed ncu --set full capture
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))
        if tj["workload"] == wl_name and tj["n_gpus"] == world:
            traffic = tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"]
            traffic_src = tj["source"]
    except Exception:
        pass
    roofline = {"kernel": "pass2_kernel<IRLS_FULL, q, diff> (binomial IRLS pass, npdr_b200/csrc/regress.cu)", "bound": "fp64",
                "achieved": achieved, "peak": peak["dfma_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peak["dfma_tflops"] if peak["dfma_tflops"] else None, "traffic": traffic,
                "traffic_unit": "bytes/launch (dram__bytes_read.sum + dram__bytes_write.sum)", "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": alg_bytes / n_full,
                "peak_source": "measured live by npdrb_measure_fp64_peak (DFMA micro-kernel; MEASURED_PEAKS.json has no FP64 figure)",
                "launches": n_full, "avg_launch_ms": km["pass_full_ms"] / n_full,
                "fp64_instr_per_eval": fp64_instr(q), "evals_per_launch_avg": km["evals_pass_full"] / n_full,
                "note": "achieved = attribute x record evaluations per launch x FP64 instructions the kernel executes per evaluation "
                        "(18 at q = 0, 31 at q = 2; SASS count, DESIGN.md 4) x 2 flops per issue slot / CUDA-event launch time, against "
                        "the measured DFMA peak: frac is the FP64 pipe's issue-slot utilisation (ncu sm__pipe_fp64_cycles_active agrees). "
                        "records = pairs after folding mutual (i,j)/(j,i) pairs into one weight-2 record.  The kernel also issues 10.6 "
                        "(q = 0) / 13.5 (q = 2) non-FP64 instructions per evaluation; an FP64 warp instruction holds the dispatch "
                        "port for 2 cycles, so the issue-limited ceiling of this mix is 36/46.6 = 77 % (q = 0), 62/75.5 = 82 % (q = 2)",
                "canonical": {"flops_per_eval_iteration": f_bin(q), "achieved": canon,
                              "frac": canon / peak["dfma_tflops"] if peak["dfma_tflops"] else None,
                              "note": "SURVEY 8d's canonical count (exp = log = 20, reciprocal = 8 flops): exceeds 1 because the kernel "
                                      "spends 6 FP64 instructions on exp, 2 on the reciprocal and replaces the per-record log by a running "
                                      "product (one multiply per record, one log per thread)"},
                "records_per_pair": km["records"] / max(info["n_pairs_used"], 1),
                "share_of_step": km["pass_full_ms"] / ms_step,
                "hbm": {"achieved": alg_bytes / (km["pass_full_ms"] * 1e-3) / 1e9 if km["pass_full_ms"] > 0 else 0.0,
                        "peak": hbm_peak, "unit": "GB/s",
                        "frac": (alg_bytes / (km["pass_full_ms"] * 1e-3) / 1e9) / hbm_peak if km["pass_full_ms"] > 0 else None}}
    dist_adds = float(p) * N * (N - 1)          # 2 DADD per element pair over the lower triangle (symmetric mode)
    if world > 1:                                               # balanced: ~half of the rank's rectangle (the rest arrives over NVLink)
        dist_adds = (1.0 if os.environ.get("NPDRB_NO_BALANCED") != "1" else 2.0) * float(p) * info["rows"][1] * N
    stage = {"distance_ms": sm["distance"], "neighbors_ms": sm["neighbors"], "prepare_ms": sm["prepare"], "regress_ms": sm["regress"],
             "irls_full_passes": passes, "pass_first_ms": km["pass_first_ms"],
             "distance_fp64_add_frac": (dist_adds / (sm["distance"] * 1e-3) / 1e12) / peak["dadd_tflops"] if sm["distance"] > 0 else None}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        N_s, p_s = (2000, 512) if wl_name != "small" else (512, 128)      # ~10 s of host work on the box's 16 cores
        ev, dt, thr, M_s = cpu_oracle_sample(wl_name, N_s, p_s, q)
        cpu_baseline = {"value": ev / dt, "unit": "evals/s", "cores": thr, "kind": "port", "seconds": dt,
                        "sample": f"first {N_s} instances x {p_s} attributes of the {wl_name} recipe (M={M_s} pairs), all stages, "
                                  f"CPU oracle (C restatement of the reference R path), OpenMP over attributes"}

    line = {"metric": "npdr_attr_x_pair_glm_evals_per_s", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": wl["scaling"],
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl_name, "description": wl["desc"], "N": N, "p": p, "q": q, "pairs_M": M,
                       "regression": "binomial", "nbd": "multisurf", "metric": "manhattan", "seed": SEED,
                       "parallelism": f"rows/{world} (distance+neighbours" + (", symmetric halves swapped in one all-to-all" if world > 1 else "")
                                      + f") -> all-gather pairs -> attrs/{world} (regression)",
                       "l2": "inputs (X = %.0f MB per GPU) exceed the 126 MB L2; no flush needed" % (N * p * 8 / 1e6)},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "stages": stage, "fp64_peak_measured": peak, "variants": variants,
            "work_saving": "exact: (i,j)/(j,i) pairs folded into one weight-2 record; the IRLS pass that would only confirm convergence "
                           "is skipped when the Newton quadratic model + cubic remainder bound puts the next deviance change >= 10x "
                           "below glm.fit's epsilon (same beta, se, iteration count); `variants` times the step without them"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
