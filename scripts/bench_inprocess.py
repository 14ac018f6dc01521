"""End-to-end timing of the in-process multi-GPU driver (npdrb_npdr_multi): ONE host process, host buffers in, statistics
out -- the call an R session makes.  Not part of the bench.py contract (that one is one process per GPU under torchrun);
wall-clock timed because the work spans several devices.  usage: python scripts/bench_inprocess.py --gpus 8 [--workload cfg5]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                      # noqa: E402
from npdr_b200 import _lib                        # noqa: E402
from npdr_b200._lib import DIFF, METRIC, NBD, REG, check, dptr, iptr   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
    ap.add_argument("--workload", default="cfg4", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    args = ap.parse_args()
    wl = bench.WORKLOADS[args.workload]
    data = bench.make_data(wl, args.gpus, torch.device("cuda", 0))
    torch.cuda.synchronize()
    N, p, q = data["N"], data["p"], data["q"]
    hXt = torch.empty((p, N), dtype=torch.float64, pin_memory=True)
    hXt.copy_(data["dX"][:, :N])
    X = hXt.numpy().T                             # (N, p) column-major view of pinned memory
    y = data["dy"].cpu().numpy()
    Cm = np.asfortranarray(data["dC"].cpu().numpy().T) if q else None
    del data
    torch.cuda.empty_cache()
    L = _lib.load()
    o = _lib.Opts()
    L.npdrb_opts_default(C.byref(o))
    o.regression_type = REG["binomial"]; o.attr_diff_type = DIFF["numeric-abs"]; o.nbd_method = NBD["multisurf"]
    o.nbd_metric = METRIC["manhattan"]; o.msurf_sd_frac = 0.5; o.n_covars = q
    for i, t in enumerate(("match-mismatch", "numeric-abs")[:q]):
        o.covar_diff_type[i] = DIFF[t]
    devs = np.arange(args.gpus, dtype=np.int32)
    times, info = [], None
    for it in range(args.warmup + args.steps):
        res = _lib.Result()
        t0 = time.perf_counter()
        check(L.npdrb_npdr_multi(args.gpus, iptr(devs), dptr(X), N, p, dptr(y), dptr(Cm), C.byref(o), C.byref(res)))
        dt = time.perf_counter() - t0
        info = {f: getattr(res, f) for f in ("n_pairs", "ms_distance", "ms_neighbors", "ms_prepare", "ms_regress", "ms_total", "irls_passes")}
        L.npdrb_result_release(C.byref(res))
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    print(json.dumps({"driver": "npdrb_npdr_multi (one host process, one thread per device)", "workload": args.workload,
                      "n_gpus": args.gpus, "N": N, "p": p, "q": q, "pairs_M": info["n_pairs"], "ms_per_call": ms,
                      "evals_per_s": float(p) * info["n_pairs"] / (ms * 1e-3), "h2d_bytes_per_device": int(N * p * 8 // args.gpus) if args.gpus > 1 else int(N * p * 8),
                      "stages_max_over_devices": info, "timing": "host wall clock around the call, host buffers in / statistics out"}))


if __name__ == "__main__":
    main()
