"""Wall-clock breakdown of one resident-input step of bench.py's workload (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import npdr_b200 as nb
from npdr_b200.distributed import GpuEngine

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg4"]
dev = torch.device("cuda", 0)
ctx = nb.context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
data = bench.make_data(wl, 1, dev)
N, p, q = data["N"], data["p"], data["q"]


def T():
    torch.cuda.synchronize()
    return time.perf_counter()


for it in range(4):
    t = [T()]
    eng = GpuEngine(device_index=0, dX=data["dX"], ldx=data["ldx"], dy=data["dy"], dC=data["dC"], N=N, p=p, q=q); t.append(T())
    ri, nn = eng.neighbors_block(0, N, "multisurf", "manhattan", 0.5, 0); t.append(T())
    torch.cuda.synchronize(); eng.session.import_pairs(ri.data_ptr(), nn.data_ptr(), ri.numel()); t.append(T())
    nu = eng.session.unique_pairs(False); t.append(T())
    st = eng.session.regress(0, p, "binomial", "numeric-abs", data["covar_diff_type"]); t.append(T())
    sm = eng.session.stage_ms(); km = eng.session.kernel_ms()
    eng.close(); t.append(T())
    names = ["engine", "dist+nbr", "import", "unique", "regress", "close"]
    print(it, " ".join(f"{n}={1e3 * (b - a):.1f}ms" for n, a, b in zip(names, t, t[1:])), f"total={1e3 * (t[-1] - t[0]):.1f}ms",
          {k: round(v, 1) for k, v in sm.items()}, {k: round(v, 1) for k, v in km.items() if k.endswith("ms")}, "M", ri.numel(), "unique", nu)
