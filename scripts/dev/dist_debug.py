"""Development aid: where does the GPU distance matrix differ from the oracle at bench sizes?"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from npdr_b200 import _lib
from oracle import oracle as O

dev = torch.device("cuda", 0)
data = bench.make_data(bench.WORKLOADS["cfg4"], 1, dev)
N, p = data["N"], data["p"]
X = np.asfortranarray(data["dX"][:, :N].T.cpu().numpy())
rows = np.r_[0:256, 1920:2048, 3872:4000]
for pp in (20000, 5000, 1024):
    for kern in ("tma", "simple"):
        os.environ["NPDRB_DIST_KERNEL"] = kern
        s = _lib.Session(_lib.context(0), N, pp, 0)
        s.set_device_inputs(data["dX"].data_ptr(), data["ldx"], data["dy"].data_ptr(), 0)
        s.distances("manhattan", False, 0, N)
        D = s.copy_distances()
        s.close()
        E = O.distances2(np.asfortranarray(X[:, :pp]), np.asfortranarray(X[rows][:, :pp]), "manhattan").T   # rows x N
        bad = D[rows] != E
        tiles = sorted({(int(rows[r]) // 128, int(c) // 128) for r, c in zip(*np.nonzero(bad))})
        print(f"p={pp} kernel={kern}: {int(bad.sum())} mismatching entries of {bad.size}; tiles {tiles[:40]}", flush=True)
        if bad.any():
            r, c = [int(v[0]) for v in np.nonzero(bad)]
            print("   first:", rows[r], c, D[rows[r], c], E[r, c], "sym:", D[c, rows[r]])
