"""world_size-2 (and 3) gloo tests of the sharding plumbing in npdr_b200/distributed.py on CPU.
The per-rank compute is supplied by an oracle-backed Engine (tests only); the checks are that the
row-block / attribute-column sharded run reproduces the unsharded pair list and statistics exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


class OracleEngine:
    """CPU stand-in for GpuEngine (test infrastructure)."""
    device = torch.device("cpu")

    def __init__(self, X, y, C):
        from oracle import oracle as O
        self.O, self.X, self.y, self.C = O, X, y, C
        self.D = None
        self.surf = None

    def _dist(self, metric):
        if self.D is None:
            self.D = self.O.distances(self.X, metric)
        return self.D

    def neighbors_block(self, row0, nrows, nbd_method, nbd_metric, sd_frac, k, hitmiss=False):
        D = self._dist(nbd_metric)
        N = D.shape[0]
        if hitmiss:
            ri, nn, _, _, _ = self.O.nearest_neighbors_hitmiss(D, self.y, nbd_method, "precomputed", sd_frac, k)
        elif nbd_method == "relieff":
            kk = k if k else self.O.knn_surf(N, sd_frac)
            ri, nn, _ = self.O.neighbors(D, "relieff", k=kk)
        elif nbd_method == "surf":
            ri, nn, _ = self.O.neighbors(D, "surf", radius=np.full(N, self.surf))
        else:
            r, _ = self.O.radius_multisurf(D, sd_frac)
            ri, nn, _ = self.O.neighbors(D, "multisurf", radius=r)
        keep = (ri > row0) & (ri <= row0 + nrows)
        return torch.from_numpy(ri[keep].copy()), torch.from_numpy(nn[keep].copy())

    def surf_partial(self, row0, nrows, nbd_metric, centre):
        D = self._dist(nbd_metric); N = D.shape[0]
        sa = su = cu = ss = 0.0
        for g in range(row0, row0 + nrows):
            sa += D[g].sum(); up = D[g, g + 1:]
            su += up.sum(); cu += up.size; ss += ((up - centre) ** 2).sum()
        return [sa, su, cu, ss]

    def set_surf_radius(self, r):
        self.surf = r

    def regress_block(self, Ri, NN, attr0, nattr, regression_type, attr_diff_type, covar_diff_type, unique):
        ri, nn = Ri.numpy(), NN.numpy()
        nu, keep = self.O.unique_neighbors(ri, nn, self.X.shape[0])
        if unique:
            ri, nn = ri[keep], nn[keep]
        yd = self.O.pair_diffs(self.y, ri, nn, "numeric-abs" if regression_type == "lm" else "match-mismatch")
        cd = None
        if self.C is not None:
            cd = np.column_stack([self.O.pair_diffs(self.C[:, c], ri, nn, t) for c, t in enumerate(covar_diff_type)])
        st = self.O.regress_attrs(np.asfortranarray(self.X[:, attr0:attr0 + nattr]), ri, nn, yd, cd, attr_diff_type, regression_type)
        return st, nu, ri.size


def _case(seed=0):
    rng = np.random.default_rng(seed)
    N, p = 300, 37
    X = np.asfortranarray(rng.normal(size=(N, p)))
    y = (rng.random(N) < 0.5).astype(float)
    C = np.asfortranarray(np.column_stack([rng.integers(0, 2, N).astype(float), rng.normal(40, 10, N).round()]))
    return X, y, C


def _worker(rank, world, port, method, q, hitmiss=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import sys
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from npdr_b200.distributed import npdr_sharded
    X, y, C = _case()
    eng = OracleEngine(X, y, C)
    stats, info = npdr_sharded(eng, X.shape[0], X.shape[1], "binomial", "numeric-abs", method, "manhattan", 0, 0.5,
                               ("match-mismatch", "numeric-abs"), separate_hitmiss_nbds=hitmiss)
    if rank == 0:
        q.put((stats, info))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


@pytest.mark.parametrize("world,method,hitmiss", [(2, "multisurf", False), (3, "multisurf", False), (2, "surf", False),
                                                  (2, "relieff", False), (2, "multisurf", True)])
def test_sharded_equals_unsharded(world, method, hitmiss):
    from oracle import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, method, q, hitmiss)) for r in range(world)]
    for p_ in procs:
        p_.start()
    stats, info = q.get(timeout=240)
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    X, y, C = _case()
    if hitmiss:
        ri, nn, D, _, _ = O.nearest_neighbors_hitmiss(X, y, method)
    else:
        ri, nn, D, r = O.nearest_neighbors(X, method)
    assert info["n_pairs"] == ri.size and sum(info["pair_counts"]) == ri.size
    cd = np.column_stack([O.pair_diffs(C[:, 0], ri, nn, "match-mismatch"), O.pair_diffs(C[:, 1], ri, nn, "numeric-abs")])
    exp = O.regress_attrs(X, ri, nn, O.pair_diffs(y, ri, nn, "match-mismatch"), cd)
    assert stats.shape == exp.shape
    if method == "surf":
        np.testing.assert_allclose(stats, exp, rtol=1e-9, equal_nan=True)    # radius from combined moments
    else:
        assert np.array_equal(stats, exp, equal_nan=True)


def test_all_gather_varlen_single_process():
    from npdr_b200.distributed import partition_attrs, partition_rows
    assert partition_rows(300, 2) == [(0, 256), (256, 44)]
    assert partition_attrs(37, 3) == [(0, 13), (13, 12), (25, 12)]
