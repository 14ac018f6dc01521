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

    def surf_radius_from_moments(self, m0, m1, N, sd_frac):
        from npdr_b200 import _lib                      # host-only entry points of the library: no GPU needed
        return _lib.surf_radius_from_moments(m0, m1, N, sd_frac)

    def surf_band_count(self, row0, nrows, nbd_metric, lo, hi):
        blk = self._dist(nbd_metric)[row0:row0 + nrows]
        return int(((blk >= lo) & (blk <= hi)).sum())

    def distance_block(self, row0, nrows, nbd_metric):
        return torch.from_numpy(np.ascontiguousarray(self._dist(nbd_metric)[row0:row0 + nrows]))

    def surf_radius_exact(self, D, sd_frac):
        from npdr_b200 import _lib
        return _lib.surf_radius_exact_host(D.numpy(), sd_frac)

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


class BalancedOracleEngine(OracleEngine):
    """Numpy mirror of the balanced-distance primitives of libnpdr_b200 (distance.cu: build_tiles / exchange_lists /
    pack / unpack), so the all-to-all plumbing of exchange_balanced() runs under gloo on CPU."""
    T = 128

    @staticmethod
    def _row_owner_computes(I, J):
        return (I < J) == ((I + J) % 2 == 0)

    def balanced_local(self, row_starts, index, nbd_metric):
        T = self.T
        Dfull = self.O.distances(self.X, nbd_metric)
        N = Dfull.shape[0]
        ts = [-(-r // T) for r in row_starts]
        a, b = ts[index], ts[index + 1]
        self.r0, self.r1 = row_starts[index], row_starts[index + 1]
        self.block = np.full((self.r1 - self.r0, N), np.nan)            # NaN = not computed here
        self.send_tiles, self.recv_tiles, sc, rc = [], [], [], []
        n_jt = -(-N // T)
        for I in range(a, b):
            for J in range(n_jt):
                own = a <= J < b
                if own or self._row_owner_computes(I, J):
                    i0, i1 = I * T, min((I + 1) * T, self.r1)
                    j0, j1 = J * T, min((J + 1) * T, N)
                    self.block[i0 - self.r0:i1 - self.r0, j0:j1] = Dfull[i0:i1, j0:j1]
        for h in range(len(ts) - 1):
            if h == index:
                sc.append(0); rc.append(0); continue
            c, d = ts[h], ts[h + 1]
            snd = [(I, J) for I in range(a, b) for J in range(c, d) if self._row_owner_computes(I, J)]
            rcv = [(I, J) for J in range(c, d) for I in range(a, b) if self._row_owner_computes(J, I)]
            self.send_tiles += snd; self.recv_tiles += rcv
            sc.append(len(snd) * T * T); rc.append(len(rcv) * T * T)
        return sc, rc

    def balanced_pack(self, n_send):
        T = self.T
        N = self.block.shape[1]
        out = np.zeros((len(self.send_tiles), T, T))
        for t, (I, J) in enumerate(self.send_tiles):                     # tile (I, J) transposed: buf[t][c][r]
            i0, i1 = I * T - self.r0, min((I + 1) * T, self.r1) - self.r0
            j0, j1 = J * T, min((J + 1) * T, N)
            out[t, :j1 - j0, :i1 - i0] = self.block[i0:i1, j0:j1].T
        assert out.size == n_send
        return torch.from_numpy(out.reshape(-1)) if n_send else torch.zeros(1, dtype=torch.float64)

    def balanced_unpack(self, recv, n_recv):
        T = self.T
        N = self.block.shape[1]
        buf = recv[:n_recv].numpy().reshape(-1, T, T)
        for t, (I, J) in enumerate(self.recv_tiles):                     # rows I (mine), columns J (the sender's rows)
            i0, i1 = I * T - self.r0, min((I + 1) * T, self.r1) - self.r0
            j0, j1 = J * T, min((J + 1) * T, N)
            self.block[i0:i1, j0:j1] = buf[t, :i1 - i0, :j1 - j0]

    def neighbors_block(self, row0, nrows, nbd_method, nbd_metric, sd_frac, k, hitmiss=False):
        # inside npdr_sharded(): the exchanged block must be this rank's rows of the full matrix before it is used
        assert self.block is not None and (self.r0, self.r1) == (row0, row0 + nrows)
        assert np.array_equal(self.block, self._dist(nbd_metric)[row0:row0 + nrows, :])
        return super().neighbors_block(row0, nrows, nbd_method, nbd_metric, sd_frac, k, hitmiss)


def _worker_balanced(rank, world, port, N, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import sys
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from npdr_b200.distributed import exchange_balanced, partition_rows
    rng = np.random.default_rng(5)
    X = np.asfortranarray(rng.normal(size=(N, 6)))
    eng = BalancedOracleEngine(X, None, None)
    parts = partition_rows(N, world)
    exchange_balanced(eng, parts, rank, "manhattan")
    ok = True
    if parts[rank][1] > 0:
        D = eng.O.distances(X, "manhattan")
        ok = bool(np.array_equal(eng.block, D[parts[rank][0]:parts[rank][0] + parts[rank][1], :]))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,N", [(2, 300), (3, 700), (4, 300)])
def test_balanced_distance_exchange_gloo(world, N):
    """exchange_balanced(): every rank ends with its full row block although it computed only half of the off-diagonal
    squares; (4, 300) has a rank without rows, which still joins the all-to-all."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_balanced, args=(r, world, port, N, q)) for r in range(world)]
    for p_ in procs:
        p_.start()
    res = dict(q.get(timeout=240) for _ in range(world))
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    assert all(res[r] for r in range(world)), res


def _case(seed=0):
    rng = np.random.default_rng(seed)
    N, p = 300, 37
    X = np.asfortranarray(rng.normal(size=(N, p)))
    y = (rng.random(N) < 0.5).astype(float)
    C = np.asfortranarray(np.column_stack([rng.integers(0, 2, N).astype(float), rng.normal(40, 10, N).round()]))
    return X, y, C


def _worker(rank, world, port, method, q, hitmiss=False, balanced=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import sys
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from npdr_b200.distributed import npdr_sharded
    X, y, C = _case()
    eng = (BalancedOracleEngine if balanced else OracleEngine)(X, y, C)
    stats, info = npdr_sharded(eng, X.shape[0], X.shape[1], "binomial", "numeric-abs", method, "manhattan", 0, 0.5,
                               ("match-mismatch", "numeric-abs"), separate_hitmiss_nbds=hitmiss)
    if rank == 0:
        q.put((stats, info))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


@pytest.mark.parametrize("world,method,hitmiss,balanced", [(2, "multisurf", False, False), (3, "multisurf", False, True),
                                                           (2, "surf", False, True), (2, "relieff", False, False),
                                                           (2, "multisurf", True, True)])
def test_sharded_equals_unsharded(world, method, hitmiss, balanced):
    from oracle import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, method, q, hitmiss, balanced)) for r in range(world)]
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
    # surf: the radius comes from combined moments, but no distance lies inside its guard band, so the pair list -- and with
    # it every statistic -- is the reference's bit for bit
    assert np.array_equal(stats, exp, equal_nan=True)


def test_surf_exact_replay_under_gloo(monkeypatch):
    """NPDRB_SURF_FORCE_EXACT=1: the ranks gather the matrix and rank 0 replays R's sequential long-double chains."""
    monkeypatch.setenv("NPDRB_SURF_FORCE_EXACT", "1")
    test_sharded_equals_unsharded(2, "surf", False, True)


def test_surf_exact_host_replay_matches_oracle_and_band_bound():
    """The library's host replay of R's SURF radius equals the oracle's bit for bit; the moment radius lies inside its own
    error bound of it (the guarantee the guard band rests on)."""
    from oracle import oracle as O
    from npdr_b200 import _lib
    rng = np.random.default_rng(8)
    for N in (37, 300, 1500):
        X = np.asfortranarray(rng.normal(size=(N, 40)))
        D = O.distances(X, "manhattan")
        for sd_frac in (0.5, 0.0, 1.3):
            r_exact = _lib.surf_radius_exact_host(D, sd_frac)
            assert r_exact == O.radius_surf(D, sd_frac)
            iu = np.triu_indices(N, 1)
            up = D[iu]
            m0 = [D.sum(), up.sum(), float(up.size), float((up ** 2).sum())]
            mean_u = m0[1] / m0[2]
            m1 = [m0[0], m0[1], m0[2], float(((up - mean_u) ** 2).sum())]
            r, eps = _lib.surf_radius_from_moments(m0, m1, N, sd_frac)
            assert abs(r - r_exact) <= eps and eps < 1e-9 * abs(r_exact)


def test_all_gather_varlen_single_process():
    from npdr_b200.distributed import partition_attrs, partition_rows
    assert partition_rows(300, 2) == [(0, 256), (256, 44)]
    assert partition_attrs(37, 3) == [(0, 13), (13, 12), (25, 12)]


def test_cyclic_blocks_layout():
    """Block-cyclic attribute layout of the upload pipelines: world * n_sub blocks covering 0..p in order, every bound but the
    last a multiple of 16 (the distance kernel's attribute chunk), no empty block, nearly equal widths; None when p is too
    small to give every block a chunk."""
    from npdr_b200.distributed import cyclic_blocks
    for p, world in ((50000, 8), (160000, 8), (40000, 2), (300, 2), (4096, 2), (20000 * 4, 4), (1600, 3), (16 * 10, 2)):
        st = cyclic_blocks(p, world)
        assert st is not None and st[0] == 0 and st[-1] == p and (len(st) - 1) % world == 0
        assert all(b % 16 == 0 for b in st[:-1]) and all(st[i] < st[i + 1] for i in range(len(st) - 1))
        w = [st[i + 1] - st[i] for i in range(len(st) - 1)]
        assert max(w) - min(w) <= 16 + (p % 16)
        # every rank owns n_sub blocks, one per round, and the rounds are consumed in attribute order
        n_sub = (len(st) - 1) // world
        owned = {r: [c * world + r for c in range(n_sub)] for r in range(world)}
        assert sorted(b for bl in owned.values() for b in bl) == list(range(world * n_sub))
    assert cyclic_blocks(100, 8) is None and cyclic_blocks(5000, 1) is None
