"""Sharding NPDR's hot path over the GPUs of one node (one process per GPU, torch.distributed).

The path has two embarrassingly parallel stages joined by ONE exchange:
  A+B  distances + neighbour lists : row-block sharded (rank g owns instances I_g and needs D[I_g, :]).  D is
       symmetric, so each rank computes its diagonal square and a checkerboard HALF of every off-diagonal square and
       the halves are swapped in one all-to-all over NVLink (packed, transposed 128 x 128 tiles): ~N^2/(2G) instance
       pairs per rank instead of N^2/G, bit-identical to the unsharded matrix.  multiSURF radii are per-row, so no
       reduction is needed; SURF needs two scalar all-reduces.
  X    all-gather of the variable-length (Ri, NN) segments -- concatenation in rank order IS the
       reference's order, because segments are contiguous in Ri.
  C    regression : attribute-column sharded, every rank holds the full pair list; statistics are
       gathered to rank 0.  No data-path reduction -> results are identical for any world size.

`Engine` is the per-rank compute interface; `GpuEngine` drives libnpdr_b200 sessions.  The collective
plumbing below only sees torch tensors, so it is exercised on CPU with the gloo backend in tests/.
"""
import os

import numpy as np
import torch
import torch.distributed as dist

ROW_ALIGN = 128          # distance tiles are 128 rows: row blocks start on tile boundaries


def partition_rows(N, world, align=ROW_ALIGN):
    """Contiguous, tile-aligned row blocks [(row0, nrows)] covering 0..N; trailing ranks may be empty."""
    n_tiles = -(-N // align)
    base, extra = divmod(n_tiles, world)
    out, t0 = [], 0
    for r in range(world):
        nt = base + (1 if r < extra else 0)
        row0 = min(t0 * align, N)
        row1 = min((t0 + nt) * align, N)
        out.append((row0, row1 - row0))
        t0 += nt
    return out


def partition_attrs(p, world):
    """Contiguous attribute blocks [(attr0, nattr)]."""
    base, extra = divmod(p, world)
    out, a0 = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((a0, n))
        a0 += n
    return out


class Engine:
    """Per-rank compute interface used by npdr_sharded()."""
    device = torch.device("cpu")

    def neighbors_block(self, row0, nrows, nbd_method, nbd_metric, sd_frac, k, hitmiss=False):
        """-> (Ri, NN) int32 tensors on self.device for the owned rows (1-based, reference order).
        hitmiss: separate hit/miss neighbourhoods (nearestNeighborsSeparateHitMiss)."""
        raise NotImplementedError

    def surf_partial(self, row0, nrows, nbd_metric, centre):
        """-> 4 partial moments (see npdrb_session_surf_partial)."""
        raise NotImplementedError

    def set_surf_radius(self, radius):
        raise NotImplementedError

    def surf_radius_from_moments(self, m0, m1, N, sd_frac):
        """-> (radius, eps): radius from the combined moments and the bound on its distance from the reference's value."""
        raise NotImplementedError

    def surf_band_count(self, row0, nrows, nbd_metric, lo, hi):
        """-> number of distances of the owned block inside [lo, hi]."""
        raise NotImplementedError

    def distance_block(self, row0, nrows, nbd_metric):
        """-> the owned block of the distance matrix, float64 tensor [nrows, N] on self.device."""
        raise NotImplementedError

    def surf_radius_exact(self, D, sd_frac):
        """-> the reference's SURF radius from the full N x N matrix (tensor), replayed in R's summation order."""
        raise NotImplementedError

    def regress_block(self, Ri, NN, attr0, nattr, regression_type, attr_diff_type, covar_diff_type, unique):
        """-> (stats float64 [nattr, 8] numpy, n_unique, n_pairs_used)."""
        raise NotImplementedError


def all_gather_varlen(t, group=None):
    """All-gather 1-D tensors of different lengths; returns the concatenation in rank order."""
    world = dist.get_world_size(group)
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(max(counts), 1)
    pad = torch.zeros(mx, dtype=t.dtype, device=t.device)
    pad[: t.numel()] = t
    bufs = [torch.empty(mx, dtype=t.dtype, device=t.device) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)]), counts


def exchange_balanced(engine, parts, rank, nbd_metric, group=None):
    """Row block of the symmetric distance matrix with the work split evenly over the ranks.  Each rank computes its
    diagonal square and a checkerboard half of every off-diagonal square (engine.balanced_local), packs the tiles its
    peers need (transposed, peers in rank order), ONE all-to-all (NCCL over NVLink) swaps the halves, and
    engine.balanced_unpack drops the received tiles into the block.  Ranks without rows join with empty messages."""
    world = len(parts)
    live = [g for g, (_, n) in enumerate(parts) if n > 0]
    row_starts = [parts[g][0] for g in live] + [parts[live[-1]][0] + parts[live[-1]][1]]
    send_counts, recv_counts = [0] * world, [0] * world
    mine = rank in live
    if mine:
        sc, rc = engine.balanced_local(row_starts, live.index(rank), nbd_metric)
        for j, g in enumerate(live):
            send_counts[g], recv_counts[g] = int(sc[j]), int(rc[j])
    n_send, n_recv = sum(send_counts), sum(recv_counts)
    send = engine.balanced_pack(n_send) if mine else torch.empty(1, dtype=torch.float64, device=engine.device)
    recv = torch.empty(max(n_recv, 1), dtype=torch.float64, device=engine.device)
    dist.all_to_all_single(recv[:n_recv], send[:n_send], recv_counts, send_counts, group=group)
    if mine:
        engine.balanced_unpack(recv, n_recv)


def pair_list_hash(ri, nn):
    """Order-sensitive 64-bit hash of a (Ri, NN) pair list held in torch tensors (wrapping int64 arithmetic): equal lists
    <=> equal hashes for all practical purposes; used to assert that N ranks and one rank find the same list."""
    if ri.numel() == 0:
        return 0
    i = ri.to(torch.int64); j = nn.to(torch.int64)
    pos = torch.arange(1, ri.numel() + 1, dtype=torch.int64, device=ri.device)
    h = (i * 1000003 + j * 7919 + 12345) * (pos * 2654435761 + 1)
    h = h ^ (h >> 29)
    return int(h.sum().item())


def npdr_sharded(engine, N, p, regression_type="binomial", attr_diff_type="numeric-abs", nbd_method="multisurf",
                 nbd_metric="manhattan", knn=0, msurf_sd_frac=0.5, covar_diff_type=(), neighbor_sampling="none",
                 group=None, separate_hitmiss_nbds=False, pair_hash=False):
    """Run the path sharded over the default process group.  Every rank returns
    (stats [p, 8] on rank 0 else None, info dict).  pair_hash: also return pair_list_hash() of the gathered list."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    row0, nrows = partition_rows(N, world)[rank]
    dev = engine.device

    if nbd_method == "surf" and separate_hitmiss_nbds and world > 1:
        raise NotImplementedError("surf hit/miss radii need the whole distance matrix on one device")
    if world > 1 and hasattr(engine, "balanced_local") and os.environ.get("NPDRB_NO_BALANCED") != "1":
        exchange_balanced(engine, partition_rows(N, world), rank, nbd_metric, group)    # collective: every rank calls it
    if nbd_method == "surf" and world > 1:
        # one global mean / sd: two scalar all-reduces (R/nearestNeighbors.R:234-241), then the guard band: if no distance
        # of any block lies within the radius' error bound the moment radius provably selects the reference's pairs;
        # otherwise the blocks are gathered on rank 0 and R's sequential long-double chains are replayed exactly
        m0 = torch.tensor(engine.surf_partial(row0, nrows, nbd_metric, 0.0) if nrows else [0.0] * 4, dtype=torch.float64, device=dev)
        dist.all_reduce(m0, group=group)
        mean_u = float(m0[1] / m0[2])
        m1 = torch.tensor(engine.surf_partial(row0, nrows, nbd_metric, mean_u) if nrows else [0.0] * 4, dtype=torch.float64, device=dev)
        dist.all_reduce(m1, group=group)
        radius, eps = engine.surf_radius_from_moments(m0.tolist(), m1.tolist(), N, msurf_sd_frac)
        band = torch.tensor([engine.surf_band_count(row0, nrows, nbd_metric, radius - eps, radius + eps) if nrows else 0],
                            dtype=torch.int64, device=dev)
        dist.all_reduce(band, group=group)
        if int(band.item()) > 0 or os.environ.get("NPDRB_SURF_FORCE_EXACT") == "1":
            blk = engine.distance_block(row0, nrows, nbd_metric) if nrows else torch.zeros((0, N), dtype=torch.float64, device=dev)
            flat, _ = all_gather_varlen(blk.reshape(-1), group)       # rare path: the whole matrix on every rank
            rad = torch.tensor([engine.surf_radius_exact(flat.reshape(N, N), msurf_sd_frac) if rank == 0 else 0.0],
                               dtype=torch.float64, device=dev)
            dist.broadcast(rad, 0, group=group)
            radius = float(rad.item())
        engine.set_surf_radius(radius)

    if nrows > 0:
        ri, nn = engine.neighbors_block(row0, nrows, nbd_method, nbd_metric, msurf_sd_frac, knn, separate_hitmiss_nbds)
    else:
        ri = torch.zeros(0, dtype=torch.int32, device=dev); nn = torch.zeros(0, dtype=torch.int32, device=dev)
    if world > 1:
        ri_all, counts = all_gather_varlen(ri, group)
        nn_all, _ = all_gather_varlen(nn, group)
    else:
        ri_all, nn_all, counts = ri, nn, [ri.numel()]

    attr0, nattr = partition_attrs(p, world)[rank]
    if nattr > 0:
        stats, n_unique, n_used = engine.regress_block(ri_all, nn_all, attr0, nattr, regression_type, attr_diff_type,
                                                       covar_diff_type, neighbor_sampling == "unique")
        st = torch.from_numpy(np.ascontiguousarray(stats, dtype=np.float64)).to(dev)
    else:
        st = torch.zeros((0, 8), dtype=torch.float64, device=dev); n_unique, n_used = -1, -1
    info = {"n_pairs": int(sum(counts)), "pair_counts": counts, "n_unique_pairs": n_unique, "n_pairs_used": n_used,
            "rows": (row0, nrows), "attrs": (attr0, nattr)}
    if pair_hash:
        info["pair_hash"] = pair_list_hash(ri_all, nn_all)
    if world == 1:
        return st.cpu().numpy(), info
    flat, fcounts = all_gather_varlen(st.reshape(-1), group)      # tiny: p x 8 doubles in total
    out = flat.reshape(-1, 8).cpu().numpy() if rank == 0 else None
    return out, info


def gather_columns_pipelined(dX, world, ncols, group=None):
    """Replicate an attribute-sharded matrix without stalling the distance stage.  dX: [p, ldx] device tensor (row a = attribute
    column a) whose rows [rank * ncols, (rank + 1) * ncols) this rank has already filled.  One asynchronous NCCL broadcast per
    rank's shard, issued in rank order = attribute order; returns (slab_cols, events) with events[g] firing when shard g has
    landed here.  Handed to GpuEngine(slab_events=...) the distance kernel consumes shard g while shards g+1.. are still on
    the wire (npdrb_session_set_device_slab_events) -- bit-identical to gathering first."""
    side = torch.cuda.Stream(device=dX.device)
    events = []
    for g in range(world):
        work = dist.broadcast(dX[g * ncols:(g + 1) * ncols], src=g, group=group, async_op=True)
        with torch.cuda.stream(side):
            work.wait()                                                 # the side stream (not the caller's) waits for the transfer
            ev = torch.cuda.Event()
            ev.record(side)
        events.append(ev)
    return ncols, events


def cyclic_blocks(p, world, n_sub=5):
    """Column bounds of the block-cyclic attribute layout of upload_gather_pipelined: world * n_sub blocks of (nearly) equal
    width, every bound but the last a multiple of 16 (the distance kernel's attribute chunk); block b = c * world + r is
    sub-slab c of rank r.  -> list of world * n_sub + 1 bounds, or None when p is too small to give every block 16 columns."""
    chunks = p // 16
    n_blocks = world * n_sub
    if world < 2 or chunks < n_blocks:
        return None
    return [16 * (b * chunks // n_blocks) for b in range(n_blocks)] + [p]


def upload_gather_pipelined(ctx, dX, ldx, host_cols, rank, world, starts, group=None):
    """Upload AND replicate an attribute-sharded matrix underneath the distance stage.  Ownership is block-cyclic (starts =
    cyclic_blocks(p, world, n_sub)), so the order in which sub-slabs become available everywhere -- sub-slab 0 of every rank,
    then sub-slab 1, ... -- IS the attribute order k = 0..p-1 the distance kernel consumes.  A helper thread stages sub-slab c
    of this rank's host columns (host_cols: N x (this rank's columns, blocks in order), Fortran order, pageable is fine)
    through the library's pinned ring, then issues one asynchronous NCCL broadcast per rank for that sub-slab, records an
    event per block and raises its host flag; the caller passes the result to GpuEngine(slab_events=...) at once and the
    distance call (npdrb_session_set_device_slab_events_v with ready flags) starts consuming block 0 while later sub-slabs
    are still crossing PCIe.  -> (starts, events, ready, thread, errors); join the thread after the distance stage."""
    import threading
    dev = dX.device
    n_blocks = len(starts) - 1
    n_sub = n_blocks // world
    side = torch.cuda.Stream(device=dev)
    issue = torch.cuda.Stream(device=dev)        # the collectives must not queue behind the distance launches of the caller's stream
    events = [torch.cuda.Event() for _ in range(n_blocks)]
    for e in events:
        e.record(side)                           # creates the handles; re-recorded when the block's transfer has been issued
    ready = np.zeros(n_blocks, dtype=np.int32)
    err = []

    def work():
        try:
            torch.cuda.set_device(dev)
            with torch.cuda.stream(issue):
                off = 0
                for c in range(n_sub):
                    b_own = c * world + rank
                    w_own = starts[b_own + 1] - starts[b_own]
                    ctx.upload_columns(dX[starts[b_own]:starts[b_own + 1]].data_ptr(), ldx, host_cols[:, off:off + w_own])
                    off += w_own
                    for r in range(world):
                        b = c * world + r
                        work_ = dist.broadcast(dX[starts[b]:starts[b + 1]], src=r, group=group, async_op=True)
                        with torch.cuda.stream(side):
                            work_.wait()
                            events[b].record(side)
                        ready[b] = 1
        except Exception as ex:                  # unblock the consumer; it will fail on the data check
            err.append(ex)
            ready[:] = 1

    th = threading.Thread(target=work, name="npdr-upload-gather")
    th.start()
    return starts, events, ready, th, err


class GpuEngine(Engine):
    """libnpdr_b200 session on this rank's GPU.  Inputs are uploaded once (or adopted from device tensors)."""

    def __init__(self, X=None, y=None, C=None, device_index=0, dX=None, ldx=None, dy=None, dC=None, N=None, p=None, q=None,
                 slab_events=None):
        from . import _lib
        self._lib = _lib
        self.device = torch.device("cuda", device_index)
        self.ctx = _lib.context(device_index)
        if X is not None:
            N, p = X.shape
            q = 0 if C is None else (1 if np.ndim(C) == 1 else np.shape(C)[1])
        self.N, self.p, self.q = N, p, q
        self.session = _lib.Session(self.ctx, N, p, q)
        if X is not None:
            self.session.upload(X, y, C)
        else:
            self._keep = (dX, dy, dC, slab_events)
            if slab_events is None:
                torch.cuda.synchronize(self.device)  # the producers of the adopted tensors ran on torch's streams, the library has its own
            else:
                torch.cuda.current_stream(self.device).synchronize()   # y / C only: the column slabs of X are still in flight
            self.session.set_device_inputs(dX.data_ptr(), ldx, dy.data_ptr(), dC.data_ptr() if dC is not None else 0)
            if slab_events is not None:              # (slab_cols, [torch.cuda.Event][, ready flags]): X arrives slab by slab
                self.session.set_device_slab_events(slab_events[0], [e.cuda_event for e in slab_events[1]],
                                                    slab_events[2] if len(slab_events) > 2 else None)
        self._metric_done = None

    def close(self):
        self.session.close()

    def _ensure_distances(self, row0, nrows, nbd_metric):
        key = (row0, nrows, nbd_metric)
        if self._metric_done != key:
            self.session.distances(nbd_metric, False, row0, nrows)
            self._metric_done = key

    # balanced symmetric distances: the three per-rank primitives exchange_balanced() drives
    def balanced_local(self, row_starts, index, nbd_metric):
        self._bal_key = (row_starts[index], row_starts[index + 1] - row_starts[index], nbd_metric)
        return self.session.distances_balanced(nbd_metric, row_starts, index)

    def balanced_pack(self, n_send):
        send = torch.empty(max(n_send, 1), dtype=torch.float64, device=self.device)
        if n_send:
            self.session.balanced_pack(send.data_ptr())
        torch.cuda.synchronize(self.device)          # the library's stream has finished packing
        return send

    def balanced_unpack(self, recv, n_recv):
        torch.cuda.synchronize(self.device)
        if n_recv:
            self.session.balanced_unpack(recv.data_ptr())
        self._metric_done = self._bal_key

    def neighbors_block(self, row0, nrows, nbd_method, nbd_metric, sd_frac, k, hitmiss=False):
        self._ensure_distances(row0, nrows, nbd_metric)
        M, _ = (self.session.neighbors_hitmiss if hitmiss else self.session.neighbors)(nbd_method, sd_frac, k)
        ri = torch.empty(max(M, 1), dtype=torch.int32, device=self.device)[:M]
        nn = torch.empty(max(M, 1), dtype=torch.int32, device=self.device)[:M]
        if M:
            self.session.copy_local_pairs_device(ri.data_ptr(), nn.data_ptr())
        return ri, nn

    def surf_partial(self, row0, nrows, nbd_metric, centre):
        self._ensure_distances(row0, nrows, nbd_metric)
        return self.session.surf_partial(centre).tolist()

    def set_surf_radius(self, radius):
        self.session.set_surf_radius(radius)

    def surf_radius_from_moments(self, m0, m1, N, sd_frac):
        return self._lib.surf_radius_from_moments(m0, m1, N, sd_frac)

    def surf_band_count(self, row0, nrows, nbd_metric, lo, hi):
        self._ensure_distances(row0, nrows, nbd_metric)
        return self.session.surf_band_count(lo, hi)

    def distance_block(self, row0, nrows, nbd_metric):
        self._ensure_distances(row0, nrows, nbd_metric)
        return torch.from_numpy(self.session.copy_distances()).to(self.device)

    def surf_radius_exact(self, D, sd_frac):
        return self._lib.surf_radius_exact_host(D.cpu().numpy(), sd_frac)

    def regress_block(self, Ri, NN, attr0, nattr, regression_type, attr_diff_type, covar_diff_type, unique):
        torch.cuda.synchronize(self.device)          # NCCL results complete before the library's stream reads them
        self._pairs = (Ri.contiguous(), NN.contiguous())
        self.session.import_pairs(self._pairs[0].data_ptr(), self._pairs[1].data_ptr(), Ri.numel())
        n_unique = self.session.unique_pairs(keep_unique_only=unique)
        stats = self.session.regress(attr0, nattr, regression_type, attr_diff_type, covar_diff_type)
        n_used = n_unique if unique else Ri.numel()
        return stats, n_unique, n_used
