"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed for
the plumbing.

Stage 1 (CWT -> admittance / coherence) shards by wavelet SCALE: every (scale, azimuth) plane is
independent given the grid's spectrum and the azimuth reduction / jackknife of a scale is local
when all 11 azimuths of that scale live on one GPU.  Stage 2 (per-cell fit) shards by grid ROW
(cells are independent but need every scale).  Between them the four (nx, ny, ns) maps are
re-sharded from [scale block][all rows] to [all scales][row block] -- the one real exchange step of
the path -- with a single all-to-all per map; the receive side needs no unpacking because scale
blocks are assigned to ranks in rank order.  The fitted (nx/nn, rows) maps are gathered at the end.

The reference has no counterpart (single process; ``estimate_grid(parallel=True)`` raises,
classes.py:1211-1216).  Everything here is host logic over torch tensors: the same code runs on
CPU tensors with the gloo backend (tests/test_sharding.py) and on CUDA tensors with NCCL.
"""
import torch
import torch.distributed as dist

__all__ = ["scale_range", "balanced_cuts", "deal_scales", "row_range", "rows_per_shard", "RowResharder", "PeerShards", "gather_rows"]


def scale_range(ns, world, rank, cost=None):
    """Contiguous block of scales of ``rank``.  Without ``cost`` the first ns % world ranks hold one
    more scale; with a per-scale cost vector (Plan.scale_cost()) the cuts minimise the most expensive
    block (small scales have wide supports and cost several times more than large ones)."""
    if cost is None:
        base, extra = divmod(ns, world)
        lo = rank * base + min(rank, extra)
        return lo, lo + base + (1 if rank < extra else 0)
    cuts = balanced_cuts([float(c) for c in cost], world)
    return cuts[rank], cuts[rank + 1]


def balanced_cuts(cost, parts):
    """Cut points [0, ..., n] of the contiguous partition of ``cost`` into ``parts`` blocks (possibly
    empty) with the smallest maximum block sum (dynamic programme, n and parts are tiny)."""
    n = len(cost)
    pre = [0.0]
    for c in cost:
        pre.append(pre[-1] + c)
    inf = float("inf")
    best = [[inf] * (n + 1) for _ in range(parts + 1)]
    arg = [[0] * (n + 1) for _ in range(parts + 1)]
    best[0][0] = 0.0
    for p in range(1, parts + 1):
        for j in range(n + 1):
            for i in range(j + 1):
                v = max(best[p - 1][i], pre[j] - pre[i])
                if v < best[p][j]:
                    best[p][j], arg[p][j] = v, i
    cuts = [n]
    for p in range(parts, 0, -1):
        cuts.append(arg[p][cuts[-1]])
    return cuts[::-1]


def deal_scales(cost, world):
    """Longest-processing-time assignment of scales to ranks: scales sorted by decreasing cost, each given
    to the least loaded rank.  Returns one ascending list of scale indices per rank.  Only the one-sided
    exchange (PeerShards) can use it: its epilogue stores every scale at its own index."""
    order = sorted(range(len(cost)), key=lambda s: (-float(cost[s]), s))
    load = [0.0] * world
    mine = [[] for _ in range(world)]
    for s in order:
        r = min(range(world), key=lambda q: (load[q], q))
        load[r] += float(cost[s])
        mine[r].append(s)
    return [sorted(m) for m in mine]


def rows_per_shard(ny, world):
    return -(-ny // world)


def row_range(ny, world, rank):
    """Grid rows (y) of ``rank``: blocks of ceil(ny / world) rows, the last ones possibly short or empty."""
    r = rows_per_shard(ny, world)
    lo = min(rank * r, ny)
    return lo, min(lo + r, ny)


def _all_to_all(recv, send, recv_splits, send_splits, group=None):
    """all_to_all_single where the backend has it (NCCL), point-to-point otherwise (gloo on CPU)."""
    backend = dist.get_backend(group)
    if backend == "nccl":
        dist.all_to_all_single(recv, send, recv_splits, send_splits, group=group)
        return
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    so = [0]
    for s in send_splits:
        so.append(so[-1] + s)
    ro = [0]
    for s in recv_splits:
        ro.append(ro[-1] + s)
    reqs = []
    for peer in range(world):
        if peer == rank:
            recv[ro[peer]:ro[peer + 1]].copy_(send[so[peer]:so[peer + 1]])
            continue
        if send_splits[peer]:
            reqs.append(dist.isend(send[so[peer]:so[peer + 1]].contiguous(), peer, group=group))
        if recv_splits[peer]:
            reqs.append(dist.irecv(recv[ro[peer]:ro[peer + 1]], peer, group=group))
    for r in reqs:
        r.wait()


class RowResharder:
    """Re-shards (nx, ny, nsl) scale-block maps (x fastest, Fortran layout: element
    i + nx*(j + ny*s)) into (nx, nrows, ns) row-block maps holding every scale."""

    def __init__(self, nx, ny, ns, world, rank, device, nmaps=4, dtype=torch.float64, group=None, cost=None):
        self.nx, self.ny, self.ns, self.world, self.rank, self.group = nx, ny, ns, world, rank, group
        self.is0, self.is1 = scale_range(ns, world, rank, cost)
        self.nsl = self.is1 - self.is0
        self.row0, row1 = row_range(ny, world, rank)
        self.nrows = row1 - self.row0
        self.rows = [row_range(ny, world, q) for q in range(world)]
        self.scales = [scale_range(ns, world, r, cost) for r in range(world)]
        self.send_splits = [self.nsl * (b - a) * nx for a, b in self.rows]
        self.recv_splits = [(b - a) * self.nrows * nx for a, b in self.scales]
        self.send = torch.empty(max(sum(self.send_splits), 1), dtype=dtype, device=device)
        self.recv = [torch.empty(max(ns * self.nrows * nx, 1), dtype=dtype, device=device) for _ in range(nmaps)]

    def exchange(self, maps):
        """maps: list of flat tensors holding this rank's (nx, ny, nsl) maps -> fills local_maps()."""
        nx, ny, nsl = self.nx, self.ny, self.nsl
        for m, dst in zip(maps, self.recv):
            src = m[:nsl * ny * nx].view(nsl, ny, nx)
            off = 0
            for (a, b), n in zip(self.rows, self.send_splits):  # pack: rows of each destination, scale-major
                if n:
                    self.send[off:off + n].view(nsl, b - a, nx).copy_(src[:, a:b, :])
                off += n
            _all_to_all(dst[:self.ns * self.nrows * nx], self.send[:off], self.recv_splits, self.send_splits,
                        self.group)

    def local_maps(self):
        """Flat tensors of this rank's (nx, nrows, ns) maps (all scales, rows [row0, row0+nrows))."""
        return self.recv


class PeerShards:
    """One-sided form of the exchange: every rank owns an IPC-mappable buffer holding its row shard
    of the four maps for all scales ([map][scale][row][x]); the other ranks map it (CUDA IPC) and
    the admittance / coherence epilogue of ``pfx_wlet_admit_coh_sharded_dev`` stores each output
    row straight into the owner's buffer over NVLink while the transform of the next scale runs.
    No pack kernel, no collective on the data path -- only a barrier before the fit reads."""

    def __init__(self, nx, ny, ns, world, rank, device_index, group=None):
        import ctypes as C

        from . import _lib
        self._lib, self._C = _lib, C
        self.nx, self.ny, self.ns, self.world, self.rank, self.group = nx, ny, ns, world, rank, group
        self.device = device_index
        self.row0, row1 = row_range(ny, world, rank)
        self.nrows = row1 - self.row0
        nbytes = _lib.lib.pfx_shard_bytes(nx, ny, ns, world, rank)
        self.own = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        _lib.check(_lib.lib.pfx_ipc_alloc(device_index, max(nbytes, 16), C.byref(self.own), handle))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.bases = (C.c_void_p * world)()
        self._opened = []
        for q in range(world):
            if q == rank:
                self.bases[q] = self.own.value
                continue
            p = C.c_void_p()
            buf = (C.c_ubyte * 64).from_buffer_copy(handles[q])
            _lib.check(_lib.lib.pfx_ipc_open(device_index, buf, C.byref(p)))
            self.bases[q] = p.value
            self._opened.append(p)

    def transform(self, plan, topo_ptr, grav_ptr, is0, is1):
        """Enqueue the CWT of scales [is0, is1) with the sharded epilogue on the plan's stream."""
        if is1 > is0:
            self._lib.check(self._lib.lib.pfx_wlet_admit_coh_sharded_dev(plan.handle, topo_ptr, grav_ptr, is0, is1,
                                                                         self.world, self.bases))

    def transform_scales(self, plan, topo_ptr, grav_ptr, scales):
        """Same for an arbitrary list of scale indices (deal_scales)."""
        if len(scales):
            arr = (self._C.c_int * len(scales))(*[int(s) for s in scales])
            self._lib.check(self._lib.lib.pfx_wlet_admit_coh_sharded_list_dev(plan.handle, topo_ptr, grav_ptr, arr,
                                                                              len(scales), self.world, self.bases))

    def local_map_ptrs(self):
        """Device pointers of this rank's four (nx, nrows, ns) maps."""
        stride = self.ns * self.nrows * self.nx * 8
        return [self.own.value + m * stride for m in range(4)]

    def close(self):
        for p in self._opened:
            self._lib.lib.pfx_ipc_close(self.device, p)
        self._opened = []
        if self.own:
            self._lib.lib.pfx_ipc_free(self.device, self.own)
            self.own = None


def gather_rows(local, nx_out, out_rows, group=None):
    """All ranks contribute their (nx_out, out_rows[rank]) output map (flat, x fastest); returns the
    concatenation along y on every rank."""
    world = dist.get_world_size(group)
    pieces = [torch.empty(max(nx_out * out_rows[q], 0), dtype=local.dtype, device=local.device) for q in range(world)]
    mx = max(nx_out * r for r in out_rows)
    # all_gather wants equal sizes: pad to the largest shard
    padded = torch.zeros(mx, dtype=local.dtype, device=local.device)
    padded[:local.numel()].copy_(local)
    bufs = [torch.empty(mx, dtype=local.dtype, device=local.device) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    for q in range(world):
        pieces[q] = bufs[q][:nx_out * out_rows[q]]
    return torch.cat(pieces)
