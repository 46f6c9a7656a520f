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

__all__ = ["scale_range", "balanced_cuts", "deal_scales", "row_range", "rows_per_shard", "RowResharder", "PeerShards", "gather_rows",
           "gather_rows_to_root", "stream_barrier", "TeMapPipeline"]


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

    def fence(self):
        """All ranks: everything enqueued so far has completed everywhere.  Needed (a) after the transforms,
        before any rank's fit reads its shard (peers store into it), (b) after the fits, before the next
        transform's remote stores overwrite a shard that is still being read, (c) before ``close`` frees memory
        other ranks have mapped.  ``stream_barrier`` is the stream-ordered form of (a) / (b) that does not
        block the host."""
        import torch
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)

    def close(self):
        if self.world > 1 and dist.is_initialized():
            self.fence()  # no peer may still be storing into (or have work pending on) a buffer about to go
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


class _RawCuda:
    """Zero-copy view of raw device memory for torch.as_tensor (the CUDA array interface, version 3)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def tensor_over(ptr, shape, dtype, device):
    """torch tensor (no copy, no ownership) over device memory at ``ptr`` -- memory of pfx_ipc_alloc here."""
    typestr = {torch.float64: "<f8", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_RawCuda(ptr, shape, typestr), device=device)


class PeerResult:
    """One-sided form of the result gather: rank ``dst`` owns an IPC-mappable buffer holding the (nmaps, my, mx)
    float64 result maps followed by the (my, mx) int32 status words; every rank maps it and its fit kernel stores
    the rows it fitted straight into place over NVLink.  No pack, no gather collective: a stream-ordered barrier
    after the fits tells ``dst`` that everything has landed."""

    def __init__(self, nmaps, my, mx, world, rank, device_index, dst=0, group=None):
        import ctypes as C

        from . import _lib
        self._lib, self._C = _lib, C
        self.nmaps, self.my, self.mx, self.rank, self.dst, self.device = nmaps, my, mx, rank, dst, device_index
        self.world, self.group = world, group
        self.map_bytes = my * mx * 8
        nbytes = nmaps * self.map_bytes + my * mx * 4
        self.own = None
        handle = (C.c_ubyte * 64)()
        if rank == dst:
            self.own = C.c_void_p()
            _lib.check(_lib.lib.pfx_ipc_alloc(device_index, max(nbytes, 16), C.byref(self.own), handle))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self._opened = None
        if rank == dst:
            self.base = self.own.value
        else:
            p = C.c_void_p()
            buf = (C.c_ubyte * 64).from_buffer_copy(handles[dst])
            _lib.check(_lib.lib.pfx_ipc_open(device_index, buf, C.byref(p)))
            self.base, self._opened = p.value, p

    def map_ptr(self, m, out_row0):
        """Where this rank's first output row of map ``m`` lives in the owner's buffer."""
        return self.base + m * self.map_bytes + out_row0 * self.mx * 8

    def status_ptr(self, out_row0):
        return self.base + self.nmaps * self.map_bytes + out_row0 * self.mx * 4

    def tensors(self):
        """On ``dst``: (nmaps, my, mx) float64 maps and (my, mx) int32 status, views of the owned buffer."""
        dev = torch.device("cuda", self.device)
        maps = tensor_over(self.base, (self.nmaps, self.my, self.mx), torch.float64, dev)
        status = tensor_over(self.base + self.nmaps * self.map_bytes, (self.my, self.mx), torch.int32, dev)
        return maps, status

    def close(self):
        if self.world > 1 and dist.is_initialized():
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
        if self._opened is not None:
            self._lib.lib.pfx_ipc_close(self.device, self._opened)
            self._opened = None
        if self.own:
            self._lib.lib.pfx_ipc_free(self.device, self.own)
            self.own = None


def stream_barrier(flag, group=None):
    """Stream-ordered barrier: a one-element all-reduce on the current stream.  Work enqueued after it on any
    rank starts only when every rank's stream has reached it -- without blocking the host, so a pipeline of
    several steps keeps streaming.  ``flag`` is a one-element tensor on the device (CPU tensors with gloo)."""
    dist.all_reduce(flag, group=group)


def gather_rows_to_root(local, nmaps, mx, out_rows, dst=0, group=None):
    """Rank ``dst`` receives the concatenation along y of every rank's (nmaps, out_rows[rank], mx) block
    (flat, x fastest) as one (nmaps, sum(out_rows), mx) tensor; other ranks get None.  One collective for all
    maps; shards are padded to the tallest one because gather wants equal sizes."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    rmax = max(out_rows)
    mine = out_rows[rank]
    padded = local.new_zeros((nmaps, rmax, mx))
    if mine:
        padded[:, :mine, :].copy_(local[:nmaps * mine * mx].view(nmaps, mine, mx))
    bufs = [local.new_empty((nmaps, rmax, mx)) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:, :r, :] for b, r in zip(bufs, out_rows)], dim=1).contiguous()


class TeMapPipeline:
    """The whole hot path on ``world`` GPUs of one node: two (nx, ny) grids on rank 0 in, the fitted
    (Te, std Te, F, std F[, alpha, std alpha], chi2) maps and the per-cell status on rank 0 out
    (Project.wlet_admit_coh + Project.estimate_grid, classes.py:916-975, 1125-1332).

      1. rank 0 broadcasts topography, gravity and water depth (NCCL);
      2. every rank transforms the scales dealt to it by cost (pfx_plan_scale_cost, longest first) and the
         epilogue of each scale stores its output rows straight into the row shard of the GPU that will fit
         them, over NVLink through CUDA-IPC mappings (PeerShards) -- the exchange is fused with the
         epilogue, no pack kernel and no collective on the data path;
      3. a stream-ordered barrier (one-element all-reduce): all shards are complete;
      4. every rank fits the cells of its rows (pfx_l2_fit_rows_dev);
      5. the result maps are gathered on rank 0 (one NCCL gather).
    The broadcast of the next step cannot overtake the gather of this one on rank 0, so no rank starts
    storing into a shard that is still being fitted.  With world == 1 there is nothing to exchange: the
    plain single-GPU entry points run (pfx_wlet_admit_coh_dev + pfx_l2_fit_dev).
    Everything is enqueued on the current CUDA stream of each rank; torch carries buffers and collectives."""

    def __init__(self, nx, ny, dx, dy, k, k0, conf, alph=False, atype="joint", nn=1, local_device=0, engine=None,
                 group=None, balance="deal", solo=False, gather=None):
        import ctypes as C

        import numpy as np

        from . import _lib
        self._lib, self._C = _lib, C
        # solo: this process alone, whatever process group exists (independent replicas, reference runs)
        multi = dist.is_initialized() and not solo
        self.world = dist.get_world_size(group) if multi else 1
        self.rank = dist.get_rank(group) if multi else 0
        self.group, self.local = group, int(local_device)
        self.nx, self.ny, self.ns, self.nn = int(nx), int(ny), len(k), int(nn)
        self.alph, self.atype, self.conf = bool(alph), atype, conf
        self.dev = torch.device("cuda", self.local)
        self.plan = _lib.Plan(nx, ny, dx, dy, k, k0, engine=_lib.default_engine() if engine is None else engine,
                              device=self.local)
        self.d_k = torch.from_numpy(np.ascontiguousarray(k, dtype=np.float64)).to(self.dev)
        self.cost = self.plan.scale_cost()
        nxy = self.nx * self.ny
        self.d_topo = torch.empty(nxy, dtype=torch.float64, device=self.dev)
        self.d_grav = torch.empty(nxy, dtype=torch.float64, device=self.dev)
        self.d_wd = torch.empty(nxy, dtype=torch.float64, device=self.dev)  # (nx, ny) x fastest: [j][i]
        self.nmaps = 7 if self.alph else 5
        self._side, self._wd_event = None, None
        self.mx = self.nx // self.nn
        if self.world > 1:
            self.peers = PeerShards(nx, ny, self.ns, self.world, self.rank, self.local, group=group)
            self.row0, self.nrows = self.peers.row0, self.peers.nrows
            if balance == "deal":
                self.scales = deal_scales(self.cost, self.world)[self.rank]
            else:
                a, b = scale_range(self.ns, self.world, self.rank, self.cost)
                self.scales = list(range(a, b))
            self.map_ptrs = self.peers.local_map_ptrs()
            self.flag = torch.zeros(1, dtype=torch.float32, device=self.dev)
        else:
            self.peers, self.row0, self.nrows = None, 0, self.ny
            self.scales = list(range(self.ns))
            self.maps = torch.empty(4 * self.ns * nxy, dtype=torch.float64, device=self.dev)
            self.map_ptrs = [self.maps.data_ptr() + m * self.ns * nxy * 8 for m in range(4)]
        self.out_rows = []
        for q in range(self.world):
            a, b = row_range(self.ny, self.world, q)
            cj0, myl = C.c_int(), C.c_int()
            _lib.check(_lib.lib.pfx_l2_fit_rows_shape(self.ny, a, b - a, self.nn, C.byref(cj0), C.byref(myl)))
            self.out_rows.append(myl.value)
        self.my = sum(self.out_rows)
        mine = max(self.out_rows[self.rank] * self.mx, 1)
        # fit outputs of this rank, one block: [map][row][x] (+ status as the last "map", stored as float64)
        self.fit_out = torch.zeros((self.nmaps + 1) * mine, dtype=torch.float64, device=self.dev)
        self.fit_status = torch.zeros(mine, dtype=torch.int32, device=self.dev)
        self.mine = mine
        # N > 1: the fit stores its maps straight into rank 0's result buffer (PeerResult); "nccl" keeps the
        # packed gather collective (PLATEFLEX_B200_GATHER, or gather="nccl")
        import os
        self.result = None
        if self.world > 1 and (gather or os.environ.get("PLATEFLEX_B200_GATHER", "p2p")) != "nccl":
            self.result = PeerResult(self.nmaps, self.my, self.mx, self.world, self.rank, self.local, 0, group)
            self.out_row0 = sum(self.out_rows[:self.rank])
            if self.rank == 0:
                self.res_maps, self.res_status = self.result.tensors()
                self.res_all = torch.zeros((self.nmaps + 1, self.my, self.mx), dtype=torch.float64, device=self.dev)

    # ---- steps -------------------------------------------------------------------------------------
    def load(self, topo, grav, wd_f):
        """rank 0: copy the inputs into the pipeline's device buffers (host tensors: pinned memory makes the
        copies asynchronous).  topo / grav are (nx, ny) C order, wd_f is the water depth with x fastest."""
        if self.rank == 0:
            self.d_topo.copy_(topo.reshape(-1), non_blocking=True)
            self.d_grav.copy_(grav.reshape(-1), non_blocking=True)
            # the water depth is needed by the fit only: its copy runs on a side stream behind the transform
            cur = torch.cuda.current_stream(self.dev)
            if self._side is None:
                self._side = torch.cuda.Stream(self.dev)
            self._side.wait_stream(cur)  # the previous step's fit has read the old values
            with torch.cuda.stream(self._side):
                self.d_wd.copy_(wd_f.reshape(-1), non_blocking=True)
            self._wd_event = self._side.record_event()

    def run(self):
        """One pass over the inputs currently in d_topo / d_grav / d_wd of rank 0; returns on rank 0 the
        (nmaps + 1, my, mx) float64 tensor of result maps (the last one the status words), None elsewhere."""
        self.run_transform()
        return self.run_fit()

    def _bind_stream(self):
        stream = torch.cuda.current_stream(self.dev)
        self.plan.set_stream(stream.cuda_stream)
        return stream

    def run_cwt_only(self):
        """This rank's share of the transform alone (no collectives): what the per-kernel timers measure."""
        self._bind_stream()
        if self.world > 1:
            self.peers.transform_scales(self.plan, self.d_topo.data_ptr(), self.d_grav.data_ptr(), self.scales)
        else:
            self._lib.check(self._lib.lib.pfx_wlet_admit_coh_dev(self.plan.handle, self.d_topo.data_ptr(),
                                                                 self.d_grav.data_ptr(), 0, self.ns, *self.map_ptrs))

    def run_transform(self):
        """Steps 1-3: broadcast, scale-sharded transform into the row shards, stream-ordered barrier."""
        if self.world > 1:
            dist.broadcast(self.d_topo, 0, group=self.group)
            dist.broadcast(self.d_grav, 0, group=self.group)
        self.run_cwt_only()
        self._wait_wd()
        if self.world > 1:
            dist.broadcast(self.d_wd, 0, group=self.group)  # needed by the fit only: behind the transform
            stream_barrier(self.flag, self.group)

    def _wait_wd(self):
        if self._wd_event is not None:
            torch.cuda.current_stream(self.dev).wait_event(self._wd_event)
            self._wd_event = None

    def run_fit(self):
        """Steps 4-5: fit of this rank's rows, gather on rank 0."""
        lib, C = self._lib.lib, self._C
        stream = self._bind_stream()
        self._wait_wd()
        p2p = self.result is not None
        if self.nrows > 0 and self.out_rows[self.rank] > 0:
            if p2p:  # every output row at its place in rank 0's buffer
                o = [self.result.map_ptr(i, self.out_row0) for i in range(self.nmaps)]
                status_ptr = self.result.status_ptr(self.out_row0)
            else:
                o = [self.fit_out.data_ptr() + i * self.mine * 8 for i in range(self.nmaps)]
                status_ptr = self.fit_status.data_ptr()
            te, te_s, f, f_s = o[0], o[1], o[2], o[3]
            al, al_s, chi2 = (o[4], o[5], o[6]) if self.alph else (None, None, o[4])
            wd = self.d_wd.data_ptr() + self.row0 * self.nx * 8
            self._lib.check(lib.pfx_l2_fit_rows_dev(
                self.local, stream.cuda_stream, C.byref(self.conf), self.d_k.data_ptr(), self.ns, self.nx, self.ny,
                self.row0, self.nrows, *self.map_ptrs, wd, None, None, None, self.nn, int(self.alph),
                self._lib.ATYPES[self.atype], te, te_s, f, f_s, al, al_s, chi2, status_ptr))
            if not p2p:
                self.fit_out[self.nmaps * self.mine:].copy_(self.fit_status)  # int32 -> float64, exact
        if p2p:
            stream_barrier(self.flag, self.group)  # every rank's rows have landed in rank 0's buffer
            if self.rank != 0:
                return None
            self.res_all[:self.nmaps].copy_(self.res_maps)
            self.res_all[self.nmaps].copy_(self.res_status)  # int32 -> float64, exact
            return self.res_all
        if self.world > 1:
            return gather_rows_to_root(self.fit_out, self.nmaps + 1, self.mx, self.out_rows, 0, self.group)
        return self.fit_out.view(self.nmaps + 1, self.out_rows[0], self.mx)

    def launches(self):
        return self.plan.launch_count()

    def close(self):
        if getattr(self, "result", None) is not None:
            self.res_maps = self.res_status = None
            self.result.close()
            self.result = None
        if self.peers is not None:
            self.peers.close()
            self.peers = None
        self.plan.close()
