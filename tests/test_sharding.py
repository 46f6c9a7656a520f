"""Host-side multi-GPU logic on CPU: world_size-2 (and 3) gloo process groups exercise the scale /
row partitioning and the all-to-all re-shard of plateflex_b200/sharding.py; the row-sharded fit's
output-row bookkeeping (pfx_l2_fit_rows_shape, no device work) is checked against the reference's
loop bounds (classes.py:1189, 1218-1219)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nx, ny, ns, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from plateflex_b200 import sharding
        rng = np.random.default_rng(7)
        full = [rng.standard_normal((ns, ny, nx)) for _ in range(4)]  # [s][j][i] == Fortran (nx, ny, ns)
        rs = sharding.RowResharder(nx, ny, ns, world, rank, torch.device("cpu"))
        is0, is1 = sharding.scale_range(ns, world, rank)
        local = [torch.from_numpy(np.ascontiguousarray(f[is0:is1]).reshape(-1)) for f in full]
        rs.exchange(local)
        ok = True
        for f, got in zip(full, rs.local_maps()):
            want = f[:, rs.row0:rs.row0 + rs.nrows, :].reshape(-1)
            ok = ok and np.array_equal(got[:want.size].numpy(), want)
        # gather of per-shard output maps back into the full map
        out_rows = [b - a for a, b in (sharding.row_range(ny, world, r) for r in range(world))]
        mine = torch.from_numpy(np.ascontiguousarray(full[0][0, rs.row0:rs.row0 + rs.nrows, :]).reshape(-1))
        whole = sharding.gather_rows(mine, nx, out_rows)
        ok = ok and np.array_equal(whole.numpy(), full[0][0].reshape(-1))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nx,ny,ns", [(2, 13, 10, 5), (2, 8, 7, 20), (3, 5, 4, 2)])
def test_reshard_scale_blocks_to_row_blocks_gloo(world, nx, ny, ns):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nx, ny, ns, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _ in res) == list(range(world)) and all(ok for _, ok in res)


def _gather_worker(rank, world, port, nmaps, mx, out_rows, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from plateflex_b200 import sharding
        rng = np.random.default_rng(11)
        full = rng.standard_normal((nmaps, sum(out_rows), mx))
        lo = sum(out_rows[:rank])
        # the rank's block lives in a buffer that may be larger than what it owns (at least one element)
        mine = np.zeros(max(nmaps * out_rows[rank] * mx, 1) + 3)
        mine[:nmaps * out_rows[rank] * mx] = full[:, lo:lo + out_rows[rank], :].reshape(-1)
        flag = torch.ones(1)
        sharding.stream_barrier(flag)  # the stream-ordered barrier is a one-element all-reduce
        got = sharding.gather_rows_to_root(torch.from_numpy(mine), nmaps, mx, out_rows, dst=0)
        ok = float(flag.item()) == world
        if rank == 0:
            ok = ok and got.shape == full.shape and np.array_equal(got.numpy(), full)
        else:
            ok = ok and got is None
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nmaps,mx,out_rows", [(2, 6, 7, [4, 3]), (3, 8, 5, [2, 2, 0]), (2, 6, 1, [1, 1])])
def test_gather_rows_to_root_gloo(world, nmaps, mx, out_rows):
    """Result maps of the row shards (unequal, possibly empty) concatenated along y on rank 0 with one gather."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, nmaps, mx, out_rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _ in res) == list(range(world)) and all(ok for _, ok in res)


def test_partitions_cover_everything():
    from plateflex_b200 import sharding
    for n, world in [(20, 8), (20, 3), (3, 8), (1024, 8), (341, 4)]:
        for fn in (sharding.scale_range, sharding.row_range):
            blocks = [fn(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
    assert [b - a for a, b in (sharding.scale_range(20, 8, r) for r in range(8))] == [3, 3, 3, 3, 2, 2, 2, 2]


@pytest.mark.parametrize("ny,nn,world", [(341, 10, 4), (1024, 1, 8), (50, 7, 3), (9, 5, 4), (64, 64, 2)])
def test_row_sharded_fit_bookkeeping_matches_reference_loop(pf, ny, nn, world):
    """Shard outputs concatenated along y must be exactly the (ny // nn) output rows of
    classes.py:1189, each output row cj mapping to grid row cj*nn inside its shard."""
    from plateflex_b200 import _lib, sharding
    rows_seen = []
    for r in range(world):
        row0, row1 = sharding.row_range(ny, world, r)
        cj0, n = C.c_int(), C.c_int()
        _lib.check(_lib.lib.pfx_l2_fit_rows_shape(ny, row0, row1 - row0, nn, C.byref(cj0), C.byref(n)))
        for cj in range(cj0.value, cj0.value + n.value):
            assert row0 <= cj * nn < row1
            rows_seen.append(cj)
    assert rows_seen == list(range(ny // nn))
    # every fitted row of the reference loop lands in exactly one shard
    assert set(range(0, ny - nn, nn)) <= {cj * nn for cj in rows_seen}


def test_balanced_cuts_minimise_the_heaviest_block():
    from plateflex_b200 import sharding
    cost = [1, 1, 1, 1, 2, 2, 4, 4, 8, 9]
    cuts = sharding.balanced_cuts(cost, 4)
    assert cuts[0] == 0 and cuts[-1] == len(cost) and cuts == sorted(cuts)
    blocks = [sum(cost[a:b]) for a, b in zip(cuts, cuts[1:])]
    assert max(blocks) == 9  # [1,1,1,1,2,2] [4,4] [8] [9]
    assert sharding.balanced_cuts([5.0], 3)[-1] == 1
    # every rank sees the same cuts; ranges tile the axis
    r = [sharding.scale_range(len(cost), 4, q, cost) for q in range(4)]
    assert r[0][0] == 0 and r[-1][1] == len(cost) and all(a[1] == b[0] for a, b in zip(r, r[1:]))


def test_deal_scales_balances_better_than_blocks():
    from plateflex_b200 import sharding
    cost = [2.8, 2.8, 2.9, 3.0, 3.1, 3.3, 3.6, 3.9, 4.1, 4.3, 4.6, 4.9, 5.2, 5.5, 5.8, 6.2, 6.6, 7.0, 7.3, 7.6]
    world = 8
    dealt = sharding.deal_scales(cost, world)
    assert sorted(s for part in dealt for s in part) == list(range(len(cost)))
    assert all(part == sorted(part) for part in dealt)
    cuts = sharding.balanced_cuts(cost, world)
    worst_block = max(sum(cost[a:b]) for a, b in zip(cuts, cuts[1:]))
    worst_dealt = max(sum(cost[s] for s in part) for part in dealt)
    assert worst_dealt <= worst_block
    assert worst_dealt <= 1.15 * sum(cost) / world
