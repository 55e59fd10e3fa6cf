"""CPU: host-side logic of the multi-GPU stripe decomposition -- partitioning, and the
torch.distributed plumbing (IPC-blob all_gather) under gloo with world_size 2."""
import os
import socket

import pytest

from simplewinderosion_b200 import stripes


def test_stripe_bounds_properties():
    for nx, w in [(16384, 8), (4096, 4), (64, 2), (100, 3), (16, 2), (4096, 1), (70, 8), (1024, 7)]:
        b = stripes.stripe_bounds(nx, w)
        assert len(b) == w and b[0][0] == 0 and b[-1][1] == nx
        for (lo, hi), (lo2, _) in zip(b, b[1:]):
            assert hi == lo2
        assert all(hi - lo >= stripes.MIN_ROWS for lo, hi in b)
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 12
        for x in (0, nx // 3, nx - 1):
            r = stripes.owner_of_row(b, x)
            assert b[r][0] <= x < b[r][1]
    assert stripes.stripe_bounds(16384, 8)[3] == (6144, 8192)
    with pytest.raises(ValueError):
        stripes.stripe_bounds(15, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b = stripes.stripe_bounds(4096, world)
        blob = bytes([rank + 1]) * stripes.STRIPE_IPC_BYTES
        got = stripes.allgather_bytes(blob, world)
        ok = all(g == bytes([r + 1]) * stripes.STRIPE_IPC_BYTES for r, g in enumerate(got))
        # every rank derives the same partition and knows its neighbours
        nbrs = [r for r in (rank - 1, rank + 1) if 0 <= r < world]
        # normalisation of a striped map: per-rank (min, max) -> the global range, exactly (fp32 values)
        mine = [(-0.5191119, 0.25), (-0.125, 0.3582528)][rank]
        rng = stripes.reduce_range(*mine)
        ok = ok and rng == (float(__import__("numpy").float32(-0.5191119)), float(__import__("numpy").float32(0.3582528)))
        q.put((rank, ok, b, nbrs))
    finally:
        dist.destroy_process_group()


def test_ipc_blob_exchange_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] and res[1][1]
    assert res[0][2] == res[1][2] == [(0, 2048), (2048, 4096)]
    assert res[0][3] == [1] and res[1][3] == [0]


def test_tile_ownership_partitions_the_tile_rows():
    """Tile-resident schedule on stripes (host arithmetic, no GPU): for every compiled tile geometry, each of the
    three shifted tilings and a range of stripe cuts, the ranks' tile-row ranges partition [0, ntx) in rank order, and
    every tile row's centre row (clamped to the map) lies in its owner's rows."""
    import ctypes as C

    from simplewinderosion_b200 import stripes
    from simplewinderosion_b200._capi import check, lib

    def geometry(shape):
        ex, ez, nt = C.c_int(), C.c_int(), C.c_int()
        check(lib.swe_tile_geometry(shape, C.byref(ex), C.byref(ez), C.byref(nt)))
        return ex.value, ez.value, nt.value

    shapes = []
    for shape in range(8):
        if lib.swe_tile_geometry(shape, None, None, None) != 0:
            break
        shapes.append(shape)
    assert len(shapes) >= 2 and geometry(0)[:2] == (60, 60)
    for shape in shapes:
        ex, ez, nt = geometry(shape)
        assert ex % 3 == 0 and ez % 3 == 0 and nt % 32 == 0 and ex * ez <= 14400
        for nx, world in ((2 * ex + 8, 1), (4 * ex + 40, 2), (4096, 4), (4096, 8), (16384, 8), (3 * ex * 3 + 17, 3)):
            if nx // world - 8 < 2 * ex:
                continue
            bounds = stripes.stripe_bounds(nx, world)
            for tiling in range(3):
                off = tiling * (ex // 3)
                prev_end, total = 0, None
                for lo, hi in bounds:
                    a, b, n = C.c_int(), C.c_int(), C.c_int()
                    check(lib.swe_tile_owned_rows(nx, shape, tiling, lo, hi, C.byref(a), C.byref(b), C.byref(n)))
                    assert a.value == prev_end and b.value >= a.value + 1, (shape, nx, world, tiling, lo, hi, a.value, b.value)
                    total = n.value
                    assert total == (nx + off + ex - 1) // ex
                    for row in range(a.value, b.value):
                        centre = min(max(row * ex - off + ex // 2, 0), nx - 1)
                        assert lo <= centre < hi, (shape, nx, world, tiling, row, centre, lo, hi)
                        # the tile reaches at most into the two neighbouring stripes
                        top, bottom = max(row * ex - off, 0), min((row + 1) * ex - off, nx)
                        r = bounds.index((lo, hi))
                        assert top >= (bounds[r - 1][0] if r > 0 else 0) and bottom <= (bounds[r + 1][1] if r + 1 < world else nx)
                    prev_end = b.value
                assert prev_end == total
    assert lib.swe_tile_owned_rows(100, 0, 0, 50, 40, None, None, None) != 0      # bad arguments are refused
