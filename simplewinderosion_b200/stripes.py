"""Multi-GPU row stripes: one process and one engine per GPU, the map cut along x
(the slow index). torch.distributed is used ONLY for set-up plumbing (exchanging CUDA IPC handles,
reducing statistics); inside World.erode the neighbours exchange ghost rows and hand over particles
with kernels over peer-mapped memory (NVLink P2P), see csrc/swe_stripe.cu.

    bounds = stripe_bounds(nx, world_size)
    w = StripedWorld(nx, nz, rank, world_size, max_particles, device=local_rank)
    w.connect()                      # all ranks (collective)
    w.upload_global(height)          # every rank passes the same global map (or only its rows)
    w.erode(cycles)                  # cycles = GLOBAL particle count of the call
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import world as _world
from ._capi import STRIPE_IPC_BYTES, StripeLink, check, lib

MIN_ROWS = 8


def stripe_bounds(nx: int, world_size: int, align: int = 4) -> list[tuple[int, int]]:
    """Balanced row ranges, boundaries on multiples of `align` (the spawn order works in strips of
    4 rows), every stripe at least MIN_ROWS rows."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    if nx < MIN_ROWS * world_size:
        raise ValueError(f"{nx} rows cannot be cut into {world_size} stripes of >= {MIN_ROWS} rows")
    cuts = [0]
    for r in range(1, world_size):
        c = int(round(nx * r / world_size / align)) * align
        c = max(c, cuts[-1] + MIN_ROWS)
        cuts.append(c)
    cuts.append(nx)
    for r in range(world_size - 1, 0, -1):          # keep the tail stripes tall enough
        if cuts[r + 1] - cuts[r] < MIN_ROWS:
            cuts[r] = cuts[r + 1] - MIN_ROWS
    out = [(cuts[r], cuts[r + 1]) for r in range(world_size)]
    assert all(hi - lo >= MIN_ROWS for lo, hi in out) and out[0][0] == 0 and out[-1][1] == nx
    return out


def owner_of_row(bounds: list[tuple[int, int]], x: int) -> int:
    for r, (lo, hi) in enumerate(bounds):
        if lo <= x < hi:
            return r
    raise ValueError(f"row {x} outside the map")


def allgather_bytes(blob: bytes, world_size: int, group=None, device=None) -> list[bytes]:
    """Fixed-size byte blobs from every rank (torch.distributed plumbing; works with gloo or nccl)."""
    import torch
    import torch.distributed as dist
    backend = dist.get_backend(group)
    dev = torch.device("cuda", device or 0) if backend == "nccl" else torch.device("cpu")
    mine = torch.tensor(list(blob), dtype=torch.uint8, device=dev)
    gathered = [torch.empty(len(blob), dtype=torch.uint8, device=dev) for _ in range(world_size)]
    dist.all_gather(gathered, mine, group=group)
    return [bytes(g.cpu().tolist()) for g in gathered]


def reduce_range(lo: float, hi: float, group=None, device=None) -> tuple[float, float]:
    """Global (min, max) of per-rank height ranges (torch.distributed plumbing; gloo or nccl): what the
    normalisation of map::init (cellpool.h:416-428) needs when the map is cut into stripes."""
    import torch
    import torch.distributed as dist
    backend = dist.get_backend(group)
    dev = torch.device("cuda", device or 0) if backend == "nccl" else torch.device("cpu")
    t = torch.tensor([-lo, hi], dtype=torch.float32, device=dev)     # one MAX reduction for both ends
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return -float(t[0].item()), float(t[1].item())


class StripeGroup:
    """Several stripes of one map driven from THIS host thread (logical stripes on one GPU, or a
    process owning several GPUs). The C library interleaves the launches exchange by exchange."""

    def __init__(self, worlds: list):
        self.worlds = list(worlds)
        self._arr = (C.c_void_p * len(worlds))(*[w._h for w in worlds])

    def sync_halos(self) -> None:
        check(lib.swe_stripe_group_sync_halos(self._arr, len(self.worlds)))

    def erode(self, cycles: int) -> None:
        check(lib.swe_stripe_group_erode(self._arr, len(self.worlds), int(cycles)))

    def cascade_sweep(self, sweeps: int = 1) -> None:
        check(lib.swe_stripe_group_sweep(self._arr, len(self.worlds), int(sweeps)))

    def max_excess(self) -> float:
        self.sync_halos()
        return max(w.max_excess_local() for w in self.worlds)

    def gather(self, plane: int = _world.HEIGHT) -> np.ndarray:
        return np.concatenate([w.download(plane) for w in self.worlds], axis=0)

    def stats(self):
        return [w.stats() for w in self.worlds]


class StripedWorld(_world.World):
    """One rank's stripe of a World spread over world_size GPUs."""

    def __init__(self, nx: int, nz: int, rank: int, world_size: int, max_particles: int,
                 params=None, device: int = 0):
        self.rank, self.world_size = int(rank), int(world_size)
        self.bounds = stripe_bounds(nx, world_size)
        super().__init__(nx, nz, params=params, device=device, stripe=self.bounds[rank],
                         max_particles=max_particles)

    # -- wiring -----------------------------------------------------------------------------------
    def export_link(self) -> StripeLink:
        link = StripeLink()
        check(lib.swe_stripe_export(self._h, C.byref(link)))
        return link

    def connect_links(self, below: StripeLink | None, above: StripeLink | None) -> None:
        check(lib.swe_stripe_connect(self._h, C.byref(below) if below is not None else None,
                                     C.byref(above) if above is not None else None))

    def connect_ring(self, links: list[StripeLink], colocated: int = 1) -> None:
        """Neighbour wiring plus every rank's barrier mailbox: enables the colour-queue schedule on stripes."""
        arr = (StripeLink * len(links))(*links)
        check(lib.swe_stripe_connect_ring(self._h, self.rank, len(links), arr, int(colocated)))

    @staticmethod
    def connect_in_process(worlds: list["StripedWorld"], ring: bool = True) -> None:
        """All stripes live in this process (several logical stripes on one or more GPUs)."""
        links = [w.export_link() for w in worlds]
        for r, w in enumerate(worlds):
            if ring and len(worlds) <= 8:
                w.connect_ring(links, colocated=sum(1 for o in worlds if o.device == w.device))
            else:
                w.connect_links(links[r - 1] if r > 0 else None, links[r + 1] if r + 1 < len(worlds) else None)

    def connect(self, group=None) -> None:
        """One process per GPU: exchange CUDA IPC handles through torch.distributed (collective)."""
        import torch.distributed as dist
        blob = (C.c_ubyte * STRIPE_IPC_BYTES)()
        check(lib.swe_stripe_export_ipc(self._h, blob, STRIPE_IPC_BYTES))
        gathered = allgather_bytes(bytes(blob), self.world_size, group, self.device)
        if self.world_size <= 8:
            # every rank's buffers are mapped: the neighbours' for halos / hand-off / peer block access, everybody's
            # barrier mailbox for the cross-GPU grid barrier of the colour-queue kernels
            links = []
            for r in range(self.world_size):
                link = StripeLink()
                if r != self.rank:
                    buf = (C.c_ubyte * STRIPE_IPC_BYTES).from_buffer_copy(gathered[r])
                    check(lib.swe_stripe_open_ipc(self._h, buf, STRIPE_IPC_BYTES, C.byref(link)))
                links.append(link)
            self.connect_ring(links)
        else:
            links = {}
            for nb in (self.rank - 1, self.rank + 1):
                if 0 <= nb < self.world_size:
                    buf = (C.c_ubyte * STRIPE_IPC_BYTES).from_buffer_copy(gathered[nb])
                    link = StripeLink()
                    check(lib.swe_stripe_open_ipc(self._h, buf, STRIPE_IPC_BYTES, C.byref(link)))
                    links[nb] = link
            self.connect_links(links.get(self.rank - 1), links.get(self.rank + 1))
        dist.barrier(group=group)

    # -- map transfer -------------------------------------------------------------------------------
    def upload_global(self, height: np.ndarray, plane: int = _world.HEIGHT) -> None:
        """Take this rank's rows out of a global (nx, nz) plane. Call sync_halos() afterwards
        (collective: every rank must)."""
        lo, hi = self.bounds[self.rank]
        self.upload(np.ascontiguousarray(height[lo:hi]), plane)

    def upload_rock_global(self, rock: np.ndarray) -> None:
        """Layered extension on a stripe: take this rank's STORED rows (owned + ghost rows) out of the global
        (nx, nz) bedrock plane."""
        x0, rows = C.c_int(), C.c_int()
        check(lib.swe_stripe_stored_rows(self._h, C.byref(x0), C.byref(rows)))
        a = np.ascontiguousarray(rock[x0.value:x0.value + rows.value], dtype=np.float32)
        check(lib.swe_upload_rock(self._h, a.ctypes.data, 4))

    def sync_halos(self) -> None:
        check(lib.swe_stripe_sync_halos(self._h))

    def max_excess(self) -> float:
        """Collective (one process per GPU): refresh the ghost rows, then the local maximum over
        the owned rows and the pairs into the ghost row above. Reduce with MAX over ranks."""
        self.sync_halos()
        return super().max_excess()

    def max_excess_local(self) -> float:
        return super().max_excess()

    def normalize_collective(self, group=None) -> tuple[float, float]:
        """map::init's normalisation (cellpool.h:416-428) on a striped map: every rank folds its owned
        rows, the ranges are reduced over the ranks, every rank rescales its rows and ghost rows with
        the global range (collective: every rank must call it). Returns the global (min, max)."""
        lo, hi = reduce_range(*self.height_range(), group=group, device=self.device)
        self.normalize(lo, hi)
        return lo, hi


def gather_global(worlds: list[StripedWorld], plane: int = _world.HEIGHT) -> np.ndarray:
    """In-process helper: stitch the owned rows of all stripes back into one (nx, nz) plane."""
    return np.concatenate([w.download(plane) for w in worlds], axis=0)
