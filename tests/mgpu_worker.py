"""torchrun worker for tests/test_gpu_stripes.py::test_real_multi_gpu: one process per GPU,
stripes wired through CUDA IPC; checks the striped cascade sweep bitwise against the oracle and
the striped parallel erode for particle/mass accounting."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import simplewinderosion_b200 as swe  # noqa: E402
from simplewinderosion_b200 import generate, stripes  # noqa: E402
from tests.oracle_lib import PortWorld  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, cycles = 512, 65536
    hm = np.random.RandomState(3).random_sample((n, n)).astype(np.float32)
    w = stripes.StripedWorld(n, n, rank, world, cycles, device=local)
    w.connect()
    w.upload_global(hm)
    w.sync_halos()
    w.cascade_sweep(3)
    ora = PortWorld(n)
    ora.set_plane(0, hm)
    ora.cascade_sweep(3)
    lo, hi = w.bounds[rank]
    got = w.download()
    assert np.array_equal(got.view(np.uint32), np.array(ora.plane(0))[lo:hi].view(np.uint32)), "striped sweep != oracle"

    # generate on the stripes: noise layers in global coordinates, collective normalisation; bitwise
    # equal to the oracle's whole-map generator, ghost rows included (a sweep follows without a sync)
    w.upload_global(np.zeros((n, n), np.float32))
    w.sync_halos()
    w.generate_noise(9, 8)
    rng = w.normalize_collective()
    og = PortWorld(n)
    og.noise_stack(9, 8)
    raw = np.array(og.plane(0))
    assert rng == (min(0.0, float(raw.min())), max(0.0, float(raw.max()))), rng
    og.normalize()
    assert np.array_equal(w.download().view(np.uint32), np.array(og.plane(0))[lo:hi].view(np.uint32)), "striped generate != oracle"
    w.cascade_sweep(1); og.cascade_sweep(1)
    assert np.array_equal(w.download().view(np.uint32), np.array(og.plane(0))[lo:hi].view(np.uint32)), "sweep after generate != oracle"

    # parallel erode on structured terrain (the noise-layer dune field generated above, relaxed by one sweep)
    hm2_local = w.download()
    assert np.array_equal(hm2_local.view(np.uint32), np.array(og.plane(0))[lo:hi].view(np.uint32))
    hm2 = np.array(og.plane(0))                      # (bitwise what the stripes hold, asserted above)
    w.srand(5)
    w.erode(cycles, swe.MODE_PARALLEL)
    st = w.stats()
    t = torch.tensor([st.particles, st.exits_oob + st.exits_age, st.steps, st.stripe_overflow + st.queue_dropped], dtype=torch.float64, device="cuda")
    m = torch.tensor([w.sum_plane(), st.sediment_lost], dtype=torch.float64, device="cuda")
    dist.all_reduce(t); dist.all_reduce(m)
    s0 = float(hm2.astype(np.float64).sum())
    assert int(t[0]) == cycles and int(t[1]) == cycles and int(t[3]) == 0, t.tolist()
    assert abs(float(m[0] + m[1]) - s0) < 2e-6 * s0 + 1e-4, (m.tolist(), s0)
    # stripes of 512 / world rows are at least two 60-row tiles tall: the call ran the tile-resident schedule (global
    # tiling, boundary tiles working in the neighbour's memory, three cross-GPU barriers per super-round)
    assert st.last_super_rounds > 0, "the tile-resident schedule did not run on the stripes"
    # against a single engine on the same schedule (same seed) and, as the yardstick, one with another seed
    single, other = swe.World(n, device=local), swe.World(n, device=local)
    single.upload(hm2); single.srand(5); single.erode(cycles, swe.MODE_PARALLEL)
    other.upload(hm2); other.srand(6); other.erode(cycles, swe.MODE_PARALLEL)
    ref, oth = single.download()[lo:hi] - hm2[lo:hi], other.download()[lo:hi] - hm2[lo:hi]
    mine = w.download() - hm2[lo:hi]
    corr = float(np.corrcoef(mine.ravel(), ref.ravel())[0, 1])
    base = float(np.corrcoef(oth.ravel(), ref.ravel())[0, 1])
    fx, fr = w.download(swe.MOMENTUMX), single.download(swe.MOMENTUMX)[lo:hi]
    fratio = float(fx.mean()) / float(fr.mean())
    seam = slice(0, 4) if rank > 0 else slice(hi - lo - 4, hi - lo)          # the rows next to a stripe boundary
    seam_corr = float(np.corrcoef(mine[seam].ravel(), ref[seam].ravel())[0, 1])
    seam_base = float(np.corrcoef(oth[seam].ravel(), ref[seam].ravel())[0, 1])
    print(f"rank {rank}: field corr vs single {corr:.4f} (two seeds of a single engine: {base:.4f}); seam rows {seam_corr:.4f} ({seam_base:.4f}); "
          f"momentumx mean ratio {fratio:.4f}; steps {st.steps}", flush=True)
    assert corr > base - 0.05, (corr, base)
    assert seam_corr > seam_base - 0.15, (seam_corr, seam_base)
    assert abs(fratio - 1.0) < 0.1, fratio
    steps_single = single.stats().steps
    single.close(); other.close()
    if rank == 0:
        assert abs(int(t[2]) / steps_single - 1.0) < 0.02, (int(t[2]), steps_single)
        print(f"MGPU_OK world={world} steps={int(t[2])} (single engine {steps_single}) super_rounds={st.last_super_rounds} "
              f"mass_drift={float(m[0] + m[1]) - s0:.3e}")
    dist.barrier()
    w.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
