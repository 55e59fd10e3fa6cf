"""GPU: the renderer read-back of SURVEY 8f rank 3 -- swe_download_vertices / swe_write_vertices_device
produce the Vertex records of quad::updatenode (cellpool.h:302-321). Bar: BIT-EXACT against the
records the reference's own updatenode wrote (tests/golden/vertices.npz) and against the oracle."""
import os

import numpy as np
import pytest

import simplewinderosion_b200 as swe
from simplewinderosion_b200 import stripes
from tests.oracle_lib import PortWorld

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vertices.npz"))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("case", ["eroded", "random"])
def test_vertices_match_reference_updatenode_golden(case):
    w = swe.World(32)
    w.upload(G[f"{case}_height"])
    assert np.array_equal(bits(w.vertices()), bits(G[f"{case}_vertices"]))


@pytest.mark.parametrize("nx,nz", [(8, 8), (97, 40), (256, 512)])
def test_vertices_match_oracle_after_deterministic_erosion(nx, nz):
    ora = PortWorld(nx, nz)
    rs = np.random.RandomState(nx)
    hm = (0.2 * rs.random_sample((nx, nz))).astype(np.float32)
    ora.set_plane(0, hm)
    w = swe.World(nx, nz)
    w.upload(hm)
    w.srand(4); ora.srand(4)
    w.erode(64, swe.MODE_DETERMINISTIC); ora.erode(64)
    v = w.vertices()
    assert np.array_equal(bits(v), bits(ora.vertices()))
    assert np.array_equal(bits(v[..., 3:6]), bits(w.normals()))      # same normals as swe_download_normals


def test_device_side_write_equals_host_download():
    import torch
    n = 512
    w = swe.World(n)
    w.generate_noise(2, 8); w.normalize()
    w.erode(50000, swe.MODE_PARALLEL)
    vbo = torch.empty((n, n, 12), dtype=torch.float32, device="cuda")   # stands in for the mapped VBO
    w.write_vertices_device(vbo.data_ptr())
    w.sync()
    assert np.array_equal(bits(vbo.cpu().numpy()), bits(w.vertices()))
    with pytest.raises(swe.SweError):
        w.write_vertices_device(vbo.data_ptr() + 4)                     # must be 16-byte aligned


def test_stripes_write_their_rows():
    nx, nz, world = 96, 64, 3
    hm = np.random.RandomState(3).random_sample((nx, nz)).astype(np.float32)
    single = swe.World(nx, nz)
    single.upload(hm)
    ws = [stripes.StripedWorld(nx, nz, r, world, 1024) for r in range(world)]
    stripes.StripedWorld.connect_in_process(ws)
    for w in ws:
        w.upload_global(hm)
    stripes.StripeGroup(ws).sync_halos()
    got = np.concatenate([w.vertices() for w in ws], axis=0)
    assert np.array_equal(bits(got), bits(single.vertices()))
