"""CPU: the oracle's restatement of quad::updatenode (cellpool.h:302-321) against the records the
reference's own updatenode produced (tests/golden/vertices.npz) and, where oracle/_ref is present,
against the reference directly. Bit-exact."""
import os

import numpy as np
import pytest

from tests.oracle_lib import PortWorld, RefWorld, ref_available

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vertices.npz"))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("case", ["eroded", "random"])
def test_vertices_match_reference_updatenode(case):
    w = PortWorld(32)
    w.set_plane(0, G[f"{case}_height"])
    got, want = w.vertices(), G[f"{case}_vertices"]
    assert np.array_equal(bits(got), bits(want))
    # structure: position = (x, 80 h, z); tangent/bitangent span one cell; edge neighbours have height 0
    h = G[f"{case}_height"]
    assert np.array_equal(want[..., 0], np.arange(32, dtype=np.float32)[:, None].repeat(32, 1))
    assert np.array_equal(want[..., 1], np.float32(80) * h)
    assert np.all(want[..., 6] == 1) and np.all(want[..., 8] == 0) and np.all(want[..., 9] == 0) and np.all(want[..., 11] == 1)
    assert np.array_equal(want[31, :, 7], -want[31, :, 1]) and np.array_equal(want[:, 31, 10], -want[:, 31, 1])
    assert np.allclose(np.linalg.norm(want[..., 3:6], axis=-1), 1.0, atol=1e-5)


@pytest.mark.skipif(not ref_available(64), reason="oracle/_ref not built")
def test_port_equals_reference_updatenode_after_erosion():
    r = RefWorld(64, seed=11)
    if not hasattr(r.L, "ref_updatenode"):
        pytest.skip("oracle/_ref predates the updatenode harness")
    w = PortWorld(64)
    w.init_gaussian()
    w.srand(11)
    r.erode(256); w.erode(256)
    assert np.array_equal(bits(r.vertices()), bits(w.vertices()))
