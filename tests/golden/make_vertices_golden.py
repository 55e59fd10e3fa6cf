"""Generates tests/golden/vertices.npz: the Vertex records the reference's own quad::updatenode
(cellpool.h:302-321) writes, captured through the recording Vertexpool stub of oracle/_ref, for an
eroded Gaussian hill and a random map (32^2).  Run from the repo root:
    python tests/golden/make_vertices_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.oracle_lib import RefWorld  # noqa: E402


def main():
    r = RefWorld(32, seed=5)
    r.erode(200)
    out = {"eroded_height": r.plane(0), "eroded_vertices": r.vertices()}
    hm = np.random.RandomState(9).uniform(-1, 1, (32, 32)).astype(np.float32)
    r.set_plane(0, hm)
    out["random_height"] = hm
    out["random_vertices"] = r.vertices()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "vertices.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
