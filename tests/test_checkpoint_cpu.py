"""CPU: the checkpoint file format (simplewinderosion_b200/checkpoint.py) -- header parsing,
plane offsets and error handling -- on a file written by hand (no GPU needed to read one)."""
import json
import struct

import numpy as np
import pytest

from simplewinderosion_b200 import checkpoint


def write_fake(path, nx=5, nz=7, rock=False):
    names = list(checkpoint.PLANE_NAMES) + (["rock"] if rock else [])
    header = {"nx": nx, "nz": nz, "row0": 0, "global_nx": nx, "planes": names, "params": {"max_age": 512},
              "state": "00" * 256, "has_rock": rock, "stats": {}, "user": {"note": "x"}}
    blob = json.dumps(header).encode()
    planes = [np.full((nx, nz), i + 0.5, np.float32) + np.arange(nz, dtype=np.float32) for i in range(len(names))]
    with open(path, "wb") as f:
        f.write(checkpoint.MAGIC + struct.pack("<II", checkpoint.VERSION, len(blob)) + blob)
        f.write(b"\0" * (-(16 + len(blob)) % 64))
        for p in planes:
            f.write(p.tobytes())
    return header, planes


@pytest.mark.parametrize("rock", [False, True])
def test_read_planes_roundtrip(tmp_path, rock):
    path = str(tmp_path / "c.swe")
    header, planes = write_fake(path, rock=rock)
    h2, got = checkpoint.read_planes(path)
    assert h2 == header
    _, off = checkpoint.read_header(path)
    assert off % 64 == 0
    assert list(got) == header["planes"]
    for name, want in zip(header["planes"], planes):
        assert np.array_equal(np.asarray(got[name]), want)


def test_rejects_foreign_files(tmp_path):
    p = tmp_path / "bad"
    p.write_bytes(b"not a checkpoint at all")
    with pytest.raises(ValueError):
        checkpoint.read_header(str(p))
    q = tmp_path / "v9"
    q.write_bytes(checkpoint.MAGIC + struct.pack("<II", 9, 2) + b"{}")
    with pytest.raises(ValueError):
        checkpoint.read_header(str(q))
