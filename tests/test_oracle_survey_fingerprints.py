"""CPU: the oracle port against the long-run fingerprints of the UNMODIFIED reference (tests/golden/
survey_fingerprints.json, generated from oracle/_ref by tests/golden/make_survey_fingerprints.py), which are also the
values SURVEY.md 8(c) obtained with an independent probe:
    256^2 x 100 frames  37,259,394 steps  c748ab2d23943aa6   (= BASELINE.json configs[0]: 100 headless erode cycles)
    1024^2 x 3 frames    9,683,394 steps  54c13b6060e887a6
    4096^2 x 1 frame    15,844,342 steps  d44f051f7ed88e6e"""
import json
import os

import numpy as np
import pytest

from tests.oracle_lib import PortWorld, RefWorld, ref_available

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FP = json.load(open(os.path.join(G, "survey_fingerprints.json")))
SURVEY = {256: ("c748ab2d23943aa6", 37259394), 1024: ("54c13b6060e887a6", 9683394), 4096: ("d44f051f7ed88e6e", 15844342)}


@pytest.mark.parametrize("run", FP["erode"], ids=lambda r: f"n{r['n']}x{r['calls']}")
def test_port_reproduces_survey_fingerprint(run):
    assert (run["hash"], run["steps"]) == SURVEY[run["n"]]
    w = PortWorld(run["n"])
    w.init_gaussian()
    w.srand(run["seed"])
    assert abs(float(w.plane(0).astype(np.float64).sum()) - run["sum_initial"]) < 1e-6   # summation order differs
    steps = oob = age = 0
    lost = 0.0
    for _ in range(run["calls"]):
        st = w.erode(run["cycles"])
        steps += st.steps; oob += st.exits_oob; age += st.exits_age; lost += st.sediment_lost
    assert steps == run["steps"] and oob == run["exits_oob"] and age == run["exits_age"]
    assert f"{w.hash_height():016x}" == run["hash"]
    assert abs(lost - run["sediment_lost"]) < 1e-9
    # no floor on the height: the 256^2 map is eroded below zero ("unbounded erosion")
    assert abs(float(w.plane(0).astype(np.float64).sum()) - run["sum_final"]) < 1e-6


@pytest.mark.skipif(not ref_available(1024), reason="oracle/_ref not built (needs /root/reference)")
def test_compiled_reference_reproduces_survey_fingerprint_1024():
    run = [r for r in FP["erode"] if r["n"] == 1024][0]
    r = RefWorld(1024, seed=1)
    steps = sum(r.erode_counted(run["cycles"])["steps"] for _ in range(run["calls"]))
    assert steps == run["steps"] and f"{r.hash_height():016x}" == run["hash"]
