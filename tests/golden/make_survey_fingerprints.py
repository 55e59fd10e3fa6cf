"""Generates tests/golden/survey_fingerprints.json from oracle/_ref (the UNMODIFIED reference compiled headless):
the three long runs SURVEY.md 8(c) lists beside the 512^2 one -- 256^2 x 100 frames (= BASELINE.json configs[0],
"100 headless erode cycles"), 1024^2 x 3 frames, 4096^2 x 1 frame; Gaussian-hill init, srand(1), 8*tilesize particles
per frame (SimpleWindErosion.cpp:217). Needs /root/reference to have been built into oracle/_ref (make -C oracle ref).
    python tests/golden/make_survey_fingerprints.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.oracle_lib import RefWorld  # noqa: E402

SURVEY = {256: ("c748ab2d23943aa6", 37259394), 1024: ("54c13b6060e887a6", 9683394), 4096: ("d44f051f7ed88e6e", 15844342)}


def main() -> None:
    runs = []
    for n, calls in ((256, 100), (1024, 3), (4096, 1)):
        r = RefWorld(n, seed=1)
        h0 = r.sum_height()
        tot = {"steps": 0, "exits_oob": 0, "exits_age": 0, "sediment_lost": 0.0}
        for _ in range(calls):
            st = r.erode_counted(8 * n)
            for k in tot:
                tot[k] += st[k]
        run = {"n": n, "calls": calls, "cycles": 8 * n, "seed": 1, "sum_initial": h0, "sum_final": r.sum_height(),
               "hash": f"{r.hash_height():016x}", **tot}
        print(run, "survey:", SURVEY[n])
        assert (run["hash"], run["steps"]) == SURVEY[n], "the compiled reference no longer reproduces SURVEY.md 8(c)"
        runs.append(run)
    with open(os.path.join(HERE, "survey_fingerprints.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_survey_fingerprints.py from oracle/_ref (unmodified reference)",
                   "erode": runs}, f, indent=1)


if __name__ == "__main__":
    main()
