"""Turn the scratch ncu outputs in gpurun_out/ into the committed summaries under profiles/."""
import csv, json, os, subprocess, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"

def launches(path, out):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    d = defaultdict(list)
    for r in rows[h + 1:]:
        d[r[4].split("(")[0].replace("void ", "")].append(float(r[-1]))
    tot = sum(sum(v) for v in d.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({os.path.basename(path)}): `ncu --metrics gpu__time_duration.sum --clock-control none` "
                f"over `python bench.py --steps 1 --warmup 3 --skip-cpu --skip-cascade --skip-layered --skip-cfg4` ({sum(len(v) for v in d.values())} launches"
                f"{' of the timed call, after --launch-skip 900' if tag == 'r2t' else ' from the start'}; "
                f"cold-cache, serialised: compare SHARES)\n\n| kernel | launches | total ms | share | mean us |\n|---|---|---|---|---|\n")
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| {k} | {len(v)} | {sum(v)/1e6:.3f} | {100*sum(v)/tot:.1f} % | {sum(v)/len(v)/1e3:.1f} |\n")

def ncu_summary(rep, out, title):
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    with open(out, "w") as f:
        f.write(f"# {title}\n# source: ncu --set full --clock-control none --import-source on, one launch\n" + txt)
    return txt

src = os.path.join(G, f"stat_parity_tiles_{tag}.txt")
if os.path.exists(src):
    open(os.path.join(P, "r2_stat_parity_tiles.txt"), "w").write(open(src).read())
if os.path.exists(os.path.join(G, f"launches_{tag}.csv")):
    launches(os.path.join(G, f"launches_{tag}.csv"), os.path.join(P, f"{tag}_launches_summary.md"))
traffic = {}
for name, title in (("tile_fly", "tile_fly_kernel<240,60,256> (Wind::fly + erosion + World::cascade per particle step; one super-round = 24 rounds x 9 micro-rounds of every 240x60 tile), 4096^2 FBm map, 4,194,304 particles per call, launch 41 of ~97 (mid-call, tiles fully populated)"),
                    ("colour_persistent", "colour_persistent_kernel (Wind::fly + erosion + World::cascade per particle step, all micro-rounds of one World::erode), 4096^2 FBm map, 4,194,304 particles"),
                    ("par_round", "par_round_kernel, 4096^2 FBm map, 4,194,304 particles, one round (step ~150 of a call)"),
                    ("cascade_stamped", "cascade_tile_kernel<stamped,TMA>, same run: World::cascade at the cells touched in the round"),
                    ("cascade_sweep8192", "cascade_tile_kernel<plain,TMA>, 8192^2 U(0,1) map, one whole-map sweep (sweep 3: dense path, everything moves)"),
                    ("noise_stack", "noise_stack_kernel, 4096^2, eight layers of FastNoiseLite OpenSimplex2 FBm (map::init generate on the device)"),
                    ("cascade_sweep_at_rest", "cascade_tile_kernel<plain,TMA>, 8192^2 map at rest (all slopes below the repose angle): pair scan only"),
                    ("cascade_sweep_converged", "cascade_tile_kernel<plain,TMA>, 4096^2 U(0,1) map, sweep 161 of the relaxation (pair-bitmap mode, ~10 % of the cells still move)")):
    rep = os.path.join(G, f"{name}_{tag}.ncu-rep")
    if os.path.exists(rep):
        txt = ncu_summary(rep, os.path.join(P, f"{tag}_{name}_ncu.txt"), title)
        rd = wr = None
        for ln in txt.splitlines():
            f = ln.split()
            if len(f) >= 3 and f[0] == "dram__bytes_read.sum": rd = float(f[1]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[f[2]]
            if len(f) >= 3 and f[0] == "dram__bytes_write.sum": wr = float(f[1]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[f[2]]
        if rd is not None and wr is not None: traffic[name] = rd + wr
old = {}
try:
    old = json.load(open(os.path.join(P, "traffic.json")))["bytes_per_launch"]
except Exception:
    pass
old.update(traffic)
json.dump({"source": "profiles/r*_*_ncu.txt (dram__bytes_read.sum + dram__bytes_write.sum, one launch; latest round per kernel)",
           "bytes_per_launch": old}, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(traffic)
