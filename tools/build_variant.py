"""Build a variant of libswe_b200.so with extra -D flags for one translation unit (kernel tuning A/B runs):
    python tools/build_variant.py swe_queue.cu -DSWE_PERSIST_BLOCKS=5 -o simplewinderosion_b200/libswe_v5.so
Use it with SWE_B200_LIB=<path> (simplewinderosion_b200/_capi.py)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

unit, out, defs = sys.argv[1], None, []
a = sys.argv[2:]
while a:
    x = a.pop(0)
    if x == "-o":
        out = a.pop(0)
    else:
        defs.append(x)
g.build_cuda()
flags = dict(g.UNITS)[unit]
obj = os.path.join(g.OBJ, unit.replace(".cu", "") + "_" + str(abs(hash(tuple(defs))) % 10**8) + ".o")
subprocess.check_call([g._nvcc(), *g.ARCH, *g.COMMON, *flags, *defs, "-c", os.path.join(g.CSRC, unit), "-o", obj])
objs = [obj if u == unit else os.path.join(g.OBJ, u.replace(".cu", ".o")) for u, _ in g.UNITS]
subprocess.check_call([g._nvcc(), *g.ARCH, "-shared", "-Xcompiler", "-fPIC", "-o", out, *objs, "-cudart", "static",
                       "-Xlinker", "--version-script=" + os.path.join(g.CSRC, "exports.map")])
print("built", out)
