"""Build a variant of the product library with extra flags for the rate TU: tools/build_variant.py <name> <flags...>
-> hmp3_b200/_lib/var_<name>.so (select it with HMP3_B200_LIB).  Experiments only."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
TU = "kernels_rate.cu"
args = sys.argv[1:]
if args and args[0] == "--tu":          # which translation unit the variant rebuilds (default: the nested serial stage)
    TU, args = args[1], args[2:]
name, flags = args[0], args[1:]
objdir = os.path.join(ROOT, "hmp3_b200", "_build")
obj = os.path.join(objdir, "%s_%s.o" % (TU.rsplit(".", 1)[0], name))
extra = [f for f in g.TUS[TU]]
for tool in ("-Xcicc", "-Xptxas"):          # a variant's own optimisation level replaces the default one
    if tool in flags and tool in extra:
        i = extra.index(tool)
        del extra[i:i + 2]
cmd = ["timeout", "1500", "/usr/local/cuda/bin/nvcc"] + g.NVCC_COMMON + extra + flags + ["-c", "-o", obj, os.path.join(g.CSRC, TU)]
r = subprocess.run(cmd, capture_output=True, text=True)
if r.returncode:
    print(r.stderr[-3000:]); sys.exit(1)
for line in r.stderr.splitlines():
    if "k_rate" in line and "Compiling" in line or ("registers" in line and "k_rate" in prev):
        print(line[:200])
    prev = line
objs = [os.path.join(objdir, t.rsplit(".", 1)[0] + ".o") for t in g.TUS if t != TU] + [obj]
out = os.path.join(g.LIBDIR, "var_%s.so" % name)
subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-shared", "-o", out] + objs + ["-lcudart"])
print("built", out)
