"""Build a VARIANT of libb2u.so for A/B runs on one box:  python tests/build_variant.py <name> [-DMACRO=..]...
writes unet_research_b200/csrc/libb2u_<name>.so (git-ignored, travels with gpurun); select it with B2U_LIB=<path>."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as G

name, flags = sys.argv[1], sys.argv[2:]
bdir = os.path.join(G.CSRC, "build", "variant_" + name)
os.makedirs(bdir, exist_ok=True)
procs, objs = [], []
for src in G.SOURCES:
    obj = os.path.join(bdir, src.replace(".cu", ".o"))
    objs.append(obj)
    cmd = [G._nvcc(), *G.NVCC_FLAGS, *flags, "-c", os.path.join(G.CSRC, src), "-o", obj]
    procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for cmd, p in procs:
    out, _ = p.communicate()
    if p.returncode != 0:
        raise SystemExit("nvcc failed: " + " ".join(cmd) + "\n" + out)
out = os.path.join(G.CSRC, f"libb2u_{name}.so")
subprocess.check_call([G._nvcc(), "-shared", "-Wno-deprecated-gpu-targets", "-o", out, *objs, "-lcudart_static", "-lrt", "-lpthread", "-ldl"])
print(out)
