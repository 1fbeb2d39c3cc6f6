"""tests/cudasim/build.py -- TEST INFRASTRUCTURE (see cuda_runtime.h in this directory).

Builds tests/cudasim/_build/libcsb200_sim.so from the UNMODIFIED product sources
csrc/kernels.cu, csrc/sles.cu, csrc/runtime.cu, csrc/assembler.cu and csrc/grid_host.cpp with g++ against the
stand-in cuda_runtime.h of this directory.  The only source transformation is syntactic:

  kernel<<<grid, block, smem, stream>>>(args);   ->  cudasim::launch(grid, block, smem, [&]{ kernel(args); });
  extern __shared__ T name[];                    ->  T *name = (T *)cudasim::dyn_smem();

persistent.cu / comm.cu / halo.cu (cooperative launches, bulk copies, peer memory) cannot
run this way; their entry points are stubs that report an error (sim_stubs.cpp), so only
the kernel-per-phase single-rank path is reachable.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "code_saturne_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libcsb200_sim.so")


def _match_paren(s, i):
    assert s[i] == "("
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced parentheses")


def transform(src: str) -> str:
    src = re.sub(r"extern\s+__shared__\s+(\w+)\s+(\w+)\[\];",
                 r"\1 *\2 = (\1 *)cudasim::dyn_smem();", src)
    out = []
    pos = 0
    while True:
        k = src.find("<<<", pos)
        if k < 0:
            out.append(src[pos:])
            break
        # kernel name (with optional template arguments) ends at k
        m = re.search(r"([A-Za-z_]\w*(?:<[^<>;(){}]*>)?)\s*$", src[:k])
        assert m, src[k - 60:k + 20]
        name_start = m.start(1)
        e = src.index(">>>", k)
        cfg = [c.strip() for c in src[k + 3:e].split(",")]
        assert len(cfg) in (2, 3, 4), cfg
        grid, block = cfg[0], cfg[1]
        smem = cfg[2] if len(cfg) > 2 else "0"
        p0 = src.index("(", e)
        assert src[e + 3:p0].strip() == ""
        p1 = _match_paren(src, p0)
        semi = src.index(";", p1)
        assert src[p1 + 1:semi].strip() == ""
        out.append(src[pos:name_start])
        out.append(f"cudasim::launch({grid}, {block}, {smem}, "
                   f"[&]{{ {m.group(1)}{src[p0:p1 + 1]}; }});")
        pos = semi + 1
    return "".join(out)


def build(force=False) -> str:
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in ("kernels.cu", "sles.cu", "runtime.cu", "assembler.cu")]
    deps = srcs + [os.path.join(CSRC, f) for f in ("grid_host.cpp", "csb_internal.h",
                                                    "persistent.h", "grid_host.h")] \
        + [os.path.join(HERE, f) for f in ("cuda_runtime.h", "cudasim.cpp", "sim_stubs.cpp",
                                           "build.py")] \
        + [os.path.join(ROOT, "include", "csb200.h")]
    if (not force and os.path.exists(LIB)
            and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps)):
        return LIB
    cpps = []
    for s in srcs:
        dst = os.path.join(OUT, os.path.basename(s)[:-3] + "_sim.cpp")
        with open(s) as f:
            txt = transform(f.read())
        with open(dst, "w") as f:
            f.write(txt)
        cpps.append(dst)
    cpps += [os.path.join(CSRC, "grid_host.cpp"), os.path.join(HERE, "cudasim.cpp"),
             os.path.join(HERE, "sim_stubs.cpp")]
    # C-ABI entry points that live in comm.cu / halo.cu (never callable here): exported as
    # aborting stubs so that the ctypes mirror can bind the whole header
    hdr = open(os.path.join(ROOT, "include", "csb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(csb_[a-z0-9_]+)\s*\(", hdr)))
    defined = set()
    for c in cpps:
        defined |= set(re.findall(r"\b(csb_[a-z0-9_]+)\s*\(", open(c).read()))
    auto = os.path.join(OUT, "auto_stubs.cpp")
    with open(auto, "w") as f:
        f.write("#include <cstdio>\n#include <cstdlib>\nextern \"C\" {\n")
        for name in declared:
            if name not in defined:
                f.write(f"void *{name}() {{ fprintf(stderr, \"cudasim: {name} needs a real "
                        f"device\\n\"); abort(); return 0; }}\n")
        f.write("}\n")
    cpps.append(auto)
    # -ffp-contract=off: no FMA, like the product's -fmad=false
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-pthread",
           "-ffp-contract=off", "-w", f"-I{HERE}", f"-I{CSRC}",
           f"-I{os.path.join(ROOT, 'include')}", "-o", LIB] + cpps
    subprocess.check_call(cmd, cwd=ROOT)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
