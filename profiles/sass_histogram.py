"""SASS opcode histogram per kernel of the built library, and the ptxas logs beside it.
  python profiles/sass_histogram.py            # -> profiles/r2_sass_opcode_histogram.json, profiles/r2_ptxas_*.log
Runs in the build container (cuobjdump only; no GPU)."""
import collections
import glob
import json
import os
import re
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hierarchicalmatrices.jl_b200", "lib")


def main():
    so = os.path.join(LIB, "libhmb200.so")
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"\(.*", "", name).replace("void ", "")
            cur = out.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    res = {"how": "cuobjdump -sass lib/libhmb200.so, opcodes (with modifiers) counted per kernel; "
                  "tcgen05 has no f64 kind, so the FP64 tensor-core work is DMMA.8x8x4 (mma.sync m8n8k4), "
                  "bulk copies UBLKCP, async copies LDGSTS",
           "kernels": {}}
    for k, c in sorted(out.items()):
        tot = sum(c.values())
        top = dict(c.most_common(14))
        for key in c:
            if re.match(r"DMMA|UBLKCP|LDGSTS|UTMA|UTC|MUFU\.RCP64H|DFMA|LDG\.E\.(?:EF\.)?128", key):
                top[key] = c[key]
        res["kernels"][k] = {"instructions": tot, "opcodes": top}
    json.dump(res, open(os.path.join(ROOT, "profiles", "r2_sass_opcode_histogram.json"), "w"), indent=1)
    for f in glob.glob(os.path.join(LIB, "*.ptxas.log")):
        base = os.path.basename(f).replace(".ptxas.log", "")
        # keep the per-kernel resource lines only, demangled names
        lines = open(f).read().splitlines()
        keep = []
        for ln in lines:
            m = re.search(r"Compiling entry function '(\S+)'", ln)
            if m:
                dn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                keep.append("== " + re.sub(r"\(anonymous namespace\)::", "", dn))
            elif "Used" in ln or "spill" in ln:
                keep.append(ln.strip())
        open(os.path.join(ROOT, "profiles", f"r2_ptxas_{base}.log"), "w").write("\n".join(keep) + "\n")
    print(len(res["kernels"]), "kernels")


if __name__ == "__main__":
    main()
