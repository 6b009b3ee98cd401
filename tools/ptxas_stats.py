#!/usr/bin/env python
"""Registers / spills / stack of every kernel in libparvar_b200.so (ptxas -v), one line per kernel.

usage: python tools/ptxas_stats.py [filter]      (rebuilds the library with -Xptxas -v; no GPU needed)
"""
import importlib
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
b = importlib.import_module("parameter-variability_b200.build")


def main():
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    b.generate_sources()
    b.LIB_DIR.mkdir(exist_ok=True)
    cmd = [b._nvcc(), "-Xptxas", "-v", *b.NVCC_FLAGS, "-o", str(b.LIB_PATH), str(b.CSRC / "pv_abi.cu")]
    err = subprocess.run(cmd, capture_output=True, text=True, cwd=str(b.CSRC)).stderr
    b.STAMP.write_text(b._source_hash() + "\n")
    names = {}
    for m in re.finditer(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                         r"ptxas info\s+: Used (\d+) registers", err):
        names[m.group(1)] = tuple(int(x) for x in m.groups()[1:])
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    for (mangled, (stack, sst, sld, regs)), d in zip(names.items(), dem):
        d = d.replace("(pv_launch_params)", "").replace("void ", "")
        if flt in d:
            print(f"{regs:4d} regs  stack {stack:5d}  spill st {sst:6d} ld {sld:6d}  {d}")


if __name__ == "__main__":
    main()
