#!/usr/bin/env python
"""Build a tuning variant of the library into build/variants/<name>.so (same C ABI; optionally only some registry models).

usage: python tools/build_variant.py NAME [--models simple_pk,...] [-DMACRO=VALUE ...] [-v FILTER]
       PARVAR_B200_LIB=build/variants/NAME.so python tools/time_c2.py          (on the GPU box)

-v FILTER prints ptxas' registers / spills for the kernels whose demangled name contains FILTER.  No GPU needed.
"""
import importlib
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
b = importlib.import_module("parameter-variability_b200.build")
cg = importlib.import_module("parameter-variability_b200.codegen")


def main():
    args = sys.argv[1:]
    name = args.pop(0)
    models, defs, flt = None, [], None
    while args:
        a = args.pop(0)
        if a == "--models":
            models = args.pop(0).split(",")
        elif a == "-v":
            flt = args.pop(0)
        elif a.startswith("-D") or a.startswith("-maxrregcount") or a.startswith("-X"):
            defs.append(a)
        else:
            raise SystemExit(f"unknown argument {a}")
    out_dir = ROOT / "build" / "variants"
    out_dir.mkdir(parents=True, exist_ok=True)
    extra = []
    if models:
        gen = out_dir / f"{name}_generated"
        gen.mkdir(exist_ok=True)
        sbml_dir = ROOT / "parameter-variability_b200" / "models" / "sbml"
        ems = [cg.generate_model(sbml_dir / f"{m}.xml", cg.BUILTIN[m]) for m in models]
        for em in ems:
            (gen / f"{em.name}.cuh").write_text(em.header.replace('#include "../pv_common.cuh"', f'#include "{b.CSRC / "pv_common.cuh"}"'))
        (gen / "model_tables.inc").write_text(cg.emit_tables(ems))
        (gen / "models.inc").write_text(cg.models_inc(ems))
        extra.append(f'-DPV_MODELS_INC="{gen / "models.inc"}"')
    out = out_dir / f"{name}.so"
    cmd = [b._nvcc(), "-Xptxas", "-v", *b.NVCC_FLAGS, *extra, *defs, "-o", str(out), str(b.CSRC / "pv_abi.cu")]
    proc = subprocess.run(cmd, capture_output=True, text=True, cwd=str(b.CSRC))
    if proc.returncode != 0:
        raise SystemExit(proc.stdout + proc.stderr)
    if flt is not None:
        names = {}
        for m in re.finditer(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                             r"ptxas info\s+: Used (\d+) registers", proc.stderr):
            names[m.group(1)] = tuple(int(x) for x in m.groups()[1:])
        dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
        for (_, (stack, sst, sld, regs)), d in zip(names.items(), dem):
            if flt in d:
                print(f"{regs:4d} regs  stack {stack:5d}  spill st {sst:6d} ld {sld:6d}  {d.replace('void ', '')[:150]}")
    print(out)


if __name__ == "__main__":
    main()
