#!/usr/bin/env python
"""Builds the experimental kernel variants (scan3d_math.cuh: S3D_VAR_*) into their own library directories
3dscan_b200/lib_var_<name>/ -- the default build in 3dscan_b200/lib/ is not touched -- and prints each variant's
registers and static instruction budget next to the default's.  The directories are git-ignored but travel with
gpurun; tools/gpu_variants.sh then measures every variant on a B200 (parity first, then the bench line).

    python tools/build_variants.py            # all variants
    python tools/build_variants.py cold rot   # some
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = {
    "v8w5": "-DS3D_K8_WARPS=5 -DS3D_K8_MINB=4 -DS3D_VAR_COLD_OUTLINE=1",
    "v8cold": "-DS3D_VAR_COLD_OUTLINE=1",
    "cold": "-DS3D_VAR_COLD_OUTLINE=1",
    "rowsel": "-DS3D_VAR_ROWSEL_RCP=1",
    "rot": "-DS3D_VAR_TERM_ROTATE=1",
    "all": "-DS3D_VAR_COLD_OUTLINE=1 -DS3D_VAR_ROWSEL_RCP=1 -DS3D_VAR_TERM_ROTATE=1",
    "pass4": "-DS3D_VAR_PASS_UNROLL=1",
    "io500": "-DS3D_VAR_IO_SLEEP_NS=500",
    "io1000": "-DS3D_VAR_IO_SLEEP_NS=1000",
    # capture-side remap (scan3d_aux_kernels.cu; timed by tools/bench_aux.py, checked by tests/aux_check_runner.py)
    "remapu2": "-DS3D_VAR_REMAP_UNROLL=2",
    "remapwin": "-DS3D_VAR_REMAP_WINDOW=1 -DS3D_VAR_REMAP_UNROLL=2",
    "remaptile": "-DS3D_VAR_REMAP_TILED=1",
    "remaptile2": "-DS3D_VAR_REMAP_TILED=1 -DS3D_VAR_REMAP_TILED_MINB=2",
}
KERNEL = "k_fused7ILi8ELi2ELi7ELi3ELb0"


def budget(libdir):
    obj = os.path.join(libdir, "scan3d_fused_kernel7.o")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_budget.py"), obj, KERNEL],
                         capture_output=True, text=True).stdout
    total = int(re.search(r"instructions: (\d+)", out).group(1))
    loops = {int(m.group(2)): int(m.group(1))
             for m in re.finditer(r"^\s+(\d+)\s+0x\w+\.\.0x\w+\s+\('scan3d_fused_kernel7\.cu', (\d+)\)", out, re.M)}
    log = open(os.path.join(libdir, "scan3d_fused_kernel7.cu.ptxas.log")).read()
    blocks = log.split("Compiling entry function")
    regs = spill = "?"
    for b in blocks:
        if KERNEL in b.split("\n")[0]:
            m = re.search(r"Used (\d+) registers", b)
            s = re.search(r"(\d+) bytes spill stores", b)
            regs = m.group(1) if m else "?"
            spill = s.group(1) if s else "?"
    big = sorted(loops.items(), key=lambda kv: -kv[1])[:5]
    return total, regs, spill, big


def main():
    names = sys.argv[1:] or list(VARIANTS)
    rows = [("default", budget(os.path.join(ROOT, "3dscan_b200", "lib")))]
    for n in names:
        libdir = os.path.join(ROOT, "3dscan_b200", "lib_var_" + n)
        env = dict(os.environ, SCAN3D_LIBDIR=libdir, SCAN3D_BUILD_DEFS=VARIANTS[n])
        subprocess.check_call([sys.executable, os.path.join(ROOT, "3dscan_b200", "build.py")], env=env,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        rows.append((n, budget(libdir)))
    print(f"{'variant':10s} {'instr':>6s} {'regs':>5s} {'spill':>6s}  largest loops (source line: static size)")
    for n, (total, regs, spill, big) in rows:
        print(f"{n:10s} {total:6d} {regs:>5s} {spill:>6s}  " + ", ".join(f"{l}: {c}" for l, c in big))


if __name__ == "__main__":
    main()
