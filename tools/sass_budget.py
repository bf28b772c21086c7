#!/usr/bin/env python
"""Static instruction budget of one kernel from its SASS + line info (no GPU needed):

    python tools/sass_budget.py [object] [substring of the mangled kernel name]
    python tools/sass_budget.py 3dscan_b200/lib/scan3d_fused_kernel7.o k_fused7ILi8ELi2ELi7ELi3ELb0

Prints the kernel's instruction count, every loop (backward branch) with its static size and source line, and
for the largest loops the opcode mix and the source lines that contribute most.  Multiply the loop sizes by their
trip counts (FP64 pass: 2 per tile and thread, triangulation: one per surviving pixel) for a dynamic estimate;
both sides of run-time branches are counted, so a loop's size is an upper bound of what one trip executes."""
import collections
import os
import re
import subprocess
import sys
import tempfile

obj = sys.argv[1] if len(sys.argv) > 1 else "3dscan_b200/lib/scan3d_fused_kernel7.o"
want = sys.argv[2] if len(sys.argv) > 2 else "k_fused7ILi8ELi2ELi7ELi3ELb0"
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
text = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(text) if l.startswith(".text.") and want in l and l.rstrip().endswith("st:"))
end = next(i for i in range(start + 1, len(text)) if text[i].startswith((".text.", ".section")))
print(text[start].strip())
cur, ins, pend, label_addr = None, [], [], {}
for l in text[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"^(\.L_x_\d+):", l)
    if m:
        pend.append(m.group(1))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        addr = int(m.group(1), 16)
        for p in pend:
            label_addr[p] = addr
        pend = []
        ins.append((addr, re.sub(r"^@!?U?P\d+\s+", "", m.group(2)), cur))
print("instructions:", len(ins))
loops = []
for addr, txt, src in ins:
    if txt.startswith("BRA"):
        m = re.search(r"(\.L_x_\d+)", txt)
        if m and label_addr.get(m.group(1), 1 << 60) <= addr:
            a = label_addr[m.group(1)]
            if src and "intrinsics" in src[0]:
                continue   # return branch of an out-of-line warp-collective trampoline (shfl / vote in divergent code), not a loop
            loops.append((sum(1 for i in ins if a <= i[0] <= addr), a, addr, src))
print("loops (size, first, last, source line of the backward branch):")
for n, a, b, src in sorted(loops, key=lambda t: t[1]):
    print(f"  {n:5d}  {a:#07x}..{b:#07x}  {src}")
for n, a, b, src in sorted(loops, reverse=True)[:6]:
    if n < 100:
        continue
    ops, lines = collections.Counter(), collections.Counter()
    for addr, txt, s in ins:
        if a <= addr <= b:
            op = txt.split()[0]
            base = op.split(".")[0]
            if base == "IMAD" and ".MOV" in op:
                base = "IMAD.MOV"
            ops[base] += 1
            lines[s] += 1
    print(f"\nloop at {src} ({n} instructions)")
    print("  opcodes:", ", ".join(f"{k} {v}" for k, v in ops.most_common(16)))
    print("  lines:  ", ", ".join(f"{k[0]}:{k[1]} {v}" for k, v in lines.most_common(12) if k))
