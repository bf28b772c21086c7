#!/usr/bin/env python
"""Join an ncu SASS-level source page with nvdisasm --print-line-info to get per-source-line
executed-instruction counts and stall samples (the GPU box's source paths differ, so ncu cannot
correlate by itself).

usage: ncu_by_line.py report.ncu-rep cubin mangled_kernel_substring [top_n]
"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep, cubin, kname = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
    dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
    # isolate the kernel's text section
    m = re.search(r"^\.text\.[^\n]*%s[^\n]*:\n" % re.escape(kname), dis, re.M)
    body = dis[m.end():]
    nxt = re.search(r"^//-+ \.text\.", body, re.M)
    if nxt:
        body = body[:nxt.start()]
    line_of, cur = {}, ("?", 0)
    for ln in body.splitlines():
        mm = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if mm:
            cur = (mm.group(1).split("/")[-1], int(mm.group(2)))
            continue
        mm = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
        if mm:
            line_of[int(mm.group(1), 16)] = (cur, mm.group(2))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    base = None
    inst = collections.Counter(); samp = collections.Counter(); tinst = collections.Counter()
    stalls = collections.defaultdict(collections.Counter)
    sr = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        addr = int(r[ix["Address"]], 16)
        if base is None:
            base = addr
        key, _ = line_of.get(addr - base, (("?", 0), ""))
        inst[key] += int(r[ix["Instructions Executed"]])
        tinst[key] += int(r[ix["Thread Instructions Executed"]])
        samp[key] += int(r[ix["# Samples"]])
        for h in sr:
            v = int(r[ix[h]])
            if v:
                stalls[key][h[6:]] += v
    tot_i, tot_s = sum(inst.values()), sum(samp.values())
    print("total warp instr %d, samples %d" % (tot_i, tot_s))
    print("%-28s %6s %7s %7s  top stalls" % ("file:line", "", "inst%", "samp%"))
    for key, n in sorted(inst.items(), key=lambda kv: -(kv[1] + 3000 * samp[kv[0]]))[:top]:
        st = ", ".join("%s %d" % kv for kv in stalls[key].most_common(3))
        print("%-28s %6d %6.2f%% %6.2f%%  %s" % ("%s:%d" % key, n // 1000, 100.0 * n / tot_i, 100.0 * samp[key] / tot_s, st))


if __name__ == "__main__":
    main()
