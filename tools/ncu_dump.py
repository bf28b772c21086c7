import collections, csv, io, re, subprocess, sys
rep, cubin, kname = sys.argv[1:4]
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
m = re.search(r"^\.text\.[^\n]*%s[^\n]*:\n" % re.escape(kname), dis, re.M)
body = dis[m.end():]
nxt = re.search(r"^//-+ \.text\.", body, re.M)
if nxt: body = body[:nxt.start()]
line_of, cur = {}, ("?", 0)
for ln in body.splitlines():
    mm = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if mm:
        cur = (mm.group(1).split("/")[-1], int(mm.group(2))); continue
    mm = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if mm: line_of[int(mm.group(1), 16)] = (cur, mm.group(2))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
sr = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = None
for r in rows[2:]:
    if len(r) < len(hdr): continue
    addr = int(r[ix["Address"]], 16)
    if base is None: base = addr
    key, op = line_of.get(addr - base, (("?", 0), ""))
    st = {h[6:]: int(r[ix[h]]) for h in sr if int(r[ix[h]])}
    print("%05x|%s:%d|%s|%s|%s|%s" % (addr - base, key[0], key[1], op[:60], r[ix["Instructions Executed"]], r[ix["# Samples"]], st))
