#!/usr/bin/env python
"""Stall samples and shared-memory bank conflicts per CUDA SOURCE LINE of a captured kernel.
ncu's SASS page has no line column, so the per-instruction rows are aligned (by instruction index) with
`nvdisasm -g` of the same cubin, which carries `//## File ..., line N` markers (-lineinfo build).
usage: ncu_lines.py <report.ncu-rep> <libcpbus.so that ran> [top] [source file as of that build]"""
import csv, io, re, subprocess, sys, tempfile, os, glob

rep, so = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
kname = raw[2][raw[0].index("Kernel Name")]
hdr, data = None, []
for r in rows:
    if r and r[0] == "Address":
        if hdr is not None:
            break
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        data.append(r)
ci = {n: i for i, n in enumerate(hdr)}
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
# template args of the captured kernel -> mangled fragment
m = re.search(r"fanout_kernel<(\d), (\d), (\d), (\d), (\d)>", kname)
frag = "fanout_kernelILi%sELb%sELb%sELb%sELb%sE" % m.groups() if m else None
lines_of = []
cur, inside = None, False
for l in dis:
    if l.startswith(".text.") and l.endswith(":"):
        inside = frag is not None and frag in l
        continue
    if not inside:
        continue
    if l.startswith("//---") or (l.startswith("\t.section") and lines_of):
        if lines_of:
            break
    mm = re.search(r'//## File ".*?", line (\d+)', l)
    if mm:
        cur = int(mm.group(1)); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines_of.append(cur)
n = min(len(lines_of), len(data))
print(f"kernel {kname[:60]}  sass rows {len(data)}  nvdisasm instructions {len(lines_of)}")
src = open(sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "containerpilot_b200", "csrc", "cpbus_kernels.cuh")).read().split("\n")
st = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
agg = {}
for i in range(n):
    r = data[i]
    a = agg.setdefault(lines_of[i], {"n": 0, "conf": 0, "st": {}})
    a["n"] += int(r[ci["# Samples"]] or 0)
    a["conf"] += int(float(r[ci["L1 Wavefronts Shared Excessive"]] or 0))
    for k in st:
        v = int(r[ci[k]] or 0)
        if v:
            a["st"][k[6:]] = a["st"].get(k[6:], 0) + v
tot = sum(a["n"] for a in agg.values())
print("total samples", tot, " excessive shared wavefronts", sum(a["conf"] for a in agg.values()))
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1]["n"])[:top]:
    t3 = dict(sorted(a["st"].items(), key=lambda kv: -kv[1])[:3])
    text = src[ln - 1].strip()[:88] if ln and ln <= len(src) else "?"
    print(f"{a['n']:7d} {100*a['n']/max(tot,1):5.1f}%  conf={a['conf']:>9d}  L{ln}: {text:88s} {t3}")
