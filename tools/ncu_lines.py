"""Developer helper: per-CUDA-source-line instruction and stall shares of one kernel from an .ncu-rep
(maps the SASS page onto line numbers with nvdisasm -g on the cubins extracted from the in-tree .so)."""
import csv, re, subprocess, sys, os, tempfile
rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "regtools_b200", "libregtools_jx.so")], cwd=tmp, capture_output=True)
dis = None
for f in os.listdir(tmp):
    if f.endswith(".cubin"):
        d = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kernel in d:
            dis = d.split("\n")
# `rep` is an .ncu-rep, or the CSV of its source page made on the GPU box (`ncu -i x.ncu-rep --page source --csv > x.source.csv`)
out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
# pick the table whose preceding "Kernel Name" row mentions the kernel
pick = his[0]
for hi in his:
    name = " ".join(rows[hi - 1]) if hi else ""
    if kernel in name:
        pick = hi; break
nxt = [x for x in his if x > pick]
end = (nxt[0] - 1) if nxt else len(rows)
h = rows[pick]; data = [r for r in rows[pick + 1:end] if len(r) > 10 and r[0] != "Address"]
ss = h.index("Warp Stall Sampling (All Samples)"); ie = h.index("Instructions Executed")
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and kernel in l][0]
lines, cur = [], None
for l in dis[start + 1:]:
    if (l.startswith(".text.") or l.startswith("\t.section")) and lines:
        break
    m = re.search(r'//## File ".*?/([A-Za-z_0-9]+\.cuh?)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
def I(x):
    try: return int(x)
    except ValueError: return 0
print("sass instrs", len(lines), "ncu rows", len(data))
agg = {}
for i, r in enumerate(data):
    if i < len(lines):
        a = agg.setdefault(lines[i], [0, 0]); a[0] += I(r[ie]); a[1] += I(r[ss])
ti = sum(v[0] for v in agg.values()) or 1; ts = sum(v[1] for v in agg.values()) or 1
srcs = {}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if k:
        path = os.path.join(root, "regtools_b200", "csrc", k[0])
        if os.path.exists(path):
            srcs.setdefault(k[0], open(path).read().split("\n"))
            text = srcs[k[0]][k[1] - 1].strip()[:105]
    print(f"{(k[0] if k else '?'):>16}:{(k[1] if k else 0):<5d} inst {100*v[0]/ti:5.1f}%  stall {100*v[1]/ts:5.1f}%  {text}")
# optional: shares of line ranges, e.g. "inflate.cu:583-650" (argv[4:])
for spec in sys.argv[4:]:
    f, rg = spec.split(":"); lo, hi = (int(x) for x in rg.split("-"))
    a = [0, 0]
    for k, v in agg.items():
        if k and k[0] == f and lo <= k[1] <= hi: a[0] += v[0]; a[1] += v[1]
    print(f"{spec:>28} inst {100*a[0]/ti:5.1f}%  stall {100*a[1]/ts:5.1f}%")
