"""Developer helper: per-CUDA-source-line stall samples of one kernel from an .ncu-rep (maps the SASS
page onto line numbers with nvdisasm -g on the cubin extracted from the in-tree .so)."""
import csv, re, subprocess, sys, os, tempfile
rep, kernel = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "regtools_b200", "libregtools_jx.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) > 10 and r[0] != "Address"]
ss = h.index("Warp Stall Sampling (All Samples)")
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and kernel in l][0]
lines, cur = [], None
for l in dis[start + 1:]:
    if (l.startswith(".text.") or l.startswith("\t.section")) and lines:
        break
    m = re.search(r'//## File ".*?", line (\d+)', l)
    if m:
        cur = int(m.group(1)); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
def I(x):
    try: return int(x)
    except ValueError: return 0
print("sass instrs", len(lines), "ncu rows", len(data))
agg = {}
for i, r in enumerate(data):
    if i < len(lines): agg[lines[i]] = agg.get(lines[i], 0) + I(r[ss])
tot = sum(agg.values())
src = open(os.path.join(root, "regtools_b200", "csrc", "kernels.cu")).read().split("\n")
for ln, c in sorted(agg.items(), key=lambda x: -x[1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
    print(str(ln).rjust(5), str(c).rjust(6), "%4.1f%%" % (100 * c / max(tot, 1)), src[ln - 1].strip()[:110] if ln else "")
