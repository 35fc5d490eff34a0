"""Per-source-line instruction counts of one kernel from an ncu report (the CSV source page has no CUDA-line
metrics, so SASS rows are joined with `nvdisasm -g` line info by instruction offset).

    python scripts/ncu_by_line.py <rep.ncu-rep> <object.o> <mangled kernel substring> [kernel index in report]
"""
import csv, collections, os, re, subprocess, sys, tempfile

rep, obj, sub = sys.argv[1:4]
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], cwd=tmp, check=True, capture_output=True, text=True).stdout
line_of, chain_of, cur, chain, on, fresh = {}, {}, ("?", 0), [], False, True
for l in dis.split("\n"):
    if l.startswith(".text."):
        on = sub in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]*)", line (\d+)', l)
    if m:
        if fresh:
            chain, fresh = [], False
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        cur = chain[0]
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", l)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
        chain_of[int(m.group(1), 16)] = list(chain)
        fresh = True
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.split("\n")))
k, hdr, data = -1, None, []
for r in rows:
    if r and r[0] == "Kernel Name":
        k += 1
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if k == which and r:
        data.append(r)
ia, ie, it, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
base = int(data[0][ia], 16)
agg = collections.defaultdict(lambda: [0, 0, 0])
lvl = [collections.defaultdict(lambda: [0, 0, 0]) for _ in range(3)]
tot = [0, 0, 0]
for r in data:
    off = int(r[ia], 16) - base
    (f, ln), _ = line_of.get(off, (("?", 0), ""))
    v = (int(r[ie]), int(r[it]), int(r[isamp]))
    ch = chain_of.get(off, [("?", 0)])[::-1]          # outermost frame first
    for j in range(3):
        agg[(f, ln)][j] += v[j]; tot[j] += v[j]
        for d in range(3):
            lvl[d][ch[min(d, len(ch) - 1)]][j] += v[j]
print(f"total warp-inst {tot[0]:,}  thread-inst {tot[1]:,}  samples {tot[2]:,}")
srcs = {}
depth = int(os.environ.get("DEPTH", "-1"))
table = agg if depth < 0 else lvl[depth]
for (f, ln), v in sorted(table.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("TOP", "70"))]:
    path = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "csrc", f)
    if f not in srcs and os.path.isfile(path):
        srcs[f] = open(path).read().split("\n")
    text = srcs[f][ln - 1].strip()[:90] if f in srcs and 0 < ln <= len(srcs[f]) else ""
    print(f"{100 * v[0] / tot[0]:5.1f}% inst {100 * v[2] / max(1, tot[2]):5.1f}% smp  lanes {v[1] / max(1, v[0]):4.1f}  {f}:{ln}  {text}")
