"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by source line and by function.
usage: ncu -i rep.ncu-rep --page source --print-source cuda,sass --csv > src.csv ; python scripts/ncu_lines.py src.csv [top]"""
import csv, re, sys, os, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = csv.reader(open(path, newline=''))
cur_file = None; hdr = None; per_line = collections.Counter(); per_line_inst = collections.Counter(); text = {}
stalls = collections.Counter()
cur_line = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = os.path.basename(r[1]); continue
    if r[0] == "Line No": hdr = r; iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed"); continue
    if r[0] in ("Function Name", "File Name") or hdr is None: continue
    if r[0] != "":
        cur_line = (cur_file, int(r[0])); text[cur_line] = r[1].strip(); continue
    try:
        per_line[cur_line] += int(r[iS]); per_line_inst[cur_line] += int(r[iI])
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h: stalls[h] += int(r[i])
    except (ValueError, IndexError):
        pass
tot_s = sum(per_line.values()); tot_i = sum(per_line_inst.values())
print(f"total samples {tot_s}  warp instructions {tot_i}")
ts = max(1, sum(stalls.values()))
print("\n== warp stall reasons (all samples)")
for k, v in stalls.most_common(10): print(f"  {100*v/ts:5.1f}%  {k}")
defs = collections.defaultdict(list)
src_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "metaworld_b200", "csrc")
for f in set(k[0] for k in per_line):
    p = os.path.join(src_dir, f)
    if not os.path.exists(p): continue
    for i, l in enumerate(open(p), 1):
        m = re.match(r"^(?:template <[^>]*> )?(?:DEV|__device__|__global__|static)[^;(]*?\b(\w+)\s*\(", l)
        if m and not l.strip().endswith(";"): defs[f].append((i, m.group(1)))
def func(k):
    name = "?"
    for i, n in defs.get(k[0], []):
        if i <= k[1]: name = n
        else: break
    return name
byf = collections.Counter(); byfi = collections.Counter()
for k, v in per_line.items(): byf[(k[0], func(k))] += v
for k, v in per_line_inst.items(): byfi[(k[0], func(k))] += v
print("\n== by function (stall samples %, warp-instruction %)")
for (f, n), v in byf.most_common(30): print(f"  {100*v/tot_s:5.1f}%  {100*byfi[(f,n)]/max(1,tot_i):5.1f}%  {f}:{n}")
print("\n== top source lines by samples")
for k, v in per_line.most_common(top): print(f"  {100*v/tot_s:5.2f}%  inst {100*per_line_inst[k]/max(1,tot_i):5.2f}%  {k[0]}:{k[1]}  {text.get(k,'')[:110]}")
