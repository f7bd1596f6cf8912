"""Aggregate an ncu report's per-SASS-instruction samples / executed instructions by CUDA source line.

usage: python tools/ncu_lines.py <report.ncu-rep> <lib.so> <kernel-substring> [top_n]
Works without a GPU: reads the report with `ncu -i`, maps SASS offsets to source lines with
`nvdisasm -g` on the cubin extracted from the library (compile with -lineinfo)."""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, lib, kern = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr = rows[h]
iA, iS, iN, iI = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
inst = []
for r in rows[h + 1:]:
    if len(r) <= iI or not r[iA]:
        continue
    try:
        inst.append((int(r[iA], 0), r[iS], int(r[iN] or 0), int(r[iI] or 0), [int(r[i] or 0) for i in stall_cols]))
    except ValueError:
        pass
base = min(a for a, *_ in inst)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
line_of = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur_fn, cur_line, take = None, None, False
    for ln in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            take = kern in m.group(1) and (len(sys.argv) <= 5 or sys.argv[5] in m.group(1))
            continue
        if not take:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            line_of.setdefault(int(m.group(1), 16), cur_line)
agg = collections.defaultdict(lambda: [0, 0, [0] * len(stall_cols)])
for a, s, n, i, st in inst:
    key = line_of.get(a - base, ("?", 0))
    agg[key][0] += n
    agg[key][1] += i
    for k, v in enumerate(st):
        agg[key][2][k] += v
tot_n = sum(v[0] for v in agg.values()) or 1
tot_i = sum(v[1] for v in agg.values()) or 1
src_cache = {}
def src(fn, ln):
    for root in ("recsys2019_deeplearning_evaluation_b200/csrc", "."):
        p = os.path.join(root, fn)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            L = src_cache[p]
            return L[ln - 1].strip()[:90] if 0 < ln <= len(L) else ""
    return ""
print("total samples %d, total warp-instructions %d" % (tot_n, tot_i))
for key, (n, i, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    top = sorted(zip(st, [hdr[c] for c in stall_cols]), reverse=True)[:2]
    print("%5.1f%% smp %5.1f%% inst  %s:%d  [%s]  %s" % (100.0 * n / tot_n, 100.0 * i / tot_i, key[0], key[1],
          ",".join("%s=%d%%" % (nm.replace("stall_", ""), 100 * v // max(n, 1)) for v, nm in top), src(*key)))
