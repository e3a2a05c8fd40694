#!/usr/bin/env python3
"""Aggregate an ncu SASS source page (ncu -i rep --page source --csv) per CUDA source line, using nvdisasm line info.
  python tools/ncu_lines.py <src.csv> <cubin> <kernel-substring> [source-file]"""
import csv, re, subprocess, sys, collections
src_csv, cubin, ksub = sys.argv[1:4]
srcfile = sys.argv[4] if len(sys.argv) > 4 else None
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# walk functions; record for the wanted function address -> (file, line)
addr2line = {}
cur_fn = None; cur_line = None
for ln in dis:
    m = re.match(r"\s*\.text\.(\S+):", ln) or re.match(r"\s*//-+ \.text\.(\S+)", ln)
    if m: cur_fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur_line = (m.group(1), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m and cur_fn and ksub in cur_fn:
        addr2line[int(m.group(1), 16)] = cur_line
r = list(csv.reader(open(src_csv)))
hi = next(i for i, row in enumerate(r) if row and row[0] == "Address")
h = r[hi]
rows = [dict(zip(h, x)) for x in r[hi + 1:] if len(x) == len(h)]
base = min(int(x["Address"], 16) for x in rows)
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for x in rows:
    a = int(x["Address"], 16) - base
    key = addr2line.get(a)
    s = float(x["# Samples"] or 0); ins = float(x["Instructions Executed"] or 0); thr = float(x["Thread Instructions Executed"] or 0)
    agg[key][0] += s; agg[key][1] += ins; agg[key][2] += thr
tot = sum(v[0] for v in agg.values()); toti = sum(v[1] for v in agg.values())
lines = {}
print("samples %d, warp instructions %d, mapped addresses %d" % (tot, toti, len(addr2line)))
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if v[0] < tot * 0.004 and v[1] < toti * 0.008: continue
    text = ""
    if key:
        f, l = key
        if f not in lines:
            try: lines[f] = open(f).read().splitlines()
            except OSError: lines[f] = []
        text = lines[f][l - 1].strip()[:110] if l - 1 < len(lines[f]) else ""
    print("%5.1f%% smp %5.1f%% inst thr/inst %4.1f  %s:%s  %s" % (100 * v[0] / tot, 100 * v[1] / toti, v[2] / max(v[1], 1), key[0].split("/")[-1] if key else "?", key[1] if key else "", text))
