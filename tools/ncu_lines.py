#!/usr/bin/env python
"""Aggregate an ncu --page source (SASS) CSV by CUDA source line using nvdisasm line info.
usage: ncu_lines.py <report.ncu-rep> <kernel-regex> <cubin> <mangled-substring> [top]"""
import collections, csv, re, subprocess, sys
rep, kre, cubin, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
start = None
for i, l in enumerate(sass):
    if l.startswith(".text.") and mangled in l and l.rstrip().endswith(":"):
        start = i; break
assert start is not None, "kernel not found in cubin"
cur = None; a2l = {}
for l in sass[start + 1:]:
    if l.startswith(".text.") or l.startswith("\t.section"):
        if a2l: break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        inl = re.search(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*)", l)
    if m: a2l[int(m.group(1), 16)] = (cur, m.group(2))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hdr = next(r for r in rows if r and r[0] == "Address")
hi = rows.index(hdr)
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None; agg = collections.Counter(); samp = collections.Counter(); ops = collections.defaultdict(collections.Counter)
for r in rows[hi + 1:]:
    if len(r) <= isamp: continue
    try: a = int(r[ia], 16)
    except ValueError: continue
    if base is None: base = a
    info = a2l.get(a - base)
    key = info[0] if info and info[0] else ("?", 0)
    agg[key] += int(r[ii] or 0); samp[key] += int(r[isamp] or 0)
tot, ts = sum(agg.values()), sum(samp.values())
print("total instr", tot, "samples", ts)
srcs = {}
for k, v in samp.most_common(top):
    f = k[0]
    if f not in srcs:
        try: srcs[f] = open("/root/repo/gnn-formation-control_b200/csrc/" + f).read().split("\n")
        except Exception: srcs[f] = []
    text = srcs[f][k[1] - 1].strip()[:95] if 0 < k[1] <= len(srcs[f]) else ""
    print("%-22s %4d  samples %5.1f%%  instr %5.1f%%  %s" % (f, k[1], 100 * v / max(ts, 1), 100 * agg[k] / max(tot, 1), text))
