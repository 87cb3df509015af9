#!/usr/bin/env python3
"""Aggregate an ncu source page (SASS) by CUDA source line using nvdisasm -g line info of the same build.
usage: ncu_lines.py <report.ncu-rep> <dis.txt> <kernel mangled substring> [top]"""
import csv, re, sys, collections, subprocess, io
rep, dis, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
base = int(data[0][ix['Address']], 16)
lines = open(dis).read().split('\n')
start = [i for i, l in enumerate(lines) if ('text.' + kern) in l][0]
off2line = {}; cur = None
for l in lines[start + 1:]:
    if l.startswith('//-----') and 'text.' in l: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s', l)
    if m and cur: off2line[int(m.group(1), 16)] = cur
reasons = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = collections.defaultdict(lambda: collections.Counter())
for r in data:
    ln = off2line.get(int(r[ix['Address']], 16) - base, ('?', 0))
    a = agg[ln]
    a['samples'] += int(r[ix['# Samples']] or 0)
    a['inst'] += int(r[ix['Instructions Executed']] or 0)
    for h in reasons: a[h] += int(r[ix[h]] or 0)
tot = sum(a['samples'] for a in agg.values()); toti = sum(a['inst'] for a in agg.values())
print(f"total samples {tot}, warp instructions {toti}")
src = {}
def text(f, ln):
    if f not in src:
        try: src[f] = open('qwen3_tts_cuda_graphs_b200/csrc/' + f).read().split('\n')
        except OSError: src[f] = None
    return src[f][ln - 1].strip()[:70] if src[f] and 0 < ln <= len(src[f]) else f
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]['samples'])[:top]:
    rs = sorted(((a[h], h[6:]) for h in reasons if a[h]), reverse=True)[:3]
    print(f"{a['samples']:6d} {100*a['samples']/tot:5.1f}% inst {a['inst']:9d} {100*a['inst']/toti:4.1f}% L{ln:<5d} {text(f, ln):70s} " + " ".join(f"{n}:{c}" for c, n in rs))
