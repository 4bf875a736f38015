"""Executed warp instructions per basic block (runs of SASS rows with the same execution count) of an ncu source-page csv:
python tools/sass_blocks.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
I, S, A, SRC = idx['Instructions Executed'], idx['# Samples'], idx['Address'], idx['Source']
def iv(x):
    try: return int(x)
    except ValueError: return 0
tot = sum(iv(r[I]) for r in data)
tots = sum(iv(r[S]) for r in data)
seg, cur = [], None
for i, r in enumerate(data):
    ex = iv(r[I])
    if cur is None or ex != cur[2]:
        if cur: seg.append(cur)
        cur = [i, i, ex, ex, iv(r[S])]
    else:
        cur[1] = i; cur[3] += ex; cur[4] += iv(r[S])
seg.append(cur)
print(f"total {tot} warp instructions, {tots} samples")
for s in sorted(seg, key=lambda s: -s[3])[:top]:
    r = data[s[0]]
    print(f"{r[A][-5:]} +{s[1]-s[0]+1:4d} instr  x{s[2]:>9d} = {s[3]:>10d} ({100*s[3]/tot:4.1f}% instr, {100*s[4]/max(tots,1):4.1f}% samples)  {r[SRC][:50]}")
