import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
n=int(sys.argv[2]) if len(sys.argv)>2 else 40
hdr=rows[1]
idx={h:i for i,h in enumerate(hdr)}
data=[r for r in rows[2:] if len(r)==len(hdr)]
S=idx['# Samples']; I=idx['Instructions Executed']; src=idx['Source']
def iv(x):
    try: return int(x)
    except: return 0
tot=sum(iv(r[S]) for r in data)
print('total samples',tot,'total instr',sum(iv(r[I]) for r in data), 'rows', len(data))
stalls=[h for h in hdr if h.startswith('stall') and 'Not Issued' not in h]
agg={h:sum(iv(r[idx[h]]) for r in data) for h in stalls}
print(sorted(agg.items(), key=lambda kv:-kv[1])[:8])
top=sorted(data,key=lambda r:-iv(r[S]))[:n]
for r in top:
    st=sorted(((iv(r[idx[h]]),h) for h in stalls),reverse=True)[:2]
    print(f"{iv(r[S]):7d} {100*iv(r[S])/max(tot,1):5.1f}% ex={r[I]:>9s} {r[idx['Address']][-5:]} {r[src][:70]:70s} {st}")
