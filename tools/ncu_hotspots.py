"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by source line."""
import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
blocks=[]; cur=None
for r in rows:
    if r and r[0]=="File Path": cur={'file':r[1],'rows':[]}; blocks.append(cur)
    elif cur is not None: cur['rows'].append(r)
tot=collections.Counter(); stall=collections.Counter(); src={}
for b in blocks:
    hdr=None
    for r in b['rows']:
        if r and r[0]=="Line No": hdr=r; continue
        if hdr is None or len(r)<len(hdr): continue
        d=dict(zip(hdr,r))
        try: ln=int(d["Line No"])
        except: continue
        def num(k):
            try: return int(float(d.get(k,"0") or "0"))
            except: return 0
        key=(b['file'].split('/')[-1],ln)
        tot[key]+=num("Instructions Executed"); stall[key]+=num("Warp Stall Sampling (All Samples)")
        src[key]=r[1][:100]
T=sum(tot.values()); S=sum(stall.values())
n=int(sys.argv[2]) if len(sys.argv)>2 else 40
print("total inst",T,"samples",S)
print("--- top by instructions")
for k,v in tot.most_common(n): print(f"{k[0]}:{k[1]:4d} {100*v/T:5.1f}% inst {100*stall[k]/max(S,1):5.1f}% stall | {src[k]}")
print("--- top by stall samples")
for k,v in stall.most_common(n//2): print(f"{k[0]}:{k[1]:4d} {100*v/S:5.1f}% stall {100*tot[k]/T:5.1f}% inst | {src[k]}")
