import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i,r in enumerate(rows) if r and r[0]=="ID")
hdr = rows[hi]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.defaultdict(lambda:[0,0.0]); seq=[]
for r in rows[hi+1:]:
    if len(r) <= vi: continue
    name = re.sub(r"\(.*","",r[ki]).replace("<unnamed>::","").replace("void ",""); t = float(r[vi].replace(",","") or 0)
    agg[name][0]+=1; agg[name][1]+=t; seq.append((name,t/1e3))
tot = sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 10]:
    print(f"{v[1]/1e6:9.3f} ms {100*v[1]/tot:5.1f}% n={v[0]:4d} {k[:70]}")
print("total ms", tot/1e6)
tc=[round(t) for n,t in seq if n.startswith("tc_kernel")]
print("tc us first:", tc[:14]); print("tc us last 73:", tc[-73:])
print("attn us:", [(n[:6],round(t)) for n,t in seq if "attn" in n or n.startswith("flash")][:9])
print("ws us:", [round(t) for n,t in seq if n.startswith("watershed")])
