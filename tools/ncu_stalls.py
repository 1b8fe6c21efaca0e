import csv, sys, collections
f=sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 40
rows=list(csv.reader(open(f)))
# find header row
hi=[i for i,r in enumerate(rows) if r and r[0]=="Address"][0]
hdr=rows[hi]; data=rows[hi+1:]
col={h:i for i,h in enumerate(hdr)}
stall_cols=[h for h in hdr if h.startswith("stall_")]
tot=collections.Counter(); total_samples=0
recs=[]
for r in data:
    if len(r)<len(hdr): continue
    try: s=int(r[col["# Samples"]] or 0)
    except: s=0
    total_samples+=s
    for h in stall_cols:
        try: tot[h]+=int(r[col[h]] or 0)
        except: pass
    recs.append((s,r))
print("total samples",total_samples)
print("stall totals:",[(k,v) for k,v in tot.most_common(12)])
recs.sort(key=lambda x:-x[0])
for s,r in recs[:top]:
    st=sorted([(int(r[col[h]] or 0),h[6:]) for h in stall_cols if (r[col[h]] or "0")!="0"],reverse=True)[:3]
    print("%6d %5.1f%%  %-8s %-70s %s"%(s,100*s/max(1,total_samples),r[col["Address"]][-5:],r[col["Source"]][:70],st))
