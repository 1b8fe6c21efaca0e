import csv, collections, re, sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if r and r[0]=="Address"]
hdr=rows[hi[0]]; col={h:i for i,h in enumerate(hdr)}
end=hi[1]-1 if len(hi)>1 else len(rows)
data=[r for r in rows[hi[0]+1:end] if len(r)>=len(hdr) and r[0]!="Address"]
ops=collections.Counter(); tot=0; thr=collections.Counter()
def I(x):
    try: return int(x)
    except: return 0
for r in data:
    src=r[col["Source"]]
    m=re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)",src)
    op=m.group(2) if m else src
    base=op.split('.')[0]
    if base=="IMAD": base="IMAD.WIDE" if ".WIDE" in op else ("IMAD.HI" if ".HI" in op else ("IMAD.MOV" if ".MOV" in op else "IMAD"))
    n=I(r[col["Instructions Executed"]]); ops[base]+=n; tot+=n; thr[base]+=I(r[col["Thread Instructions Executed"]])
print("total warp instr:",tot,"static:",len(data))
for k,v in ops.most_common(22): print("%-12s %9d %5.1f%%  avg thr %.1f"%(k,v,100*v/tot, thr[k]/max(1,v)))
cum=0
print("--- markers (cumulative executed warp-instr up to marker)")
for r in data:
    n=I(r[col["Instructions Executed"]]); cum+=n; s=r[col["Source"]]
    if re.search(r"BAR\.|SYNCS|UBLKCP|EXIT|VOTE|ATOMS",s): print(r[col["Address"]][-5:], s.strip()[:60], n, cum)
