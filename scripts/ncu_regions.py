import csv, io, subprocess, sys
from collections import defaultdict
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
    rows=list(csv.reader(io.StringIO(out)))
    print(rows[0][1][:80])
    hdr=rows[1]
    si,samp,inst=hdr.index("Source"),hdr.index("# Samples"),hdr.index("Instructions Executed")
    ops=[]
    for r in rows[2:]:
        src=r[si].split()
        op=(src[1] if src[0].startswith("@") else src[0]).split(".")[0]
        ops.append((op,float(r[samp] or 0),float(r[inst] or 0),r[si]))
    popc=[i for i,o in enumerate(ops) if o[0]=="POPC"]
    clusters=[]; start=popc[0]; prev=popc[0]
    for i in popc[1:]:
        if i-prev>150: clusters.append((start,prev)); start=i
        prev=i
    clusters.append((start,prev))
    tot_s=sum(o[1] for o in ops); tot_i=sum(o[2] for o in ops)
    print(" total samples",tot_s,"warp instrs %.1fM"%(tot_i/1e6),"POPC clusters",clusters, "static", len(ops))
    bounds=[0]+[c for cl in clusters for c in cl]+[len(ops)]
    names=["before"]+sum(([f"popc{j}",f"after{j}"] for j in range(len(clusters))),[])
    for name,(lo,hi) in zip(names,zip(bounds[:-1],bounds[1:])):
        s=sum(x[1] for x in ops[lo:hi+1]); n=sum(x[2] for x in ops[lo:hi+1])
        print(f"   {name:8s} lines {lo}-{hi}: samples {100*s/tot_s:5.1f}%  instrs {100*n/tot_i:5.1f}% ({n/1e6:.1f}M)")
    # opcode histogram in the last region (epilogue)
    lo=bounds[-2]
    agg=defaultdict(lambda:[0,0])
    for o in ops[lo:]:
        agg[o[0]][0]+=o[1]; agg[o[0]][1]+=o[2]
    print("   epilogue opcodes:", ", ".join(f"{k} {v[1]/1e6:.1f}M/{100*v[0]/tot_s:.1f}%" for k,v in sorted(agg.items(),key=lambda kv:-kv[1][1])[:16]))
