import csv,subprocess,io,re,collections,sys
rep=sys.argv[1]
def fnmap(path):
    src=open(path).read().split("\n"); fn_at={}; cur="?"
    for i,l in enumerate(src,1):
        m2=re.search(r'(?:NB2_HD|__global__|static inline|static)\s+(?:[\w<>:,\s\*&]+?)\s+(\w+)\(',l)
        if m2 and not l.startswith("  "): cur=m2.group(1)
        fn_at[i]=cur
    return fn_at
maps={}
for kern in sys.argv[2:]:
    out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass","--kernel-name","regex:"+kern],capture_output=True,text=True).stdout
    rows=list(csv.reader(io.StringIO(out)))
    cur_file=None; hdr=None; agg=collections.Counter(); inst=collections.Counter(); tot=0; lines=[]
    for r in rows:
        if not r: continue
        if r[0]=="File Path": cur_file=r[1]; continue
        if r[0]=="Line No": hdr=r; iS=hdr.index("# Samples"); iI=hdr.index("Instructions Executed"); continue
        if hdr is None or not r[0].isdigit(): continue
        if cur_file not in maps:
            try: maps[cur_file]=fnmap(cur_file)
            except Exception: maps[cur_file]={}
        fn=maps[cur_file].get(int(r[0]),"?")
        key=cur_file.split("/")[-1]+":"+fn
        try: s_=int(r[iS]); ie=int(r[iI])
        except ValueError: continue
        agg[key]+=s_; inst[key]+=ie; tot+=s_
        lines.append((s_,cur_file.split("/")[-1],int(r[0]),r[1][:100]))
    print(kern,"total samples",tot)
    for k,v in agg.most_common(24): print("  %-45s %6d  %5.1f%%  inst %d"%(k,v,100*v/max(tot,1),inst[k]))
    print("  top lines:")
    for s_,f,ln,t in sorted(lines,reverse=True)[:14]: print("   %5d %s:%d %s"%(s_,f,ln,t))
