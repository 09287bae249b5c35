import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
which=int(sys.argv[2]) if len(sys.argv)>2 else 0
starts=[i for i,r in enumerate(rows) if r and r[0]=="Kernel Name"]
starts.append(len(rows))
s=starts[which]; e=starts[which+1]
print(rows[s][1])
hdr=rows[s+1]; data=[r for r in rows[s+2:e] if len(r)==len(hdr)]
ix={h:i for i,h in enumerate(hdr)}
tot=sum(int(r[ix["# Samples"]]) for r in data)
print("total samples",tot, "instructions", len(data))
stall_cols=[h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg={h:sum(int(r[ix[h]]) for r in data) for h in stall_cols}
for h,v in sorted(agg.items(), key=lambda x:-x[1])[:10]: print("  %-26s %7d %5.1f%%"%(h,v,100*v/tot))
cls=collections.Counter(); cnt=collections.Counter()
for r in data:
    t=r[ix["Source"]].split()
    op=t[1] if t[0].startswith("@") else t[0]
    cls[op]+=int(r[ix["# Samples"]]); cnt[op]+=int(r[ix["Instructions Executed"]])
for op,v in cls.most_common(16): print("  %-14s samples %7d %5.1f%%  executed %d"%(op,v,100*v/tot,cnt[op]))
print("top non-DMMA instructions:")
k=0
for r in sorted(data,key=lambda r:-int(r[ix["# Samples"]])):
    if 'DMMA' in r[ix["Source"]]: continue
    st={h:int(r[ix[h]]) for h in stall_cols if int(r[ix[h]])>0}
    top=sorted(st.items(), key=lambda x:-x[1])[:3]
    print("  %6s %-60s %s"%(r[ix["# Samples"]], r[ix["Source"]].strip()[:60], top))
    k+=1
    if k>22: break
