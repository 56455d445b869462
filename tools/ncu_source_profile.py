"""ncu_source_profile.py <report.ncu-rep> <launch index> [exec-count [top N]] - instruction classes of one profiled launch
(ncu --import-source on): how many SASS instructions run how often, where the stall samples fall."""
import csv, collections, sys, subprocess
rep, skip = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--launch-skip",skip,"--launch-count","1"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
print(rows[0][1][:100])
hdr=rows[1]; idx={h:i for i,h in enumerate(hdr)}
data=[r for r in rows[2:] if len(r) > idx['Instructions Executed'] and r[idx['Instructions Executed']].isdigit()]
cls=collections.Counter(); cnt=collections.Counter(); smp=collections.Counter()
for r in data:
    n=int(r[idx['Instructions Executed']]); cls[n]+=n; cnt[n]+=1; smp[n]+=int(r[idx['# Samples']])
tot=sum(cls.values()); ts=sum(smp.values())
print("total inst %.2fM samples %d"%(tot/1e6, ts))
for n,v in sorted(cls.items(), key=lambda kv:-kv[1])[:8]:
    print(f"exec/instr {n:9d}  static {cnt[n]:5d}  total {v/1e6:8.2f}M {100*v/tot:5.1f}%  samples {smp[n]:6d} {100*smp[n]/ts:5.1f}%")
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg={h:sum(int(r[idx[h]]) for r in data if len(r) > idx[h] and r[idx[h]].isdigit()) for h in stalls}
print(sorted(agg.items(), key=lambda kv:-kv[1])[:8])
if len(sys.argv)>3:
    want=int(sys.argv[3])
    for r in data:
        if int(r[idx['Instructions Executed']])==want:
            print(r[idx['Address']][-5:], r[idx['# Samples']].rjust(4), r[idx['Source']].strip()[:90])
if len(sys.argv)>4 and sys.argv[4]=="top":
    top=sorted(data,key=lambda r:-int(r[idx['# Samples']]))[:int(sys.argv[5]) if len(sys.argv)>5 else 25]
    for r in top:
        st={h:int(r[idx[h]]) for h in stalls if int(r[idx[h]])>0}
        st=sorted(st.items(), key=lambda kv:-kv[1])[:2]
        print(r[idx['# Samples']].rjust(5), r[idx['Instructions Executed']].rjust(8), r[idx['Source']].strip()[:75], st)
