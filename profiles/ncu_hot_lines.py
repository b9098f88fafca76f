"""Usage: python profiles/ncu_hot_lines.py <report.ncu-rep> <kernel regex> <top N>  — instruction / stall share per CUDA source line."""
import csv, sys, subprocess, collections
rep, kern, topn = sys.argv[1], sys.argv[2], int(sys.argv[3])
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass','--kernel-name','regex:'+kern],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur=None; agg=collections.Counter(); thr=collections.Counter(); samp=collections.Counter()
for r in rows:
    if len(r)>=2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)<10 or r[0] in ('','Line No'): continue
    try: ln=int(r[0]); ie=int(r[7]); te=int(r[8]); sm=int(r[6])
    except: continue
    k=(cur,ln,r[1].strip()[:95]); agg[k]+=ie; thr[k]+=te; samp[k]+=sm
tot=sum(agg.values()); ts=sum(samp.values())
print(tot, ts)
for k,v in agg.most_common(topn): print(f"{100*v/tot:5.1f}% samp={100*samp[k]/max(ts,1):5.1f}% thr={thr[k]/max(v,1):5.1f} {k[0][5:]}:{k[1]} {k[2]}")
