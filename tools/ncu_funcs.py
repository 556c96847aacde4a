#!/usr/bin/env python3
"""Aggregates an .ncu-rep's per-line instruction counts and stall samples by function of one source file."""
import csv, subprocess, re, sys
rep, srcfile = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur=None; data=[]
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split("/")[-1]; continue
    if r[0]=="Line No": ci=r.index("Instructions Executed"); ti=r.index("Thread Instructions Executed"); si=r.index("# Samples"); continue
    if len(r)>2 and r[2]=="-" and r[0].isdigit():
        try: data.append((cur,int(r[0]),int(r[ci]),int(r[ti]),int(r[si])))
        except ValueError: pass
src=open(srcfile).read().split('\n')
base=srcfile.split('/')[-1]
bounds=[(i+1,l) for i,l in enumerate(src) if re.match(r'^(__device__|__global__)',l)]
def fn(line):
    name='?'
    for b,l in bounds:
        if b<=line: name=l[:80]
        else: break
    return name
agg={}
tot=sum(d[2] for d in data); tots=sum(d[4] for d in data)
for f,l,n,t,s in data:
    key = fn(l) if f==base else f
    a=agg.setdefault(key,[0,0,0]); a[0]+=n; a[1]+=t; a[2]+=s
print(f"total warp instructions {tot/1e6:.1f}M, samples {tots}")
for k,(n,t,s) in sorted(agg.items(), key=lambda x:-x[1][0])[:int(sys.argv[3]) if len(sys.argv)>3 else 25]:
    print(f"{n/tot*100:5.1f}% {n/1e6:7.1f}M lanes {t/max(n,1):4.1f} smp {s/tots*100:4.1f}%  {k}")
