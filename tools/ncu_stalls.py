"""Stall samples of one kernel by function and reason (source page of an .ncu-rep captured with --import-source on).
Usage: python tools/ncu_stalls.py rep"""
import csv, io, subprocess, re, sys
from collections import defaultdict
rep=sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
src = open("merfish3d-analysis_b200/csrc/decode.cu").read().splitlines()
starts=[]
for n,l in enumerate(src,1):
    m = re.match(r"^(?:__device__ __forceinline__|static|__global__|template).*?\b(\w+)\(", l) if not l.startswith(" ") else None
    if m and "template <" not in l: starts.append((n,m.group(1)))
    m2 = re.match(r"^(\w+)\(const T\* __restrict__ stack", l)
    if m2: starts.append((n,m2.group(1)))
def func_of(f,ln):
    if f!="decode.cu": return f
    name="?"
    for s,nm in starts:
        if s<=ln: name=nm
    return name
def num(x):
    try: return int(x)
    except ValueError: return 0
H=None;cur=None;cur_file=None
agg=defaultdict(lambda: defaultdict(int))
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0]=="File Path": cur_file=r[1].split("/")[-1]; continue
    if r[0]=="Line No": H=r; stall_cols=[(i,h) for i,h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]; si=H.index("# Samples"); continue
    if H is None or len(r)!=len(H): continue
    if r[0]!="": cur=func_of(cur_file,num(r[0])); continue
    for i,h in stall_cols:
        agg[cur][h]+=num(r[i])
    agg[cur]['n']+=num(r[si])
tot=sum(v['n'] for v in agg.values())
print("function  samples%  top stalls")
for k,v in sorted(agg.items(), key=lambda kv:-kv[1]['n']):
    st=sorted(((c,h) for h,c in v.items() if h!='n'),reverse=True)[:4]
    print(f"{k:32s} {100*v['n']/tot:5.1f}%  "+", ".join(f"{h[6:]}={100*c/tot:.1f}%" for c,h in st))
allst=defaultdict(int)
for v in agg.values():
    for h,c in v.items():
        if h!='n': allst[h]+=c
print({h[6:]:round(100*c/tot,1) for h,c in sorted(allst.items(), key=lambda kv:-kv[1])})
