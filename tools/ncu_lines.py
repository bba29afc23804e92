"""Per-source-line and per-function instruction / sample totals of one kernel from an .ncu-rep (needs --import-source on).
Usage: python tools/ncu_lines.py rep [N]   -- the source page is printed with CUDA lines heading their SASS rows."""
import csv, io, re, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur_file = None; H = None; ix = None
lines = defaultdict(lambda: [0, 0, ""])
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": H = r; ix = {h: i for i, h in enumerate(H)}; si = H.index("# Samples"); ii = H.index("Instructions Executed"); continue
    if H is None or len(r) != len(H): continue
    if r[0] != "":  # a CUDA source line heading its SASS rows: totals are on this row
        key = (cur_file, int(r[0]))
        lines[key][0] += int(r[ii] or 0); lines[key][1] += int(r[si] or 0); lines[key][2] = r[1].strip()
tot_i = sum(v[0] for v in lines.values()); tot_s = sum(v[1] for v in lines.values())
print(f"total warp instr {tot_i} samples {tot_s}")
for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:N]:
    print(f"{f}:{ln:4d} {100*v[0]/tot_i:5.1f}% instr {100*v[1]/max(tot_s,1):5.1f}% smp  {v[2][:110]}")
# by function of decode.cu (line ranges from the source file)
src = open("merfish3d-analysis_b200/csrc/decode.cu").read().splitlines()
starts = []
for n, l in enumerate(src, 1):
    m = re.match(r"^(?:__device__ __forceinline__|static|__global__|template).*?\b(\w+)\(", l) if not l.startswith(" ") else None
    if m and "template <" not in l: starts.append((n, m.group(1)))
    m2 = re.match(r"^(\w+)\(const T\* __restrict__ stack", l)
    if m2: starts.append((n, m2.group(1)))
def func_of(f, ln):
    if f != "decode.cu": return f
    name = "?"
    for s, nm in starts:
        if s <= ln: name = nm
    return name
by = defaultdict(lambda: [0, 0])
for (f, ln), v in lines.items():
    k = func_of(f, ln); by[k][0] += v[0]; by[k][1] += v[1]
print("---- by function")
for k, v in sorted(by.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:34s} {100*v[0]/tot_i:5.1f}% instr {100*v[1]/max(tot_s,1):5.1f}% samples")
# static code size (SASS rows) and stall_no_inst samples by function
cur_file = None; H = None; cur = None
size = defaultdict(int); noinst = defaultdict(int)
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": H = r; continue
    if H is None or len(r) != len(H): continue
    if r[0] != "": cur = func_of(cur_file, int(r[0])); continue
    size[cur] += 1
print("---- SASS instructions by function (static)")
for k, v in sorted(size.items(), key=lambda kv: -kv[1]): print(f"{k:34s} {v:6d}")
print("total", sum(size.values()))
