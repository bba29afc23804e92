"""Per-kernel SASS opcode histogram of libm3d_b200.so (cuobjdump -sass), with the mnemonics that would prove tensor-core /
TMA paths called out.  Usage: python tools/sass_opcodes.py [path/to/lib.so] > profiles/r2_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
lib = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "merfish3d-analysis_b200" / "libm3d_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = {}
kernels = collections.OrderedDict()
for part in txt.split("Function : ")[1:]:
    name = part.split("\n", 1)[0].strip()
    ops = collections.Counter()
    for line in part.split("\n"):
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            ops[m.group(1)] += 1
    kernels[name] = ops
try:
    names = subprocess.run(["c++filt"] + list(kernels), capture_output=True, text=True, check=True).stdout.split("\n")
    demangle = dict(zip(kernels, names))
except Exception:
    pass
WATCH = ("HMMA", "DMMA", "IMMA", "UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "REDUX",
         "DADD", "DMUL", "DFMA", "FFMA", "MUFU", "ATOMS", "ATOMG", "RED", "SHFL", "VOTE", "LDG", "STG", "LDS", "STS")
total = collections.Counter()
print(f"SASS opcode histogram of {Path(lib).name} (sm_100a), {len(kernels)} kernels; static instruction counts\n")
print("Mnemonics that would show Blackwell tensor-core / TMA paths: UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld/st), UTMALDG / UTMASTG /")
print("UBLKCP (cp.async.bulk[.tensor]).  HMMA = warp-level mma.sync (the dense-candidate marking).  See DESIGN.md section 4 for why the")
print("streaming / sparse kernels of this path use neither.\n")
for name, ops in kernels.items():
    total.update(ops)
    short = re.sub(r"\(anonymous namespace\)::", "", demangle.get(name, name))
    short = re.sub(r"\(.*", "", short)
    n = sum(ops.values())
    watched = ", ".join(f"{k} {ops[k]}" for k in WATCH if ops.get(k))
    top = ", ".join(f"{k} {v}" for k, v in ops.most_common(6))
    print(f"{short}\n    {n} instructions; top: {top}\n    watched: {watched}\n")
print("library total:", ", ".join(f"{k} {total[k]}" for k in WATCH if total.get(k)))
print("UTC*MMA:", sum(v for k, v in total.items() if k.startswith("UTC")), " LDTM/STTM:", total.get("LDTM", 0) + total.get("STTM", 0),
      " UTMALDG/UTMASTG/UBLKCP:", total.get("UTMALDG", 0) + total.get("UTMASTG", 0) + total.get("UBLKCP", 0))
