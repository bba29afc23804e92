"""Top stalled SASS lines of one kernel from an .ncu-rep.  Usage: python tools/ncu_hot.py rep kernel_regex [N]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]; N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]; ix = {h: i for i, h in enumerate(H)}
body = [r for r in rows[hdr + 1:] if len(r) == len(H) and r[0] != "Address"]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
inst = sum(int(r[ix["Instructions Executed"]] or 0) for r in body)
print(f"kernel rows={len(body)} samples={tot} warp-instr={inst}")
stall_cols = [h for h in H if h.startswith("stall_")]
top = sorted(enumerate(body), key=lambda t: -int(t[1][ix["# Samples"]] or 0))[:N]
for i, r in sorted(top):
    st = sorted(((int(r[ix[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {int(r[ix['# Samples']]):7d} {100*int(r[ix['# Samples']])/max(tot,1):5.1f}% exec={r[ix['Instructions Executed']]:>9s} {r[ix['Source']].strip()[:70]:70s} {st}")
