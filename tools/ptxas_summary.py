"""Summarise `nvcc -Xptxas -v` output: kernel, registers, spills, smem.  Usage: python tools/ptxas_summary.py file.cu [filter]"""
import re, subprocess, sys
src = sys.argv[1]; flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
                      "-Xptxas", "-v", "-c", src, "-o", "/tmp/_ptxas.o"], capture_output=True, text=True).stderr
name = None
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name); name = re.sub(r"\(.*", "", name)
        spill = ""
    m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m: spill = f"spill {m.group(1)}/{m.group(2)}"
    m = re.search(r"Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", line)
    if m and name and flt in name:
        print(f"{name:90s} regs {m.group(1):>4s} {spill} {line.split('Used')[1][:80]}")
