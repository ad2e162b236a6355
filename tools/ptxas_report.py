#!/usr/bin/env python
"""Build the library with -Xptxas -v and list registers / spills / stack of the numeric kernels.
usage: python tools/ptxas_report.py [substring filter]   (EFG_LIB / EFG_NVCC_EXTRA respected; REBUILDS the library)
       python tools/ptxas_report.py --log build.log [substring filter]   (parses the stderr of an earlier -Xptxas -v build)"""
import os, re, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elfel_jl_b200 import _lib
if len(sys.argv) > 2 and sys.argv[1] == "--log":
    lines = open(sys.argv[2]).read().splitlines()
    del sys.argv[1:3]
else:
    nvcc = "/usr/local/cuda/bin/nvcc"
    extra = os.environ.get("EFG_NVCC_EXTRA", "").split()
    cmd = [nvcc] + _lib.NVCC_FLAGS + extra + ["-Xptxas", "-v", "-o", _lib.SO_PATH, os.path.join(_lib.CSRC, "elfel_gpu.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        print(r.stderr[-3000:]); sys.exit(1)
    lines = r.stderr.splitlines()
filt = sys.argv[1] if len(sys.argv) > 1 else "k_tl_numeric"
dem = {}
for i, l in enumerate(lines):
    m = re.search(r"Compiling entry function '(\S+)'", l)
    if not m or filt not in m.group(1):
        continue
    name = m.group(1)
    info = " ".join(lines[i + 1:i + 5])
    regs = re.search(r"Used (\d+) registers", info)
    sp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", info)
    short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.split("(")[0].replace("void ", "")
    print(f"{short:75s} regs {regs.group(1) if regs else '?':>4s}  stack {sp.group(1):>4s}  spill st/ld {sp.group(2)}/{sp.group(3)}")
