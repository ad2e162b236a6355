#!/bin/bash
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/ab2.py tools/ab/lib_prev.so elfel.jl_b200/libelfelgpu.so 2>&1 | tee gpurun_out/e17_ab.log
