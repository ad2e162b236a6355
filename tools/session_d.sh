#!/bin/bash
tools/session_c.sh 2>&1 | tee gpurun_out/e3_session.log
python tools/ab2.py elfel.jl_b200/libelfelgpu.so tools/ab/lib_m3b192.so tools/ab/lib_m4b160.so tools/ab/lib_m3b224.so 2>&1 | tee gpurun_out/e3_ab.log
