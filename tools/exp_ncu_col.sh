#!/bin/bash
# one full ncu capture (with source-level sampling) of the column-walk conv at the bench shape
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_col -s 2 -c 1 -f -o gpurun_out/r02_prof_col python tools/gpu_col_probe.py > gpurun_out/r02_prof_col.log 2>&1
tail -3 gpurun_out/r02_prof_col.log
