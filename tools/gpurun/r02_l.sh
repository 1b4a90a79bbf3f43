#!/bin/bash
mkdir -p gpurun_out
SB200_TRACE=1 timeout 300 python bench.py --no-cpu-baseline --steps 8 --warmup 5 --visual-threshold max > /dev/null 2> gpurun_out/r02l_trace.err
grep "grows\|waits\|predict:" gpurun_out/r02l_trace.err | grep -v " 0 -> " | cut -c1-170 | tail -50
