#!/bin/bash
mkdir -p gpurun_out
SB200_TRACE=1 timeout 300 python bench.py --no-cpu-baseline --steps 8 --warmup 5 > /dev/null 2> gpurun_out/r02l_trace.err
grep "grows\|waits" gpurun_out/r02l_trace.err | cut -c1-170 | tail -60
