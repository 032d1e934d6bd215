#!/bin/bash
mkdir -p gpurun_out
compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/j46_memcheck.log 2>&1; tail -4 gpurun_out/j46_memcheck.log
compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/j46_racecheck.log 2>&1; tail -4 gpurun_out/j46_racecheck.log
compute-sanitizer --tool synccheck python tools/sanitize_small.py > gpurun_out/j46_synccheck.log 2>&1; tail -3 gpurun_out/j46_synccheck.log
