#!/usr/bin/env bash
# round 2, GPU call af: timeline of one bench step (torch profiler)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python profiles/step_timeline.py > gpurun_out/r02af_timeline.txt 2>&1
head -3 gpurun_out/r02af_timeline.txt; tail -2 gpurun_out/r02af_timeline.txt
D3H_FUSE_FRAMES=0 timeout 300 python profiles/step_timeline.py > gpurun_out/r02af_timeline_lanes.txt 2>&1
head -3 gpurun_out/r02af_timeline_lanes.txt; tail -2 gpurun_out/r02af_timeline_lanes.txt
