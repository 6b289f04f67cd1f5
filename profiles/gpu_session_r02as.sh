#!/usr/bin/env bash
# round 2, GPU call as: compute-sanitizer on the final kernels (shared corner ids, shared prepare, register caps)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02as
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/sanitizer_case.py > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/${T}_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/sanitizer_case.py > gpurun_out/${T}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/${T}_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 3 python profiles/sanitizer_case.py > gpurun_out/${T}_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -2 gpurun_out/${T}_synccheck.log
