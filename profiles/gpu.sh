#!/usr/bin/env bash
# build the library from the current sources, then run a session script on the GPU box (the .so travels with the snapshot)
set -e
cd "$(dirname "$0")/.."
python d3human-code_b200/build.py --force > /dev/null 2>&1 || { echo "BUILD FAILED"; exit 1; }
exec gpurun --timeout "${2:-2400}" -- "bash $1"
