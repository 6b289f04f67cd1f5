#!/usr/bin/env bash
# round 2, GPU call ar: smoke() with the steady-state and batch checks
set -u
cd "$(dirname "$0")/.."
timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
