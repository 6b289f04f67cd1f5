#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
GROUPS=1 D3H_DEBUG_REUSE=1 timeout 300 python profiles/step_timeline.py 2>&1 | grep "round i0" | sort | uniq -c | head -12
