#!/bin/bash
# round 2, call 7: wide form of bm_vh (preFilterCap 32..63): parity + timing against the bm_ws fallback; full GPU tests
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t7.log
tail -5 gpurun_out/r2_t7.log
timeout 600 python tools/fuzz_parity.py 150 77 > gpurun_out/r2_fuzz7.log 2>&1; tail -3 gpurun_out/r2_fuzz7.log
(CAP=63 timeout 120 python tools/time_bm.py C4 10 1; CAP=63 B200S_KERNEL=4 timeout 120 python tools/time_bm.py C4 10 1; CAP=40 timeout 120 python tools/time_bm.py C3 10 8; CAP=31 timeout 120 python tools/time_bm.py C4 10 1) > gpurun_out/r2_wide7.log 2>&1
cat gpurun_out/r2_wide7.log
