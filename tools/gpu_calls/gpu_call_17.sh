#!/bin/bash
# round 2, call 17: border strips inside the matcher's edge tiles -- parity, fuzz, and the timing with / without
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "lr_check_border or batched_frames or bench_configuration or vh_kernel or ragged or tall_band" > gpurun_out/r2_t17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t17.log; tail -3 gpurun_out/r2_t17.log
timeout 600 python tools/fuzz_parity.py 300 777 > gpurun_out/r2_fuzz17_parity.log 2>&1; tail -1 gpurun_out/r2_fuzz17_parity.log
for c in C4 C1 C2 C3; do
  timeout 120 python tools/time_bm.py $c 20 4 2>&1 | tail -1
done
for c in C4 C2 C1; do
  DISP12=1 timeout 120 python tools/time_bm.py $c 20 4 2>&1 | tail -1
  DISP12=1 B200S_VH_EDGES=0 timeout 120 python tools/time_bm.py $c 20 4 2>&1 | tail -1
done
DISP12=1 CAP=63 timeout 120 python tools/time_bm.py C4 20 4 2>&1 | tail -1
timeout 300 python bench.py --config C4r --steps 6 --warmup 3 --no-cpu --table '' > gpurun_out/r2_b17_c4r.json 2> gpurun_out/r2_b17_c4r.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_b17_c4r.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "parity", d["parity_checked"]["frames"], d["parity_checked"]["mismatches"])
    except Exception as e: print(f, "ERR", e)
PY
