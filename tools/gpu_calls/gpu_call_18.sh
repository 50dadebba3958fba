#!/bin/bash
# round 2, call 18: role-split strip kernel (C warps / winner warps over named barriers) -- parity, racecheck, timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "lr_check_border or batched_frames or bench_configuration or ragged or validate or lr_check or disp12" > gpurun_out/r2_t18.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t18.log; tail -3 gpurun_out/r2_t18.log
timeout 600 python tools/fuzz_parity.py 300 4711 > gpurun_out/r2_fuzz18_parity.log 2>&1; tail -1 gpurun_out/r2_fuzz18_parity.log
timeout 400 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests -m gpu -x -q -k "lr_check_border and 37 and case0" > gpurun_out/r2_race18.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r2_race18.log | tail -4
for c in C4 C2 C1; do
  DISP12=1 timeout 120 python tools/time_bm.py $c 20 4 2>&1 | tail -1
done
timeout 120 python tools/time_bm.py C4 20 4 2>&1 | tail -1
DISP12=1 timeout 120 python tools/time_bm.py C4 20 1 2>&1 | tail -1
timeout 300 python bench.py --config C4r --steps 6 --warmup 3 --no-cpu --table '' > gpurun_out/r2_b18_c4r.json 2> gpurun_out/r2_b18_c4r.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_b18_c4r.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "parity", d["parity_checked"]["frames"], d["parity_checked"]["mismatches"])
        for k,v in d["configs"].items(): print("   ",k, round(v["frames_per_s"]), {a:round(x,1) for a,x in v.get("stage_us",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
