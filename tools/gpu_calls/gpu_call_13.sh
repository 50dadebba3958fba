#!/bin/bash
# round 2, call 13: frames per launch for the large shapes (C4 batch 2 / 4, C5 batch 2), slots 4 vs 8; tests after the b200s_create change
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t13.log; tail -3 gpurun_out/r2_t13.log
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --table '' --config ${CFG} > gpurun_out/r2_b13_${CFG}_${name}.json 2> gpurun_out/r2_b13_${CFG}_${name}.err
}
CFG=C4; run b1 B200S_BENCH_BATCH=1; run b2 B200S_BENCH_BATCH=2; run b4 B200S_BENCH_BATCH=4; run b2s8 B200S_BENCH_BATCH=2 B200S_BENCH_SLOTS=8 B200S_BENCH_FRAMES=32; run b1s8 B200S_BENCH_BATCH=1 B200S_BENCH_SLOTS=8
CFG=C5; run b1 B200S_BENCH_BATCH=1; run b2 B200S_BENCH_BATCH=2
CFG=C3; run b8 B200S_BENCH_BATCH=8; run b16 B200S_BENCH_BATCH=16 B200S_BENCH_FRAMES=64
for f in gpurun_out/r2_b13_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    for k,v in d["configs"].items(): print(sys.argv[1].split("r2_b13_")[1],k, "B",v["batch"],"S",v["slots"],"fps", round(v["frames_per_s"]), "e2e", round(v["e2e_frames_per_s"]), "matcher", round(v.get("matcher_us",0),1), {a:round(x,1) for a,x in v.get("stage_us",{}).items()}, v["parity_checked"]["mismatches"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
