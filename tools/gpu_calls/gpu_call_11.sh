#!/bin/bash
# round 2, call 11: pack kernel A/B (rows prefetched at 40 registers vs loads inside the row loop), HX off again
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for sfx in "" _np; do
  B200S_LIB_SUFFIX=$sfx timeout 300 python bench.py --config C4 --steps 10 --warmup 3 --no-cpu --table 'C3' > gpurun_out/r2_b11${sfx}.json 2> gpurun_out/r2_b11${sfx}.err
done
timeout 600 python -m pytest tests -m gpu -x -q -k "pointcloud or fused_chain or colour or batched or bench_configuration_parity" > gpurun_out/r2_t11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t11.log; tail -3 gpurun_out/r2_t11.log
for f in gpurun_out/r2_b11*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    for k,v in d["configs"].items(): print(sys.argv[1],k, "fps", round(v["frames_per_s"]), "e2e", round(v["e2e_frames_per_s"]), "matcher", round(v.get("matcher_us",0),1), {a:round(x,1) for a,x in v.get("stage_us",{}).items()}, v["parity_checked"]["mismatches"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
