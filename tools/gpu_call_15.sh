#!/bin/bash
# round 2, call 15: rectify kernel with the interior fast path: parity + bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t15.log; tail -3 gpurun_out/r2_t15.log
timeout 600 python tools/fuzz_chain.py 40 5 > gpurun_out/r2_fuzzchain15.log 2>&1; tail -2 gpurun_out/r2_fuzzchain15.log
SECONDS=0; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_b15.json 2> gpurun_out/r2_b15.err; echo "bench rc=$? in ${SECONDS}s"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_b15.json").read().strip().splitlines()[-1])
print("fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "parity", d["parity_checked"]["frames"], d["parity_checked"]["mismatches"])
for k,v in d["configs"].items(): print("   ",k, round(v["frames_per_s"]), round(v["e2e_frames_per_s"]), round(v.get("matcher_us",0),1), {a:round(x,1) for a,x in v.get("stage_us",{}).items()})
PY
