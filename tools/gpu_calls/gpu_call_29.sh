#!/bin/bash
# round 2, call 29: the tree as it ships -- smoke, every GPU test, the default bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke29.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke29.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t29.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t29.log; tail -3 gpurun_out/r2_t29.log
SECONDS=0; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_b29.json 2> gpurun_out/r2_b29.err; echo "bench rc=$? in ${SECONDS}s"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_b29.json").read().strip().splitlines()[-1])
print("fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["frac_of_copy_ceiling"],3), "parity", d["parity_checked"]["frames"], d["parity_checked"]["mismatches"], "roof", round(d["roofline"]["frac"],3), "cpu", (d.get("cpu_baseline") or {}).get("value"))
for k,v in d["configs"].items(): print("   ",k, round(v["frames_per_s"]), round(v["e2e_frames_per_s"]), round(v["e2e"]["frac_of_copy_ceiling"],3), round(v.get("matcher_us",0),1), round(v.get("matcher_tevals_per_s",0),3), {a:round(x,1) for a,x in v.get("stage_us",{}).items()})
PY
