#!/bin/bash
# round 2, call 28: table-path pack kernel with all row loads requested up front -- chain parity, random chains, C4 timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "pack_table or fused or batched or bench_configuration or graph or colour or process_pair" > gpurun_out/r2_t28.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t28.log; tail -3 gpurun_out/r2_t28.log
timeout 300 python tools/fuzz_chain.py 40 2828 > gpurun_out/r2_fuzz28_chain.log 2>&1; tail -1 gpurun_out/r2_fuzz28_chain.log
timeout 200 python bench.py --config C4 --steps 10 --warmup 4 --no-cpu --table '' > gpurun_out/r2_b28.json 2> gpurun_out/r2_b28.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_b28.json").read().strip().splitlines()[-1])
print("fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "parity", d["parity_checked"]["frames"], d["parity_checked"]["mismatches"])
for k,v in d["configs"].items(): print("   ",k, round(v["frames_per_s"]), {a:round(x,1) for a,x in v.get("stage_us",{}).items()})
PY
