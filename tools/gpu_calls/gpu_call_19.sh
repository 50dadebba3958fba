#!/bin/bash
# round 2, call 19: per-disparity table in the pack kernel -- parity of every chain test, A/B against the arithmetic path
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "not vh_kernel and not lr_check_border and not speckle" > gpurun_out/r2_t19.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t19.log; tail -3 gpurun_out/r2_t19.log
timeout 600 python tools/fuzz_chain.py 60 919 > gpurun_out/r2_fuzz19_chain.log 2>&1; tail -1 gpurun_out/r2_fuzz19_chain.log
for lut in 1 0; do
  B200S_PACK_LUT=$lut timeout 300 python bench.py --config C4 --steps 10 --warmup 4 --no-cpu --table '' > gpurun_out/r2_b19_lut$lut.json 2> gpurun_out/r2_b19_lut$lut.err
done
timeout 300 python bench.py --config C1 --steps 10 --warmup 4 --no-cpu --table '' > gpurun_out/r2_b19_c1.json 2> gpurun_out/r2_b19_c1.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_b19_lut1.json","gpurun_out/r2_b19_lut0.json","gpurun_out/r2_b19_c1.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "parity", d["parity_checked"]["frames"], d["parity_checked"]["mismatches"])
        for k,v in d["configs"].items(): print("   ",k, round(v["frames_per_s"]), {a:round(x,1) for a,x in v.get("stage_us",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
