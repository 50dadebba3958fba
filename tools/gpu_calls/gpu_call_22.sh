#!/bin/bash
# round 2, call 22: rectify quad form with saturated map entries, cheaper validity mask and a two-row Sobel phase
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "not vh_kernel and not lr_check_border and not speckle and not tall_band" > gpurun_out/r2_t22.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t22.log; tail -3 gpurun_out/r2_t22.log
timeout 600 python tools/fuzz_chain.py 80 2222 > gpurun_out/r2_fuzz22_chain.log 2>&1; tail -1 gpurun_out/r2_fuzz22_chain.log
timeout 300 python bench.py --config C4 --steps 10 --warmup 4 --no-cpu --table '' > gpurun_out/r2_b22_c4.json 2> gpurun_out/r2_b22_c4.err
B200S_BENCH_RECT_FLY=1 timeout 300 python bench.py --config C4 --steps 4 --warmup 3 --no-cpu --table '' > gpurun_out/r2_b22_fly.json 2> gpurun_out/r2_b22_fly.err
timeout 300 python bench.py --config C3 --steps 6 --warmup 3 --no-cpu --table '' > gpurun_out/r2_b22_c3.json 2> gpurun_out/r2_b22_c3.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_b22_c4.json","gpurun_out/r2_b22_fly.json","gpurun_out/r2_b22_c3.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["frac_of_copy_ceiling"],3), "parity", d["parity_checked"]["frames"], d["parity_checked"]["mismatches"])
        for k,v in d["configs"].items(): print("   ",k, round(v["frames_per_s"]), {a:round(x,1) for a,x in v.get("stage_us",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
