#!/bin/bash
# round 2, call 23: software-pipelined table-path pack kernel -- parity, A/B against the generic kernel's table path
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "not vh_kernel and not lr_check_border and not speckle and not tall_band" > gpurun_out/r2_t23.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t23.log; tail -3 gpurun_out/r2_t23.log
timeout 600 python tools/fuzz_chain.py 80 2323 > gpurun_out/r2_fuzz23_chain.log 2>&1; tail -1 gpurun_out/r2_fuzz23_chain.log
for l in 1 0; do
  B200S_PACK_LEAN=$l timeout 300 python bench.py --config C4 --steps 10 --warmup 4 --no-cpu --table '' > gpurun_out/r2_b23_lean$l.json 2> gpurun_out/r2_b23_lean$l.err
done
B200S_BENCH_RECT_FLY=1 timeout 300 python bench.py --config C4 --steps 4 --warmup 3 --no-cpu --table '' > gpurun_out/r2_b23_fly.json 2> gpurun_out/r2_b23_fly.err
timeout 300 python bench.py --config C1 --steps 6 --warmup 3 --no-cpu --table '' > gpurun_out/r2_b23_c1.json 2> gpurun_out/r2_b23_c1.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_b23_lean1.json","gpurun_out/r2_b23_lean0.json","gpurun_out/r2_b23_fly.json","gpurun_out/r2_b23_c1.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["frac_of_copy_ceiling"],3), "parity", d["parity_checked"]["frames"], d["parity_checked"]["mismatches"])
        for k,v in d["configs"].items(): print("   ",k, round(v["frames_per_s"]), {a:round(x,1) for a,x in v.get("stage_us",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
