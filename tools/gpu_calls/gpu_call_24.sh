#!/bin/bash
# round 2, call 24: final state of the tree -- smoke, all GPU tests, the four fuzz sweeps, default bench line, C4r / C4s
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke24.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke24.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t24.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t24.log; tail -3 gpurun_out/r2_t24.log
timeout 600 python tools/fuzz_parity.py 300 2424 > gpurun_out/r2_fuzz24_parity.log 2>&1; tail -1 gpurun_out/r2_fuzz24_parity.log
timeout 400 python tools/fuzz_speckle.py 150 2424 > gpurun_out/r2_fuzz24_speckle.log 2>&1; tail -1 gpurun_out/r2_fuzz24_speckle.log
timeout 400 python tools/fuzz_chain.py 60 2424 > gpurun_out/r2_fuzz24_chain.log 2>&1; tail -1 gpurun_out/r2_fuzz24_chain.log
timeout 400 python tools/fuzz_cuda_compat.py 60 2424 > gpurun_out/r2_fuzz24_cc.log 2>&1; tail -1 gpurun_out/r2_fuzz24_cc.log
SECONDS=0; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_b24.json 2> gpurun_out/r2_b24.err; echo "bench rc=$? in ${SECONDS}s"
timeout 300 python bench.py --config C4r --steps 6 --warmup 3 --no-cpu --table '' > gpurun_out/r2_b24_c4r.json 2> gpurun_out/r2_b24_c4r.err
timeout 300 python bench.py --config C4s --steps 6 --warmup 3 --no-cpu --table '' > gpurun_out/r2_b24_c4s.json 2> gpurun_out/r2_b24_c4s.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_b24.json","gpurun_out/r2_b24_c4r.json","gpurun_out/r2_b24_c4s.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["frac_of_copy_ceiling"],3), "parity", d["parity_checked"]["frames"], d["parity_checked"]["mismatches"], "roof", round(d["roofline"]["frac"],3), "cpu", d.get("cpu_baseline",{}).get("value"))
        for k,v in d["configs"].items(): print("   ",k, round(v["frames_per_s"]), round(v["e2e_frames_per_s"]), round(v["e2e"]["frac_of_copy_ceiling"],3), round(v.get("matcher_us",0),1), round(v.get("matcher_tevals_per_s",0),3), {a:round(x,1) for a,x in v.get("stage_us",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
