#!/bin/bash
# round 2, call 12 (2 GPUs): the driver's sequence -- smoke, GPU tests, reference arm, bench at N=1 and under torchrun at N=2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke12.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke12.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t12.log; tail -3 gpurun_out/r2_t12.log
SECONDS=0; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_b12_n1.json 2> gpurun_out/r2_b12_n1.err; echo "bench n=1 rc=$? in ${SECONDS}s"
SECONDS=0; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_b12_n2.json 2> gpurun_out/r2_b12_n2.err; echo "bench n=2 rc=$? in ${SECONDS}s"
SECONDS=0; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29578 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2_b12_ref2.json 2> gpurun_out/r2_b12_ref2.err; echo "ref n=2 rc=$? in ${SECONDS}s"
python - <<'PY'
import json
for f in ("gpurun_out/r2_b12_n1.json","gpurun_out/r2_b12_n2.json","gpurun_out/r2_b12_ref2.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "n", d["n_gpus"], "fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1))
        if "configs" in d:
            print("   parity", d["parity_checked"]["frames"], d["parity_checked"]["mismatches"], "launches", d["gpu_launches"], "cpu", (d["cpu_baseline"] or {}).get("value"), "roof", d["roofline"] and round(d["roofline"]["frac"],3), "ceil", round(d["e2e"]["copy_ceiling_frames_per_s"]), round(d["e2e"]["copy_ceiling_equal_shares_frames_per_s"]))
            for k,v in d["configs"].items(): print("   ",k, round(v["frames_per_s"]), round(v["e2e_frames_per_s"]), round(v["e2e"]["frac_of_copy_ceiling"],3), round(v.get("matcher_us",0),1), round(v.get("matcher_tevals_per_s",0),3), round(v.get("frac",0),3), {a:round(x,1) for a,x in v.get("stage_us",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
