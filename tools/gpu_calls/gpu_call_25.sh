#!/bin/bash
# round 2, call 25: two GPUs with the final kernels -- the driver's launch line (default bench) and the reference arm under torchrun
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
SECONDS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --steps 10 --warmup 4 \
    > gpurun_out/r2_b25_n2.json 2> gpurun_out/r2_b25_n2.err; echo "bench n=2 rc=$? in ${SECONDS}s"
python - <<'PY'
import json
for f in ("gpurun_out/r2_b25_n2.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "n", d["n_gpus"], "fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "parity", d["parity_checked"]["frames"], d["parity_checked"]["mismatches"], "ceil", round(d["e2e"]["copy_ceiling_frames_per_s"]), round(d["e2e"]["copy_ceiling_equal_shares_frames_per_s"]))
        for k,v in d["configs"].items(): print("   ",k, round(v["frames_per_s"]), round(v["e2e_frames_per_s"]), round(v["e2e"]["frac_of_copy_ceiling"],3))
    except Exception as e: print(f, "ERR", e)
PY
