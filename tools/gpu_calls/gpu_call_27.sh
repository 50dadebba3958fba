#!/bin/bash
# round 2, call 27: the table-miss test (short table through the test hook) and the chain tests around it on the final library
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke27.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke27.log
timeout 900 python -m pytest tests -m gpu -x -q -k "pack_table or fused or batched or bench_configuration or graph or colour or process_pair" > gpurun_out/r2_t27.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t27.log; tail -3 gpurun_out/r2_t27.log
timeout 200 python bench.py --config C4 --steps 6 --warmup 3 --no-cpu --table '' > gpurun_out/r2_b27.json 2> gpurun_out/r2_b27.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_b27.json").read().strip().splitlines()[-1])
print("fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "parity", d["parity_checked"]["frames"], d["parity_checked"]["mismatches"])
PY
