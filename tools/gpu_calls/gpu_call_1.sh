#!/bin/bash
# round 2, call 1: parity of the timed path, bench A/B of the pack mode, bare D2H probe, host topology
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
(nvidia-smi topo -m; echo; lscpu | head -40; echo; numactl -H 2>&1; cat /sys/devices/system/node/node*/meminfo 2>/dev/null | grep MemTotal; nvidia-smi -q | grep -i -A3 "pci" | head -40) > gpurun_out/r2_host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t1.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_b1.json 2> gpurun_out/r2_b1.err
B200S_PACK_DIRECT=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_b1_direct.json 2> gpurun_out/r2_b1_direct.err
timeout 300 python tools/d2h_probe.py --gpus 1 > gpurun_out/r2_probe1.json 2> gpurun_out/r2_probe1.err
timeout 300 python tools/d2h_probe.py --gpus 1 --with-h2d >> gpurun_out/r2_probe1.json 2>> gpurun_out/r2_probe1.err
timeout 300 python tools/d2h_probe.py --gpus 1 --mode registered >> gpurun_out/r2_probe1.json 2>> gpurun_out/r2_probe1.err
tail -3 gpurun_out/r2_t1.log; cat gpurun_out/r2_probe1.json
