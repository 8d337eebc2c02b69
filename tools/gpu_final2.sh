#!/bin/bash
# Scratch: closing run after a shade-only change — GPU test suite, bench line, then (if time is left) the shade captures.
# usage: bash tools/gpu_final2.sh GIT_HASH TAG
H=$1; T=${2:-r2g}
rm -f gpurun_out/parity_report.jsonl
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.log
tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
cut -c1-200 gpurun_out/${T}_bench.json
timeout 300 python tools/ncu_profile.py --git-hash $H --workloads breaktime cornell veach --kernels shade --tag $T > gpurun_out/${T}_ncu_profile.log 2>&1; echo "ncu rc=$?"
tail -1 gpurun_out/${T}_ncu_profile.log | cut -c1-300
