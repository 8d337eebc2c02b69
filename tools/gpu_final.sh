#!/bin/bash
# Scratch: the round's closing run on the GPU box — GPU test suite, ncu captures of the committed kernels (the figures
# bench.py's roofline blocks read), the bench line with the driver's arguments, the launch list and smoke().
# usage: bash tools/gpu_final.sh GIT_HASH TAG
H=$1; T=${2:-r2f}
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.log
tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 900 python tools/ncu_profile.py --git-hash $H --workloads breaktime cornell veach --tag $T > gpurun_out/${T}_ncu_profile.log 2>&1 \
    && cp gpurun_out/profiles/kernel_profiles.json profiles/kernel_profiles.json
tail -2 gpurun_out/${T}_ncu_profile.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/${T}_bench.json
RPT_GRAPHS=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_breaktime_spp8.csv \
    python bench.py --quick --steps 2 --warmup 1 --spp 8 > /dev/null 2>&1
python tools/launch_shares.py gpurun_out/${T}_launches_breaktime_spp8.csv > gpurun_out/${T}_launch_shares.txt; head -4 gpurun_out/${T}_launch_shares.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log
