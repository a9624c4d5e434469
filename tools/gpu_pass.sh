#!/bin/bash
# round 2, GPU pass (script reused per build: pass the tag as $1): parity tests, the default bench line, the launch list and one ncu --set full capture of the three step kernels
# (superseded by tools/gpu_evidence.sh; kept for the round-1 profile names it produced)
TAG=${1:-r02x}
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc $?"
cat gpurun_out/${TAG}_bench.json | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --years 1 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ptgsk_ --launch-skip 6 --launch-count 3 -f -o gpurun_out/${TAG}_pipeline \
   python bench.py --years 1 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc $?"
ls -la gpurun_out | tail -8
