#!/bin/bash
# Multi-GPU evidence on one box with N GPUs (N = $1): the hardware parity test for this world size and the bench line of BASELINE configs[3]
N=${1:-2}; TAG=${2:-r02}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus_${N}.txt
if [ $N -le 4 ]; then
  timeout 900 python -m pytest "tests/test_gpu_multi.py::test_all_reduced_catchment_series_equal_the_single_gpu_run[$N]" -m gpu -q > gpurun_out/${TAG}_multi_pytest_${N}gpu.log 2>&1
  tail -3 gpurun_out/${TAG}_multi_pytest_${N}gpu.log
fi
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 2 --warmup 3 \
   > gpurun_out/${TAG}_bench_${N}gpu.out 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "bench $N rc $?"
grep '^{' gpurun_out/${TAG}_bench_${N}gpu.out > gpurun_out/${TAG}_bench_${N}gpu.json; cut -c1-400 gpurun_out/${TAG}_bench_${N}gpu.json
if [ $N -eq 8 ]; then
  SB2_C5_SETS=4096 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 tools/bench_configs.py 5 \
     2> gpurun_out/${TAG}_c5_8gpu.err | grep '^{' > gpurun_out/${TAG}_bench_configs_c5_8gpu.jsonl; cat gpurun_out/${TAG}_bench_configs_c5_8gpu.jsonl
fi
