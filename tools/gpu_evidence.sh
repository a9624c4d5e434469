#!/bin/bash
# Round-end evidence on ONE B200 (tag = $1): the -m gpu suite, the default bench line and the CPU arm, the ncu launch list of the default
# command's one-year form, ncu --set full of the three pt_gs_k step kernels (a chunk of the shipped configuration), of the HBV step kernels
# and of the interpolation contractions, and the secondary configurations.  Everything lands in gpurun_out/<tag>_*.
# gpurun brings back at most 64 MiB per call: part "a" = tests, bench lines, launch list, secondary configurations; part "b" = the three ncu
# --set full captures (about 50 MiB); no second argument = both (only where nothing limits the size of gpurun_out).
TAG=${1:-r02}; PART=${2:-ab}
mkdir -p gpurun_out
if [[ $PART == *a* ]]; then
nvidia-smi -L > gpurun_out/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err; echo "bench rc $?"
timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference_arm.err; echo "reference arm rc $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --years 1 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc $?"
fi
if [[ $PART == *b* ]]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ptgsk_ --launch-skip 3 --launch-count 3 -f -o gpurun_out/${TAG}_pipeline \
   python bench.py --years 1 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${TAG}_ncu_pipeline.log 2>&1; echo "ncu pipeline rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_apply --launch-skip 10 --launch-count 3 -f -o gpurun_out/${TAG}_dense_apply \
   python bench.py --years 1 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${TAG}_ncu_dense.log 2>&1; echo "ncu dense rc $?"
SB2_C3_YEARS=0.06 SB2_C3_WINDOW=256 timeout 900 ncu --set full --clock-control none --import-source on -k regex:hbv_run_kernel --launch-skip 8 --launch-count 2 -f \
   -o gpurun_out/${TAG}_hbv python tools/bench_configs.py 3 > gpurun_out/${TAG}_ncu_hbv.log 2>&1; echo "ncu hbv rc $?"
fi
if [[ $PART == *a* ]]; then
SB2_C3_YEARS=1 timeout 900 python tools/bench_configs.py 3 > gpurun_out/${TAG}_bench_configs_c3.jsonl 2> gpurun_out/${TAG}_bench_configs_c3.err; echo "c3 rc $?"; cat gpurun_out/${TAG}_bench_configs_c3.jsonl
SB2_C5_SETS=128 timeout 900 python tools/bench_configs.py 5 > gpurun_out/${TAG}_bench_configs_c5.jsonl 2> gpurun_out/${TAG}_bench_configs_c5.err; echo "c5 rc $?"; cat gpurun_out/${TAG}_bench_configs_c5.jsonl
fi
ls -la gpurun_out | grep ${TAG}_; du -sm gpurun_out
