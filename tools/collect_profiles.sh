#!/bin/bash
# gpurun_out/<tag>_* (tools/gpu_evidence.sh) -> the tracked summaries under profiles/ (named per round)
TAG=${1:-r02j}; R=${2:-r02}
cp gpurun_out/${TAG}_bench_1gpu.json profiles/bench_${R}_1gpu.json
cp gpurun_out/${TAG}_bench_reference_arm.json profiles/bench_${R}_reference_arm.json
cp gpurun_out/${TAG}_launches.csv profiles/launches_${R}.csv
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv > profiles/launches_${R}_summary.txt
python tools/ncu_summary.py gpurun_out/${TAG}_pipeline.ncu-rep profiles/ncu_pipeline_${R}_shipped.txt
python tools/ncu_pipeline_json.py gpurun_out/${TAG}_pipeline.ncu-rep 100000 4096 profiles/ncu_pipeline_${R}_shipped.json
python tools/ncu_summary.py gpurun_out/${TAG}_dense_apply.ncu-rep profiles/ncu_dense_apply_${R}.txt
for f in gpurun_out/${TAG}_hbv*.ncu-rep; do b=$(basename $f .ncu-rep); python tools/ncu_summary.py $f profiles/ncu_${b#${TAG}_}_${R}.txt; done
cp gpurun_out/${TAG}_bench_configs_c3.jsonl profiles/bench_configs_${R}_c3.jsonl
cp gpurun_out/${TAG}_bench_configs_c5.jsonl profiles/bench_configs_${R}_c5.jsonl
cp gpurun_out/${TAG}_pytest.log profiles/pytest_gpu_${R}.log
for s in pt_gs_k pt_hs_k hbv_stack pt_ss_k pt_hps_k; do [ -f gpurun_out/e2e_census_$s.json ] && cp gpurun_out/e2e_census_$s.json profiles/e2e_census_${R}_$s.json; done
ls profiles | grep ${R}
