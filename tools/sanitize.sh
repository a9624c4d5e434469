#!/bin/bash
# compute-sanitizer over the step / interpolation / routing / goal kernels.  Output: gpurun_out/sanitizer_<tool>_<case>.log
# racecheck covers shared-memory hazards (TMA staging buffers, routing tiles, goal reductions); the ticket / progress protocol of the
# time-sliced kernels lives in global memory, where memcheck + initcheck (no read of a state row before the previous slice published it
# would show as an uninitialised or out-of-bounds access) and the bit-reproducibility tests are the evidence.
mkdir -p gpurun_out
run() {  # tool case cells steps
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $1 --error-exitcode 9 --print-limit 20 python tools/sanitizer_case.py $2 $3 $4 \
     > gpurun_out/sanitizer_$1_$2.log 2>&1
  echo "$1 $2 rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$1_$2.log | tail -1) $(grep ' ok ' gpurun_out/sanitizer_$1_$2.log | tail -1)"
}
run memcheck ptgsk 60000 640
run memcheck routing 4000 500
run memcheck goal 2000 480
run racecheck ptgsk 20000 320
run racecheck routing 4000 300
run racecheck goal 2000 240
run synccheck ptgsk 20000 320
run initcheck ptgsk 20000 320
