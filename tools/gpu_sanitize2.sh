#!/bin/bash
# memcheck over the whole GPU suite; racecheck (shared-memory hazards) over the small parity cases
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 2400 $S --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_gram_tc.py --deselect tests/test_cpp_updaters.py --deselect tests/test_cpp_host_mirror.py --deselect tests/test_cpp_gnss.py > gpurun_out/san_full.log 2>&1
echo "rc=$?" >> gpurun_out/san_full.log; grep -E "passed|failed|ERROR SUMMARY|rc=" gpurun_out/san_full.log | tail -4
timeout 1800 $S --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_landmarks.py -m gpu -q -x -k "layout or propagate or marginalize or ekf_update or all_obs_frames or gnss_update or delayed or landmark" > gpurun_out/san_race.log 2>&1
echo "rc=$?" >> gpurun_out/san_race.log; grep -E "passed|failed|RACECHECK SUMMARY|rc=|hazard" gpurun_out/san_race.log | sort | uniq -c | sort -rn | head -8
