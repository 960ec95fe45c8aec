#!/bin/bash
# profiles/ab_fix.sh TAG VARIANT...: on one GPU box, (1) step time of each build_variants/VARIANT.so in the default
# (two-pass) mode, interleaved, twice; (2) the step/fix parity tests with each variant.  Output: gpurun_out/TAG_*.log
TAG=$1; shift
mkdir -p gpurun_out
for rep in 1 2; do
  for v in "$@"; do
    echo -n "$v: " >> gpurun_out/${TAG}_time.log
    ATACOM_B200_LIB=$PWD/build_variants/$v.so timeout 120 python profiles/profile_step.py time 2>&1 | grep "us/launch\|rror" | cut -c1-70 >> gpurun_out/${TAG}_time.log
  done
done
cat gpurun_out/${TAG}_time.log
for v in "$@"; do
  ATACOM_B200_LIB=$PWD/build_variants/$v.so timeout 600 python -m pytest tests -m gpu -q -x -k "fullsize or step_vs_oracle or several_active or full_batch or host_buffer_entry_point" > gpurun_out/${TAG}_pytest_$v.log 2>&1
  echo "$v pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_$v.log
done
