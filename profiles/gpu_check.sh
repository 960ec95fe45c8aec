#!/bin/bash
# One bounded GPU pass used while developing: smoke of the step at full batch, the parity tests, the launch list and
# a short bench.  Usage (through gpurun): bash profiles/gpu_check.sh TAG ["pytest -k expression"]
TAG=${1:-chk}; KEXPR=${2:-"step_vs_oracle or generic or full_batch or several"}
mkdir -p gpurun_out
for args in "7 65536 0" "6 65536 0"; do timeout 60 python profiles/dbg_step.py $args >> gpurun_out/${TAG}_dbg.log 2>&1; echo "rc=$? ($args)" >> gpurun_out/${TAG}_dbg.log; done
cat gpurun_out/${TAG}_dbg.log
timeout 600 python -m pytest tests -m gpu -q -x -k "$KEXPR" > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:atacom_ -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --replays 5 > /dev/null 2> gpurun_out/${TAG}_ncu.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
