#!/bin/bash
# Round-2 evidence run (one B200): GPU test suite, the default bench line, the ncu launch list of the
# bench command and full captures of the step kernel and of the crowd step's kernels.
set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/k1_gputests.log 2>&1
(time python bench.py) > gpurun_out/k1_bench.json 2> gpurun_out/k1_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/k1_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/k1_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 9 -c 1 -f -o gpurun_out/k1_step_full python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-e2e > gpurun_out/k1_ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:'agent_scan_kernel|policy_features_umma2|dense_umma|peds_plan|peds_move|peds_advance' -s 60 -c 12 -f -o gpurun_out/k1_crowd_full python tools/bench_configs.py crowd > gpurun_out/k1_ncu_crowd.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 60 --csv --log-file gpurun_out/k1_crowd_launches.csv python tools/bench_configs.py crowd > /dev/null 2>&1
tail -3 gpurun_out/k1_gputests.log; tail -c 300 gpurun_out/k1_bench.err; ls -la gpurun_out/k1_*
