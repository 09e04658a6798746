set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/f1_gputests.log 2>&1
(time python bench.py) > gpurun_out/f1_bench.json 2> gpurun_out/f1_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f1_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/f1_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 30 -c 1 -f -o gpurun_out/f1_step_full python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-e2e > gpurun_out/f1_ncu_full.log 2>&1
tail -3 gpurun_out/f1_gputests.log; tail -c 600 gpurun_out/f1_bench.err
