#!/bin/bash
# env-var values x batch sizes (device leg, no flush):  tools/ab_sizes.sh VAR "values" "sizes"
for v in $2; do for n in $3; do
  env $1=$v python bench.py --envs $n --steps 300 --warmup 40 --no-configs --no-e2e --no-cpu-baseline --no-flush | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1=$v envs=$n ms %.4f  ns/env %.2f' % (d['ms_per_step'], d['ms_per_step']*1e6/$n))"
done; done
