#!/bin/bash
# sweep an environment variable on the C2 device leg:  tools/ab_env.sh VAR "v1 v2 ..." [extra bench args]
VAR=$1; VALS=$2; shift 2
for v in $VALS; do
  for fl in "" "--no-flush"; do
  env $VAR=$v python bench.py --steps 400 --warmup 50 --no-configs --no-e2e --no-cpu-baseline $fl "$@" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$VAR=$v $fl ms %.4f' % d['ms_per_step'])"
  done
done
