#!/bin/bash
# A/B of library builds / launch shapes on the C2 device leg:  tools/ab_libs.sh "lib1 lib2" "wpe values"
LIBS=${1:-"libnavgym_b200.so libnavgym_b200_nocoop.so"}
WPES=${2:-"2"}
for lib in $LIBS; do for wpe in $WPES; do
  for fl in "" "--no-flush"; do
  NAVGYM_WPE=$wpe NAVGYM_LIB=/root/repo/nav_gym_b200/$lib python bench.py --steps 400 --warmup 50 --no-configs --no-e2e --no-cpu-baseline $fl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib wpe=$wpe $fl ms %.4f' % d['ms_per_step'])"
  done
done; done
