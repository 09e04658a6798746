#!/bin/bash
# ncu --set full of the pedestrian-lidar kernel on the crowd step, one capture per library build:
#   tools/prof_agent_scan.sh libnavgym_b200.so [other builds in nav_gym_b200/ ...]   -> gpurun_out/agent_<lib>.ncu-rep
for lib in ${@:-libnavgym_b200.so}; do
  NAVGYM_LIB=/root/repo/nav_gym_b200/$lib ncu --set full --clock-control none --import-source on -k regex:agent_scan -s 30 -c 1 -f \
    -o gpurun_out/agent_${lib%.so} python tools/bench_configs.py crowd > gpurun_out/agent_ncu_$lib.log 2>&1
done
ls -la gpurun_out/agent_*
