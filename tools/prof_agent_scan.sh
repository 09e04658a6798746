for lib in libnavgym_b200_old.so libnavgym_b200_ilp1.so libnavgym_b200.so; do
NAVGYM_LIB=/root/repo/nav_gym_b200/$lib ncu --set full --clock-control none --import-source on -k regex:agent_scan -s 30 -c 1 -f -o gpurun_out/f3_agent_${lib%.so} python tools/bench_configs.py crowd > gpurun_out/f3_ncu_$lib.log 2>&1
done
ls -la gpurun_out/f3*
