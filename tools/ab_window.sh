#!/bin/bash
# A/B of the shared-memory EDT window (build flag NAVGYM_SMEM_WINDOW=R: the (2R)^2 cells around the
# robot staged per scan) against the L2 path, C2 device leg.  Builds are made by the caller:
#   nvcc ... -DNAVGYM_SMEM_WINDOW=16                         -> libnavgym_b200_win16.so (16 CTAs/SM as the product)
#   nvcc ... -DNAVGYM_SMEM_WINDOW=32 -DNAVGYM_CTAS_PER_SM=9  -> libnavgym_b200_win32.so (24 KB smem per CTA)
#   nvcc ... -DNAVGYM_CTAS_PER_SM=9                          -> libnavgym_b200_occ9.so  (the product kernel at that occupancy)
OUT=${1:-gpurun_out/ab_window.txt}
M="gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum"
: > $OUT
for lib in libnavgym_b200.so libnavgym_b200_win16.so libnavgym_b200_win32.so libnavgym_b200_occ9.so; do
  export NAVGYM_LIB=$PWD/nav_gym_b200/$lib
  echo "=== $lib" >> $OUT
  timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -1 >> $OUT
  for fl in "" "--no-flush"; do
    python bench.py --steps 400 --warmup 50 --no-configs --no-e2e --no-cpu-baseline $fl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C2 4096 envs $fl ms/step %.4f' % d['ms_per_step'])" >> $OUT
  done
  python bench.py --envs 32768 --steps 200 --warmup 30 --no-configs --no-e2e --no-cpu-baseline --no-flush | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('32768 envs no-flush ms/step %.4f' % d['ms_per_step'])" >> $OUT
  timeout 300 ncu --metrics $M --clock-control none -k regex:step_kernel --launch-skip 60 -c 3 --csv --log-file /tmp/ab_$lib.csv python bench.py --steps 20 --warmup 50 --no-configs --no-e2e --no-cpu-baseline > /dev/null 2>&1
  python - /tmp/ab_$lib.csv >> $OUT <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
acc = collections.defaultdict(list)
for r in rows[1:]:
    acc[r[hdr.index('Metric Name')]].append(float(r[hdr.index('Metric Value')].replace(',', '')))
for k, v in acc.items():
    print('  ncu %-62s %14.1f' % (k, sum(v) / len(v)))
PY
done
cat $OUT
