#!/bin/bash
# Runs on the GPU box (under gpurun): for each workload the plain bench command first, then -- only if it exited 0 --
# the ncu launch list of the same command and one `--set full` capture of the dominant kernel.  Outputs -> gpurun_out/.
set -u
TAG=${1:-r01b}
K='regex:l2norm|filter_mma|recheck|filter_fp32|pack_results'
for w in ${WORKLOADS:-cfg3 cfg4 cfg1}; do
  CMD="python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
  $CMD > gpurun_out/plain_$w.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 9 -c 12 --csv \
      --log-file gpurun_out/${TAG}_launches_$w.csv $CMD > gpurun_out/ncu_list_$w.log 2>&1
  if [ "$w" != "cfg1" ] || true; then
    $CMD > gpurun_out/plain2_$w.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:filter_mma -s 3 -c 1 \
        -f -o gpurun_out/${TAG}_k2_$w $CMD > gpurun_out/ncu_full_$w.log 2>&1
  fi
done
ls -la gpurun_out/*.ncu-rep
