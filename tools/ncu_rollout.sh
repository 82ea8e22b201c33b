#!/bin/bash
# Profiling recipe of the rollout kernel (run under gpurun, one GPU): the plain command first, then the launch list
# of the same command, then ONE `--set full` capture of the dominant kernel.  Outputs go to gpurun_out/; the
# summaries that are committed under profiles/ are made from them by tools/ncu_src_summary.py and tools/ncu_constants.py.
#   tools/ncu_rollout.sh <mode: fast|x3> <kernel regex: decode_tc2_kernel|decode_tc_kernel> <tag>
set -e
MODE=${1:-fast}; KREGEX=${2:-decode_tc2_kernel}; TAG=${3:-r02_$MODE}
CMD="python bench.py --mode $MODE --batch 296 --steps 1 --warmup 3 --cpu-sample 1 --no-e2e"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 3 -c 1 -o gpurun_out/${TAG}_full -f $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
git rev-parse HEAD 2>/dev/null > gpurun_out/${TAG}_git_head.txt || true
