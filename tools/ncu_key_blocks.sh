#!/bin/bash
# Profiling recipe of the key-block rollout kernel (N = G = 500, 74 instances = 296 single-tile CTAs = one wave at two
# CTAs per SM), run under gpurun on one GPU: the plain command first, then its launch list, then ONE `--set full`
# capture.  Outputs go to gpurun_out/; profiles/r02_decode_tc2_kb500_{raw,source}_summary.txt and
# profiles/r02_launches_kb500_b74.csv are made from them (ncu -i ... --page raw|source --csv, tools/ncu_src_summary.py).
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
CMD="python tools/ab_rollout.py --n 500 --g 500 --instances 74 --mode fast --reps 1"
timeout 120 $CMD > gpurun_out/r02c_kb500_plain.log 2>&1 &&
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_kb500_launches.csv $CMD > gpurun_out/r02c_kb500_ncu_list.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:decode_tc2_kernel -s 1 -c 1 -o gpurun_out/r02c_kb500_full -f $CMD > gpurun_out/r02c_kb500_ncu_full.log 2>&1
tail -2 gpurun_out/r02c_kb500_ncu_full.log
