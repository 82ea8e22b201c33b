#!/bin/bash
# A/B the in-tree library against build_lab/libhtsp_base.so on the GPU box: tools/ab.sh <out.jsonl> [extra ab_rollout args]
out=$1; shift
: > $out
for cfg in "--mode x3" "--mode fast" "--mode x3 --n 100 --g 100"; do
  HTSP_LIB_PATH=$PWD/build_lab/libhtsp_base.so python tools/ab_rollout.py --tag base $cfg "$@" >> $out 2>>$out.err
  python tools/ab_rollout.py --tag new $cfg "$@" >> $out 2>>$out.err
done
cat $out; tail -3 $out.err
