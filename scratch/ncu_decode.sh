set -e
CMD="python bench.py --mode x3 --batch 296 --steps 1 --warmup 1 --cpu-sample 1"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:decode_mma -s 1 -c 1 -o gpurun_out/decode_mma_cur -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
