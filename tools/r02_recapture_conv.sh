set -x
B="python bench.py --steps 1 --warmup 3 --no-other-configs --no-cpu-baseline"
timeout 200 $B > gpurun_out/plain_bench.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02o_launches.csv $B > gpurun_out/ncu_bench.log 2>&1
cap() { name=$1; rx=$2; skip=$3; shift 3
  timeout 120 "$@" > gpurun_out/plain_$name.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/r02o_$name "$@" > gpurun_out/ncu_$name.log 2>&1; }
cap spconv_l1_48_48 spconv_tc_kernel 2 python tools/run_spconv.py 1 48 48 3
cap spconv_l2_96_96 spconv_tc_kernel 2 python tools/run_spconv.py 2 96 96 3
cap spconv_l4_384_384 spconv_tc_kernel 2 python tools/run_spconv.py 4 384 384 3
ls -la gpurun_out/r02o_*
