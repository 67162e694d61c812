# ncu --set full captures of the conv kernel at the three profiled layer shapes (each command first runs plain)
set -x
TAG=${1:-r02p}
cap() { name=$1; rx=$2; skip=$3; shift 3
  timeout 120 "$@" > gpurun_out/plain_$name.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/${TAG}_$name "$@" > gpurun_out/ncu_$name.log 2>&1; }
cap spconv_l1_48_48 spconv_tc_kernel 2 python tools/run_spconv.py 1 48 48 3
cap spconv_l2_96_96 spconv_tc_kernel 2 python tools/run_spconv.py 2 96 96 3
cap spconv_l4_384_384 spconv_tc_kernel 2 python tools/run_spconv.py 4 384 384 3
ls -la gpurun_out/${TAG}_*
