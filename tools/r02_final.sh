# final runs of the round (one GPU): tests, smoke, the bench lines kept under profiles/
set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench_line.err; tail -c 400 gpurun_out/r02_bench_line.err
timeout 300 python bench.py --frames 1 --steps 30 --no-other-configs --no-cpu-baseline > gpurun_out/r02_bench_batch1.json 2> gpurun_out/r02_bench_batch1.err
timeout 600 python bench.py --mode train --frames 8 --steps 4 --warmup 2 > gpurun_out/r02_bench_train.json 2> gpurun_out/r02_bench_train.err; tail -c 300 gpurun_out/r02_bench_train.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; tail -c 300 gpurun_out/r02_bench_reference_arm.err
for f in gpurun_out/r02_bench_*.json; do echo $f; tail -1 $f | cut -c1-300; done
