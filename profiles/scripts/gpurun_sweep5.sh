#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s5_pytest.log 2>&1
tail -n 8 gpurun_out/s5_pytest.log
python bench.py > gpurun_out/s5_bench.json 2> gpurun_out/s5_bench.err
cat gpurun_out/s5_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s5_bench_ref.json 2>> gpurun_out/s5_bench.err
cat gpurun_out/s5_bench_ref.json
ncu --set full --clock-control none --import-source on -k regex:sell_kernel --launch-skip 6 --launch-count 1 -f -o gpurun_out/s5_sell_push python tests/kernel_timing.py random > gpurun_out/s5_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s5_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s5_bench_under_ncu.log 2>&1
tail -n 3 gpurun_out/s5_ncu.log
