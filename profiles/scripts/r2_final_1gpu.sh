#!/bin/bash
# round 2, final single-GPU evidence run: full GPU suite, smoke, bench lines (headline, C2, banded), C3/C4 timings, one warm
# ncu --set full capture of the dominant kernel and the launch list of a short bench run.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2z_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; cut -c1-200 gpurun_out/r2z_bench.json
python bench.py --workload c2_n1M_nnz10M > gpurun_out/r2z_bench_c2.json 2>> gpurun_out/r2z_bench.err
python bench.py --workload banded_n10M_nnz100M --no-cpu-baseline > gpurun_out/r2z_bench_banded.json 2>> gpurun_out/r2z_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_reference.json 2>> gpurun_out/r2z_bench.err
timeout 500 python tests/config_timing.py both 2>&1 | grep "^{" > gpurun_out/r2z_config_c3_c4.jsonl
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:slab_kernel --launch-skip 14 --launch-count 1 -f -o gpurun_out/r2z_slab_push python tests/kernel_timing.py random > gpurun_out/r2z_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2z_bench_under_ncu.log 2>&1
tail -n 2 gpurun_out/r2z_ncu.log
