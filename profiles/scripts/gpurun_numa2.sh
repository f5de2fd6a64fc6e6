#!/bin/bash
mkdir -p gpurun_out
true
ncu --metrics lts__t_sectors.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_sectors_srcunit_tex.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum,gpu__time_duration.sum,l1tex__m_xbar2l1tex_read_sectors.sum --clock-control none -k regex:gather_by_die --csv --log-file gpurun_out/numa2_ncu.csv ./sublinear-time-solver_b200/bench/numa_probe gpurun_out/numa_probe80b.bin 80 > gpurun_out/numa2_under_ncu.log 2>&1
rm -f gpurun_out/numa_probe80b.bin gpurun_out/numa_probe80.bin
tail -n 4 gpurun_out/numa2_under_ncu.log
