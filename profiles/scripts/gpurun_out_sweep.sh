for b in 0 1; do for w in 0 3 4; do
SUBLINEAR_B200_WARP_BYPASS=$b SUBLINEAR_B200_WARP_CTAS=$w python tests/kernel_timing.py random 2>&1 | tail -1
done; done
SUBLINEAR_B200_WARP_BYPASS=1 python tests/kernel_timing.py banded 2>&1 | tail -1
SUBLINEAR_B200_WARP_BYPASS=1 python tests/kernel_timing.py random 1000000 2>&1 | tail -1
SUBLINEAR_B200_WARP_BYPASS=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
