#!/bin/bash
mkdir -p gpurun_out
cd sublinear-time-solver_b200
make clean >/dev/null
make -j16 NVFLAGS="-gencode arch=compute_100a,code=sm_100a -O1 -g -lineinfo -std=c++17 -fmad=false -ccbin /usr/bin/g++ -Xcompiler -fPIC,-fopenmp,-fsanitize=address,-fno-omit-frame-pointer" > ../gpurun_out/asan_build.log 2>&1
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -Xcompiler -fopenmp,-fsanitize=address -o libsublinear_b200.so build/*.o -ldl >> ../gpurun_out/asan_build.log 2>&1
cd ..
export LD_PRELOAD=$(/usr/bin/gcc -print-file-name=libasan.so)
export ASAN_OPTIONS=protect_shadow_gap=0:detect_leaks=0:halt_on_error=1
timeout 600 python -X faulthandler -m pytest tests/test_state_streaming.py -m gpu -q -x -s -p no:cacheprovider -k streaming > gpurun_out/asan_state.log 2>&1
echo rc=$?
grep -v "^Extension modules" gpurun_out/asan_state.log | head -80
