#!/bin/bash
# dev loop: build the product library and a -DDMG_FAST_TIMING twin here, run parity + a short bench on the GPU box
set -e
cd /root/repo
python dismember_b200/build.py > /dev/null
DMG_NVCC_EXTRA="-DDMG_FAST_TIMING" python -c "
from dismember_b200 import build as b
b.build(force=True, out='/root/repo/build/libdismember_gpu_timing.so')"
/usr/local/graft/bin/gpurun --timeout 600 -- "python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4; python bench.py --steps ${STEPS:-32} --no-cpu-baseline $BENCH_ARGS > gpurun_out/dev_bench.json 2> gpurun_out/dev_bench.err; python -c \"import json;d=json.load(open('gpurun_out/dev_bench.json'));print(d['value'], d['roofline']['kernel_ms_avg'], d['config']['fast_stats'])\"; DMG_LIB=build/libdismember_gpu_timing.so python bench.py --steps 16 --no-cpu-baseline $BENCH_ARGS 2>&1 >/dev/null | grep 'fast timing'" 2>&1 | grep -v "^\[gpurun\] sending"
