#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_mirror.py -m gpu -x -q 2>&1 | tail -n 15
timeout 900 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/c9_ref.json 2> gpurun_out/c9_ref.err; echo "ref rc=$?"; cat gpurun_out/c9_ref.json; tail -n 5 gpurun_out/c9_ref.err
timeout 900 python bench.py --steps 30 --warmup 3 > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err; echo "bench rc=$?"; cat gpurun_out/c9_bench.json; tail -n 15 gpurun_out/c9_bench.err
