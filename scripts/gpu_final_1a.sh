#!/bin/bash
# round-2 final single-GPU measurements, part a: bench (both arms), launch list + full capture of one step, latency tables
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 30 --warmup 3 > gpurun_out/bench_r2_reference.json 2> gpurun_out/f1_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_r2.json 2> gpurun_out/f1_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/f1_bench.err
python scripts/profile_step.py > gpurun_out/f1_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_final_launches.csv python scripts/profile_step.py > gpurun_out/f1_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"reach_build|hyperplane|constraint_eval" -s 6 -c 3 -o gpurun_out/r2_final python scripts/profile_step.py > gpurun_out/f1_ncu2.log 2>&1
python scripts/eval_latency.py 20 > gpurun_out/r2_eval_latency_20.json 2>&1
python scripts/eval_latency.py 10 > gpurun_out/r2_eval_latency_10.json 2>&1
scripts/microbench/pcie_write > gpurun_out/r2_pcie_write.txt 2>&1
ls -la gpurun_out | head -30; du -sh gpurun_out
