#!/bin/bash
# round-2 final single-GPU measurements of the committed binary
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/f1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f1_pytest.log; tail -n 3 gpurun_out/f1_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 900 python bench.py --impl reference --steps 30 --warmup 3 > gpurun_out/bench_r2_reference.json 2> gpurun_out/f1_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_r2.json 2> gpurun_out/f1_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/f1_bench.err
python scripts/profile_step.py > gpurun_out/f1_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_final_launches.csv python scripts/profile_step.py > gpurun_out/f1_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"reach_build|hyperplane|constraint_eval" -s 6 -c 3 -o gpurun_out/r2_final python scripts/profile_step.py > gpurun_out/f1_ncu2.log 2>&1
python scripts/eval_latency.py 20 > gpurun_out/r2_eval_latency_20.json 2>&1
python scripts/eval_latency.py 10 > gpurun_out/r2_eval_latency_10.json 2>&1
scripts/microbench/pcie_write > gpurun_out/r2_pcie_write.txt 2>&1
python scripts/run_episode.py > gpurun_out/r2_episode.json 2>&1
python scripts/run_sweep.py --problems 1024 --batch 256 > gpurun_out/r2_sweep_standin_1gpu.json 2>&1
python scripts/run_sweep.py --problems 4096 --batch 256 --no-solve > gpurun_out/r2_sweep_nosolve_1gpu.json 2>&1
python scripts/tune_sweep.py 256 10 > gpurun_out/r2_tune_sweep.log 2>&1
tail -n 2 gpurun_out/r2_episode.json gpurun_out/r2_sweep_standin_1gpu.json gpurun_out/r2_sweep_nosolve_1gpu.json; cat gpurun_out/r2_tune_sweep.log
# sweep shape: plain run, then one full capture of the batched reach kernel (16 problems x 128 intervals)
python scripts/tune_sweep.py one 16 10 > gpurun_out/f1_sweep_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:reach_build -s 1 -c 1 -o gpurun_out/r2_final_sweep python scripts/tune_sweep.py one 16 10 > gpurun_out/f1_ncu3.log 2>&1
python scripts/phase_timing.py sweep > /dev/null 2>&1 || true
