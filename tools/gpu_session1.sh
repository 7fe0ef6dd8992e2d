#!/bin/bash
# First GPU session: microbenchmarks, parity tests, smoke, bench, sweeps, ncu.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== microbench"; timeout 120 ./tools/microbench > gpurun_out/microbench.jsonl 2>&1; tail -30 gpurun_out/microbench.jsonl
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
echo "== sweep n"; timeout 600 python tools/sweep.py --logn 10 12 14 16 18 20 22 --out gpurun_out/sweep_n.jsonl 2>&1 | tail -12
echo "== sweep opts"; timeout 600 python tools/sweep.py --logn 20 --windows 13 14 15 16 17 --sort 0 1 --out gpurun_out/sweep_opts.jsonl 2>&1 | tail -12
timeout 300 python tools/sweep.py --logn 16 --windows 10 11 12 13 14 --out gpurun_out/sweep_opts.jsonl 2>&1 | tail -6
timeout 300 python tools/sweep.py --logn 20 --radix 2 3 4 5 --out gpurun_out/sweep_opts.jsonl 2>&1 | tail -6
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-check --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/launches.csv
echo "== ncu full accumulate"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:KAccumulate -s 2 -c 2 -o gpurun_out/prof_acc -f \
    python bench.py --steps 2 --warmup 3 --no-check --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
