#!/bin/bash
# Quick GPU iteration: microbench, parity tests, bench, n-sweep.
set -u
mkdir -p gpurun_out
echo "== microbench"; timeout 120 ./tools/microbench > gpurun_out/microbench.jsonl 2>&1; grep -E "fe_|ge_|imad" gpurun_out/microbench.jsonl | cut -c1-200
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
echo "== sweep n"; timeout 600 python tools/sweep.py --logn 10 12 14 16 18 20 22 --out gpurun_out/sweep_n.jsonl 2>&1 | tail -12
echo "== sweep opts"; timeout 600 python tools/sweep.py --logn 20 --windows 15 16 17 --out gpurun_out/sweep_opts.jsonl 2>&1 | tail -12
