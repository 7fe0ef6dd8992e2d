#!/bin/bash
# Profiling session: parity tests, bench (with CPU baseline), ncu launch list, ncu --set full of the dominant kernels.
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
echo "== sweep n"; timeout 600 python tools/sweep.py --logn 10 12 14 16 18 20 22 --out gpurun_out/sweep_n.jsonl 2>&1 | tail -8
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-check --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/launches.csv
# .ncu-rep files are large (gpurun_out/ is capped at 64 MiB): export the raw page as CSV on the box, drop the report
export_rep() {  # $1 = report stem
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2> gpurun_out/$1_export.err
    rm -f gpurun_out/$1.ncu-rep
}
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:KAccumulate -s 3 -c 2 -o gpurun_out/prof_acc -f \
    python bench.py --steps 3 --warmup 3 --no-check --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-300
ncu -i gpurun_out/prof_acc.ncu-rep --page source --csv --kernel-name-base demangled 2>/dev/null | head -400 > gpurun_out/prof_acc_source_head.csv
export_rep prof_acc
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -k regex:'KScatter|KDigitsHist|KReduce' -s 9 -c 4 -o gpurun_out/prof_other -f \
    python bench.py --steps 3 --warmup 3 --no-check --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
tail -3 gpurun_out/ncu_full2.log | cut -c1-300
export_rep prof_other
echo "== ncu full: fold / scalar-vector / text kernels of one AC20 proof (N = 2^16)"
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -k regex:'KFold|KScalarAxpy|KScalarDotPartial|KPointText|KScalarText|KTextCompact' -c 12 -o gpurun_out/prof_ac20 -f \
    python tools/bench_ac20.py --log2n 16 --repeat 1 > gpurun_out/ncu_full3.log 2>&1
tail -3 gpurun_out/ncu_full3.log | cut -c1-300
export_rep prof_ac20
echo "== ncu full: generator fold kernels (one thread per element at 2^16 outputs, four lanes per element at 2^12)"
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k regex:'KFold|KNormalize' -c 9 -o gpurun_out/prof_fold -f \
    python tools/bench_fold.py --log2half 12 16 > gpurun_out/ncu_full4.log 2>&1
tail -2 gpurun_out/ncu_full4.log | cut -c1-200
export_rep prof_fold
ls -la gpurun_out
