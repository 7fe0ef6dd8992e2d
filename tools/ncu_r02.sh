#!/bin/bash
# Round-2 ncu captures (run on the GPU box through gpurun; one GPU).  Writes CSV summaries under gpurun_out/ncu/.
set -u
mkdir -p gpurun_out/ncu
N="ncu --clock-control none --kernel-name-base demangled"
M="gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum,l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum"
# 1. launch list of the headline bench (shares of the step)
[ -n "${SKIP_LAUNCHES:-}" ] || $N --metrics gpu__time_duration.sum -c 1500 --csv --log-file gpurun_out/ncu/launches_bench_2p20.csv python bench.py --steps 1 --warmup 3 --msms-per-step 4 --no-extras --no-cpu-baseline > gpurun_out/ncu/bench_under_ncu.log 2>&1
# 2. Ed25519 accumulate kernels: per-bucket at 2^20 (plain), segmented over tables at 2^16 / 2^17, per-bucket over tables at 2^20
$N --metrics $M -k regex:KAccumulate -c 6 --launch-skip 6 --csv --log-file gpurun_out/ncu/ed_accumulate_plain_2p20.csv python tools/sweep.py --logn 20 --steps 6 > /dev/null 2>&1
$N --metrics $M -k regex:KAccumulate -c 4 --launch-skip 4 --csv --log-file gpurun_out/ncu/ed_accumulate_tables_2p17.csv python tools/sweep.py --logn 17 --precompute 16 --steps 6 > /dev/null 2>&1
$N --metrics $M -k regex:KAccumulate -c 4 --launch-skip 4 --csv --log-file gpurun_out/ncu/ed_accumulate_tables_2p20.csv python tools/sweep.py --logn 20 --precompute 16 --steps 6 > /dev/null 2>&1
$N --metrics $M -k regex:"KAccumulate" -c 4 --launch-skip 4 --csv --log-file gpurun_out/ncu/ed_accumulate_plain_2p22.csv python tools/sweep.py --logn 22 --steps 5 > /dev/null 2>&1
# 3. counting sort + tail kernels at 2^20 (one MSM's worth each)
$N --metrics $M -k regex:"KDigitsHist|KScatter|vmsm_scan|vmsm_order|KReduce|KFinal|KOverflow|KCombine|KSegFix" -c 40 --launch-skip 60 --csv --log-file gpurun_out/ncu/ed_sort_tail_2p20.csv python tools/sweep.py --logn 20 --steps 6 > /dev/null 2>&1
# 4. BN256 kernels at 2^14 (G1 then G2)
for cv in 1 2; do
  # plain path first (segments, c = 11, Horner chain), then the MSMs over key tables (two shared bucket sets): the last
  # 40 matched launches belong to the table MSMs
  $N --metrics $M -k regex:"KAccumulate|KSegFix|KSegLongFix|KReduceWQ|KFinalWQ|KOverflowW|KCombineW" -c 200 --csv --log-file gpurun_out/ncu/bn_g${cv}_2p14.csv python tools/bench_bn256.py --log2n 14 --steps 2 --curves $cv --no-proof > /dev/null 2>&1
done
# 5. transcript text kernels + fold kernels at N = 2^16
$N --metrics $M -k regex:"KPointText|KTextCompact|KScalarText|KFold|KNormalize|vmsm_lens" -c 40 --csv --log-file gpurun_out/ncu/ac20_kernels_2p16.csv python tools/bench_ac20.py --log2n 16 --repeat 1 > /dev/null 2>&1
ls -la gpurun_out/ncu
