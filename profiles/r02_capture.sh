#!/bin/bash
# ncu evidence of round 2 (run on the GPU box through gpurun; one GPU).  Every profiled command line first runs
# plain (&&) as B200_PROFILING.md asks.  Outputs go to gpurun_out/; the summaries are copied to profiles/ by
# profiles/summarize_ncu.py r02.
set -x
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-alt-operator --no-e2e"
$B > gpurun_out/r02_plain_cfg4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_cfg4.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
for K in gram_sym_topk_kernel refine_tiled_kernel synth_kernel sigma_pairs_f32_kernel knn_merge_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o gpurun_out/r02_prof_$K $B > gpurun_out/r02_ncu_$K.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:sparse_block_matvec -s 130 -c 2 -o gpurun_out/r02_prof_sparse_matvec $B > gpurun_out/r02_ncu_spmv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gram_topk_kernel -s 1 -c 1 -o gpurun_out/r02_prof_gram_topk_limits $B > gpurun_out/r02_ncu_limits.log 2>&1
# north-star kernel 1 (direct analytical synthesis): FP32 pipe utilisation
D="python tests/gpu_synth_direct_run.py"
$D > gpurun_out/r02_plain_direct.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:synth_direct -s 1 -c 1 -o gpurun_out/r02_prof_synth_direct $D > gpurun_out/r02_ncu_direct.log 2>&1
# dense mat-vec (the HBM-bound kernel the north star names) at cfg 2
C2="python bench.py --workload cfg2 --operator dense --steps 1 --warmup 1 --no-cpu-baseline --no-alt-operator --no-e2e"
$C2 > gpurun_out/r02_plain_cfg2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:block_matvec_kernel -s 150 -c 2 -o gpurun_out/r02_prof_dense_matvec $C2 > gpurun_out/r02_ncu_dmv.log 2>&1
ls -la gpurun_out/r02_*
