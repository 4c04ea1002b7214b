#!/bin/bash
# ncu evidence for profiles/: launch list of a short bench run + full captures of the top kernels.
# Run under gpurun on ONE GPU.  A number printed by a run under ncu is never a bench value.
set -u
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
# kernel regex : launches to skip : launches to capture : output name  (the GEMM capture covers the four input projections
# of the first ranking forward, the largest being the entity-image projection)
for spec in "gemm_tcgen05:120:4:prof_gemm" "gcn_layer_bwd_warp:1:1:prof_layer_bwd" "gcn_layer0_bwd_col:1:1:prof_layer0_bwd_col" "frontend_kernel:1:1:prof_frontend" "gcn_layer_fwd_warp:2:2:prof_layer_fwd" "score_bwd_warp:1:1:prof_score_bwd" "score_warp:1:1:prof_score_fwd"; do
  IFS=: read -r pat skip cnt out <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c $cnt -f -o gpurun_out/$out $CMD > gpurun_out/ncu_$out.log 2>&1
  echo "$out rc=$?"
done
ls -la gpurun_out
