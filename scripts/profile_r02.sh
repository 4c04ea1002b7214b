#!/bin/bash
# Round-2 ncu evidence for profiles/: launch list of a short bench run + full captures of every kernel family, of the
# bf16-mode GEMMs and of the WikiMEL front end.  Run under gpurun on ONE GPU (each command first runs WITHOUT ncu and
# must exit 0).  A number printed by a run under ncu is never a bench value.
set -u
BASE="--steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-legs"
CMD="python bench.py $BASE"
mkdir -p gpurun_out
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
cap() {   # cap "<cmd args>" <kernel regex> <skip> <count> <outname>
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/$5 python bench.py $1 > gpurun_out/ncu_$5.log 2>&1
  echo "$5 rc=$?"
}
# second train step: all 21 GEMMs in launch order (projections, fu, g, W_h x2; dZ / dW per layer; dfu .. projection dW)
cap "$BASE" gemm_tcgen05 21 21 prof_gemm
cap "$BASE" frontend_kernel 1 1 prof_frontend
cap "$BASE" gcn_layer_fwd_warp 2 2 prof_layer_fwd
cap "$BASE" gcn_layer_bwd_warp 1 1 prof_layer_bwd
cap "$BASE" gcn_layer0_bwd_col 1 1 prof_layer0_bwd_col
cap "$BASE" score_bwd_warp 1 1 prof_score_bwd
cap "$BASE" score_warp 1 1 prof_score_fwd
# bf16 mode: the single-pass GEMMs of one train step (configs[3])
python bench.py $BASE --precision bf16 > gpurun_out/plain_bf16.log 2>&1 && cap "$BASE --precision bf16" gemm_tcgen05 21 21 prof_gemm_bf16
# WikiMEL shape: front end (the 12.4 MB / mention stream) and the ranking GEMMs (configs[2])
python bench.py $BASE --dataset wikimel > gpurun_out/plain_wm.log 2>&1 && cap "$BASE --dataset wikimel" frontend_kernel 1 1 prof_frontend_wm
# summarise on the box (only gpurun_out/ travels back, <= 64 MiB): tables -> gpurun_out/profiles_r02/, then drop the big reports
DRIN_PROFILE_OUT=gpurun_out/profiles_r02 python scripts/summarize_profiles.py r02 "bench.py $BASE (1 warm-up + 2 timed + 3 event-profiled train steps, ranking passes), 4096 WikiDiverse mentions, fp32-parity"
for r in prof_gemm prof_gemm_bf16 prof_frontend prof_score_bwd; do
  ncu -i gpurun_out/$r.ncu-rep --page details --csv > gpurun_out/profiles_r02/${r}_details.csv 2>/dev/null
done
rm -f gpurun_out/prof_gemm.ncu-rep gpurun_out/prof_gemm_bf16.ncu-rep gpurun_out/prof_layer_fwd.ncu-rep gpurun_out/prof_layer0_bwd_col.ncu-rep gpurun_out/prof_frontend_wm.ncu-rep
du -sh gpurun_out; ls gpurun_out/profiles_r02
