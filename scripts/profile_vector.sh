#!/bin/bash
# ncu full captures of the vector-edge kernels (gcn_edge_feature="vector"); run under gpurun on ONE GPU.
set -u
CMD="python bench.py --edge-feature vector --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_vector.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_vector.log; exit 1; }
for spec in "vec_layer_bwd:0:2:vprof_layer_bwd" "vec_layer_fwd:0:2:vprof_layer_fwd" "vec_rows_bwd:0:2:vprof_rows_bwd"; do
  IFS=: read -r pat skip cnt out <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c $cnt -f -o gpurun_out/$out $CMD > gpurun_out/ncu_$out.log 2>&1
  echo "$out rc=$?"
done
