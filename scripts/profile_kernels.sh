#!/bin/bash
# ncu full captures of named kernels in a short bench run:  profile_kernels.sh "<regex>:<skip>:<count>:<outname>" ...
set -u
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=: read -r pat skip cnt out <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c $cnt -f -o gpurun_out/$out $CMD > gpurun_out/ncu_$out.log 2>&1
  echo "$out rc=$?"
done
