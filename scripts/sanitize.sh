#!/bin/bash
# compute-sanitizer over the hand-written kernels (run on the GPU box through gpurun).  Logs -> gpurun_out/sanitizer_*.log;
# scripts/summarize_sanitizer.py condenses them into profiles/.
#   usage: scripts/sanitize.sh [per-tool timeout in s, default 420]
T=${1:-420}
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
python scripts/sanitize_cases.py all > gpurun_out/sanitizer_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitizer_plain.log; exit 1; }
timeout $T $CS --tool memcheck --error-exitcode 9 --print-limit 20 python scripts/sanitize_cases.py all > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/sanitizer_memcheck.log
timeout $T $CS --tool initcheck --error-exitcode 9 --print-limit 20 python scripts/sanitize_cases.py small > gpurun_out/sanitizer_initcheck.log 2>&1
echo "initcheck rc=$?" | tee -a gpurun_out/sanitizer_initcheck.log
timeout $T $CS --tool racecheck --error-exitcode 9 --print-limit 20 python scripts/sanitize_cases.py small > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/sanitizer_racecheck.log
timeout $T $CS --tool synccheck --error-exitcode 9 --print-limit 20 python scripts/sanitize_cases.py small > gpurun_out/sanitizer_synccheck.log 2>&1
echo "synccheck rc=$?" | tee -a gpurun_out/sanitizer_synccheck.log
