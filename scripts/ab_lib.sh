#!/bin/bash
# A/B of two builds of the library on the same box: build_ab/libdrin_prev.so (DRIN_B200_LIB override) vs the in-tree one.
# Prints per-stage device times of the headline train step and the bf16 / WikiMEL legs, alternating the two builds.
for rep in 1 2; do
  for lib in build_ab/libdrin_prev.so ""; do
    DRIN_B200_LIB=$lib python bench.py --no-e2e --no-cpu-baseline --steps 10 --warmup 3 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
s=d['stage_ms_per_step']; l=d['legs']
print('${lib:-NEW}'.ljust(28), 'step %.3f' % d['ms_per_step'], ' '.join('%s %.3f' % (k, s[k]) for k in ('gemm','frontend','gcn_fwd','gcn_bwd','score')), '| rank %.3f' % d['ranking']['ms_per_step'], '| bf16 %.3f' % l['bf16_train']['ms_per_step'], '| wm %.3f' % l['wikimel_ranking']['ms_per_pass'])"
  done
done
