#!/bin/sh
# A/B kernel variants on the GPU box: tools/ab.sh lib1.so lib2.so ...   ("default" = in-tree build)
for lib in "$@"; do
  if [ "$lib" = default ]; then unset PICHA_B200_LIB; else export PICHA_B200_LIB=$lib; fi
  for w in ${WORKLOADS:-cfg3 cfg5 cfg4}; do
    timeout 120 python bench.py --workload $w --also none --no-cpu --no-e2e --steps 5 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$lib', '$w', 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'])
    elif 'rror' in l: print(l.strip()[:200])
"
  done
done
