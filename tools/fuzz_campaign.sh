#!/bin/sh
# tools/fuzz_campaign.sh [runs per seed] : wild-range parity fuzz over several seeds; prints one line per run
# (cases by kernel, unsupported, out of tolerance) and the totals.  ~20 s per run of 500 cases on one B200.
runs=${1:-4}
total=0; bad=0; faults=0
for seed in 21 22 23 24 25 26 27; do
  for r in $(seq $runs); do
    out=$(python tests/fuzz_parity.py 500 $((seed + 100 * (r - 1))) wild 2>&1)
    line=$(echo "$out" | grep "cases by kernel" | tail -1)
    echo "seed $((seed + 100 * (r - 1))): $line"
    echo "$out" | grep -q "rror" && { faults=$((faults+1)); echo "$out" | grep "rror" | tail -2; }
    echo "$out" | grep "OUT OF TOLERANCE"
    n=$(echo "$line" | python -c "import sys,re; s=sys.stdin.read(); d=re.search(r'\{(.*?)\}', s); print(sum(int(x.split(':')[1]) for x in d.group(1).split(',')) if d else 0)")
    total=$((total + n))
  done
done
echo "TOTAL cases executed: $total, runs with a CUDA error: $faults"
