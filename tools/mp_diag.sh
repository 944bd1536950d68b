#!/bin/bash
# usage: diag.sh <label> <max_iter> <tol> [env assignments...]
label=$1; mi=$2; tol=$3; shift 3
d=$(mktemp -d)
for r in 0 1; do
  env "$@" timeout 45 python tests/mp_worker.py $r 2 $d c2_hex_p2_small $mi $tol 1 > $d/out$r.txt 2>&1 &
done
wait
echo "== $label"; for r in 0 1; do tail -4 $d/out$r.txt | sed "s/^/   [$r] /"; done
