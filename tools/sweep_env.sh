#!/bin/bash
# usage: tools/sweep_env.sh <workload> <steps> "VAR=V VAR2=V2" ...   -- one bench run per env set, kernel table only
w=$1; st=$2; shift 2
for e in "$@"; do
  echo "=== $w [$e]"
  env $e timeout 200 python bench.py --workload $w --steps $st --warmup 3 --no-cpu-baseline > /tmp/sw.json 2> /tmp/sw.err || tail -3 /tmp/sw.err
  timeout 10 python tools/bench_summary.py < /tmp/sw.json | grep -v "share 0.00"
done
