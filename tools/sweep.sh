#!/bin/bash
# Round record sweep: every shipped workload through bench.py (runs of the real schedule) into gpurun_out/$1/.
# Usage: tools/sweep.sh <tag>      (run on the GPU box; stdin must not be a terminal)
exec < /dev/null
tag=${1:-sweep}
out=gpurun_out/$tag
mkdir -p $out
run() {  # name, extra args...
  name=$1; shift
  timeout 400 python bench.py "$@" --no-cpu-baseline > $out/bench_$name.json 2> $out/bench_$name.err
  echo "$name rc=$? $(timeout 20 python tools/bsum.py $out/bench_$name.json 2>/dev/null | head -1 | cut -c1-64)"
}
run qm9_cc            --workload qm9_cc --steps 30 --warmup 3 --profile-steps 2
run qm9_cc_s4         --workload qm9_cc --sampler S4 --steps 30 --warmup 3 --profile-steps 2
run enzymes_small_cc  --workload enzymes_small_cc --steps 30 --warmup 3 --profile-steps 2
run ego_small_cc      --workload ego_small_cc --steps 30 --warmup 3 --profile-steps 2
run grid_small_cc     --workload grid_small_cc --steps 6 --warmup 3 --profile-steps 2
run community_small   --workload community_small --steps 1000 --warmup 5 --profile-steps 2
run ego_small         --workload ego_small --steps 1000 --warmup 5 --profile-steps 2
run qm9               --workload qm9 --steps 1000 --warmup 5 --profile-steps 2
run enzymes_small     --workload enzymes_small --steps 200 --warmup 5 --profile-steps 2
run zinc250k          --workload zinc250k --steps 10 --warmup 3 --profile-steps 2
run enzymes           --workload enzymes --steps 30 --warmup 3 --profile-steps 2
run grid              --workload grid --steps 12 --warmup 3 --profile-steps 2
run qm9_base_cc       --workload qm9_base_cc --steps 6 --warmup 3 --profile-steps 2
run community_small_base_cc --workload community_small_base_cc --steps 6 --warmup 3 --profile-steps 2
