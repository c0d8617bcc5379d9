#!/bin/bash
# Collect the round's measured evidence on one B200 into gpurun_out/ (copied to profiles/ afterwards).
# Usage (from the repo root, under gpurun): bash scripts/collect_profiles.sh r01b
tag=${1:-r01b}
out=gpurun_out
mkdir -p $out scripts/_build
for pr in probe_rsq probe_latency probe_block probe_exact_arith; do
  [ -x scripts/_build/$pr ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 $([ $pr = probe_exact_arith ] || echo -fmad=false) -std=c++17 -o scripts/_build/$pr scripts/$pr.cu 2>/dev/null
done
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err
python bench.py --steps 3 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python scripts/run_configs.py --out $out/${tag}_configs.json > $out/${tag}_configs.log 2>&1
python scripts/step_profile.py --out $out/${tag}_step_profile.json > $out/${tag}_step_profile.log 2>&1
python scripts/coop_probe.py > $out/${tag}_coop_vs_graph.txt 2>&1
scripts/_build/probe_rsq > $out/${tag}_probe_rsq.txt 2>&1
scripts/_build/probe_latency > $out/${tag}_probe_latency.txt 2>&1
scripts/_build/probe_block > $out/${tag}_probe_block.txt 2>&1
scripts/_build/probe_exact_arith > $out/${tag}_probe_exact_arith.txt 2>&1
# profiler passes last (numbers printed under ncu are never bench values)
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/${tag}_launches_bench.csv python bench.py --steps 3 --warmup 3 > $out/${tag}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_exact_tiled -s 1 -c 1 -o $out/${tag}_k_exact_tiled -f python scripts/ncu_exact_driver.py > $out/${tag}_ncu_exact.log 2>&1
# the full capture of the 870 ms headline kernel takes ~7 GPU-minutes of replays: only when asked for
[ -n "$FULL_NCU" ] && ncu --set full --clock-control none --import-source on -k regex:k_fast_tiled -s 3 -c 1 -o $out/${tag}_k_fast_tiled -f python bench.py --steps 2 --warmup 3 > $out/${tag}_ncu_full.log 2>&1
ls -la $out | tail -20
