#!/bin/bash
# Collect round-2 evidence on one B200 into gpurun_out/ (copied to profiles/ afterwards).  Usage: bash scripts/collect_r02.sh r02g
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
python -m pytest tests -x -q -m gpu > $out/${tag}_pytest_gpu.txt 2>&1; tail -4 $out/${tag}_pytest_gpu.txt
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; tail -c 600 $out/${tag}_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2>&1
python bench.py --workload flowfield --steps 2 --warmup 3 > $out/${tag}_bench_flowfield_n1.json 2> $out/${tag}_ff.err; tail -c 400 $out/${tag}_ff.err
python bench.py --workload sweep --steps 2 --warmup 3 > $out/${tag}_bench_sweep_n1.json 2> $out/${tag}_sw.err; tail -c 400 $out/${tag}_sw.err
[ -f scripts/_build/libludvm_trace.so ] && (LUDVM_B200_LIB=$PWD/scripts/_build/libludvm_trace.so python scripts/coop_trace.py exact; LUDVM_B200_LIB=$PWD/scripts/_build/libludvm_trace.so python scripts/coop_trace.py fast) > $out/${tag}_coop_trace.txt 2>&1
cat $out/${tag}_coop_trace.txt
# profiler passes last (numbers printed under ncu are never bench values)
python bench.py --steps 2 --warmup 3 --no-extra-legs > $out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/${tag}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extra-legs > $out/${tag}_ncu_bench.log 2>&1
python scripts/run_configs.py --only 1,2,4 --out $out/${tag}_configs.json > $out/${tag}_configs.log 2>&1; tail -4 $out/${tag}_configs.log
python scripts/ncu_fast_driver.py 1048576 > $out/${tag}_plain_fused.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_fast_fused -s 1 -c 1 -o $out/${tag}_k_fast_fused -f python scripts/ncu_fast_driver.py 1048576 > $out/${tag}_ncu_fused.log 2>&1
cat $out/${tag}_plain_fused.txt
# far-field path: probe (time / pairs left / error), launch list and one full capture of the evaluation kernel
PROBE_ORDERS=12,14,16,18 python scripts/tree_probe.py 18 20 22 24 > $out/${tag}_tree_probe.txt 2>&1; cut -c1-200 $out/${tag}_tree_probe.txt
python scripts/ff_tree_probe.py > $out/${tag}_ff_tree_probe.txt 2>&1
python scripts/tree_ncu_driver.py 20 18 > $out/${tag}_plain_tree.txt 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/${tag}_tree_launches.csv python scripts/tree_ncu_driver.py 20 18 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tree_eval -s 1 -c 1 -o $out/${tag}_k_tree_eval -f python scripts/tree_ncu_driver.py 20 18 > $out/${tag}_ncu_tree.log 2>&1
ls -la $out | tail -14
