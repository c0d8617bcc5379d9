#!/bin/bash
# Reduced end-of-round collection on one B200 (the full one is collect_r02.sh).  Usage: bash scripts/collect_final.sh r03g
tag=${1:-r03g}; out=gpurun_out; mkdir -p $out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -x -q -m gpu > $out/${tag}_pytest_gpu.txt 2>&1; tail -2 $out/${tag}_pytest_gpu.txt
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; tail -c 300 $out/${tag}_bench_n1.err
python bench.py --workload flowfield --steps 2 --warmup 3 > $out/${tag}_bench_flowfield_n1.json 2> $out/${tag}_ff.err; tail -c 300 $out/${tag}_ff.err
PROBE_ORDERS=12,14,16,18 python scripts/tree_probe.py 18 20 22 24 > $out/${tag}_tree_probe.txt 2>&1; cut -c1-150 $out/${tag}_tree_probe.txt
python scripts/ff_tree_probe.py > $out/${tag}_ff_tree_probe.txt 2>&1; cut -c1-200 $out/${tag}_ff_tree_probe.txt | head -3
python scripts/tree_ncu_driver.py 20 18 > $out/${tag}_plain_tree.txt 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 24 --csv --log-file $out/${tag}_tree_launches.csv python scripts/tree_ncu_driver.py 20 18 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tree_m2l_dmma -s 1 -c 1 -o $out/${tag}_k_tree_m2l_dmma -f python scripts/tree_ncu_driver.py 20 18 > $out/${tag}_ncu_gemm.log 2>&1
ls -la $out | tail -8
