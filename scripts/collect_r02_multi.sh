#!/bin/bash
# Multi-GPU evidence (run under `gpurun --gpus G`).  Usage: bash scripts/collect_r02_multi.sh r02h 2
tag=${1:-r02h}; G=${2:-2}
out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_multi.py -x -q -m gpu > $out/${tag}_pytest_multi.txt 2>&1; tail -5 $out/${tag}_pytest_multi.txt
$TR --master-port 29511 bench.py --gpus $G --steps 3 --warmup 3 > $out/${tag}_bench_n$G.json 2> $out/${tag}_bench_n$G.err; tail -c 300 $out/${tag}_bench_n$G.err
$TR --master-port 29512 bench.py --gpus $G --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference_n$G.json 2> /dev/null
$TR --master-port 29513 bench.py --gpus $G --workload flowfield --steps 2 --warmup 3 > $out/${tag}_bench_flowfield_n$G.json 2> $out/${tag}_ff.err; tail -c 300 $out/${tag}_ff.err
$TR --master-port 29514 bench.py --gpus $G --workload sweep --steps 2 --warmup 3 > $out/${tag}_bench_sweep_n$G.json 2> $out/${tag}_sw.err; tail -c 300 $out/${tag}_sw.err
$TR --master-port 29515 bench.py --gpus $G --transport nccl --steps 2 --warmup 3 --no-extra-legs > $out/${tag}_bench_n${G}_nccl.json 2> /dev/null
python - <<PY
import json
for f in ["${tag}_bench_n$G.json", "${tag}_bench_reference_n$G.json", "${tag}_bench_flowfield_n$G.json", "${tag}_bench_sweep_n$G.json", "${tag}_bench_n${G}_nccl.json"]:
    try:
        d = json.loads([l for l in open("$out/" + f) if l.startswith("{")][-1])
        print(f, "value %.4g" % d["value"], d.get("parity", {}).get("ok"), d.get("cpu_baseline", {}).get("cores"), d.get("config", {}).get("transport"))
    except Exception as e:
        print(f, "ERR", e)
PY
