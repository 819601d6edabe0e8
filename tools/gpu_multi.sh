#!/bin/bash
# Multi-GPU visit: sharded parity worker + bench at N ranks.  usage: tools/gpu_multi.sh <tag> <nranks> [workloads...]
tag=$1; n=$2; shift 2
out=gpurun_out
mkdir -p $out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q -k "test_sharded_assembly_and_cg[$n]" > $out/${tag}_multi_pytest.log 2>&1; echo "multi pytest rc=$?"; tail -3 $out/${tag}_multi_pytest.log
for w in "${@:-auto}"; do
  timeout 600 bash -c "$(declare -f run); n=$n; run 29711 bench.py --gpus $n --steps 3 --warmup 3 --workload $w" > $out/${tag}_bench_n${n}_${w}.json 2> $out/${tag}_bench_n${n}_${w}.err; echo "bench $w rc=$?"
  cat $out/${tag}_bench_n${n}_${w}.json
done
