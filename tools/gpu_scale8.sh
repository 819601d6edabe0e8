#!/bin/bash
# 8-GPU visit: bench of the weak-scaling workload and of the C3/C5 sizes.  usage: tools/gpu_scale8.sh <tag>
tag=$1; n=8
out=gpurun_out
mkdir -p $out
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $n "${@:2}"; }
timeout 200 bash -c "$(declare -f tr); n=$n; tr 29721 --steps 3 --warmup 3 --workload auto" > $out/${tag}_bench_n8_auto.json 2> $out/${tag}_bench_n8_auto.err; echo "auto rc=$?"; cat $out/${tag}_bench_n8_auto.json
timeout 200 bash -c "$(declare -f tr); n=$n; tr 29724 --steps 2 --warmup 1 --e2e-steps 1 --max-iters 1500 --workload bolt" > $out/${tag}_bench_n8_bolt.json 2> $out/${tag}_bench_n8_bolt.err; echo "bolt (config C3 geometry) rc=$?"; cat $out/${tag}_bench_n8_bolt.json
timeout 200 bash -c "$(declare -f tr); n=$n; tr 29722 --steps 2 --warmup 1 --e2e-steps 1 --workload c3" > $out/${tag}_bench_n8_c3.json 2> $out/${tag}_bench_n8_c3.err; echo "c3 rc=$?"; cat $out/${tag}_bench_n8_c3.json
timeout 260 bash -c "$(declare -f tr); n=$n; tr 29723 --steps 2 --warmup 1 --e2e-steps 1 --workload c5" > $out/${tag}_bench_n8_c5.json 2> $out/${tag}_bench_n8_c5.err; echo "c5 rc=$?"; cat $out/${tag}_bench_n8_c5.json
tail -3 $out/${tag}_bench_n8_c5.err
