#!/bin/bash
# First GPU visit of the next round: the experimental paths that were written without a GPU.
#   tools/gpu_experimental.sh <tag> 1        cap blocks + refinement on one GPU
#   tools/gpu_experimental.sh <tag> 2|8      sharded worker with caps (and the peer exchange), then
#                                            the C3-size bench with the cap-block preconditioner
tag=${1:-exp}; n=${2:-1}
out=gpurun_out
mkdir -p $out
export ICQB_EXPERIMENTAL=1
if [ "$n" = "1" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q -k "cap_blocks or subdivide" > $out/${tag}_exp_pytest.log 2>&1
  echo "experimental pytest rc=$?"; tail -5 $out/${tag}_exp_pytest.log
  exit 0
fi
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 400 bash -c "$(declare -f tr); n=$n; tr 29731 tests/dist_worker.py" > $out/${tag}_exp_worker_nccl.log 2>&1
echo "worker (caps, NCCL) rc=$?"; grep -c " ok" $out/${tag}_exp_worker_nccl.log
ICQB_PEER_EXCHANGE=1 timeout 400 bash -c "$(declare -f tr); n=$n; tr 29732 tests/dist_worker.py" > $out/${tag}_exp_worker_peer.log 2>&1
echo "worker (caps, peer exchange) rc=$?"; grep -c " ok" $out/${tag}_exp_worker_peer.log
for mode in nccl peer; do
  [ $mode = peer ] && export ICQB_PEER_EXCHANGE=1
  timeout 300 bash -c "$(declare -f tr); n=$n; tr 29733 bench.py --gpus $n --steps 2 --warmup 1 --e2e-steps 1 --preconditioner block+caps" \
      > $out/${tag}_exp_bench_weak_${mode}.json 2> $out/${tag}_exp_bench_weak_${mode}.err
  echo "bench weak ($mode) rc=$?"; cat $out/${tag}_exp_bench_weak_${mode}.json
done
unset ICQB_PEER_EXCHANGE
if [ "$n" = "8" ]; then
  timeout 300 bash -c "$(declare -f tr); n=$n; tr 29734 bench.py --gpus $n --steps 2 --warmup 1 --e2e-steps 1 --workload c3" \
      > $out/${tag}_exp_bench_c3.json 2> $out/${tag}_exp_bench_c3.err
  echo "bench c3 rc=$?"; cat $out/${tag}_exp_bench_c3.json
fi
