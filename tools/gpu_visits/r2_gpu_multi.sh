#!/bin/bash
# Round 2, multi-GPU visit: sharded parity worker (NCCL all-reduce, then the peer-memory exchange),
# then bench.py on the refined bolt (strong scaling) with both exchanges.
# usage: tools/r2_gpu_multi.sh <tag> <nranks> [extra workloads ...]
tag=$1; n=$2; shift 2
out=gpurun_out
mkdir -p $out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    r = d['roofline_assembly']; p = d['phases']; e = d.get('e2e') or {}; o = p['over_ranks']
    print('  N=%d P=%d offdiag max %.1f min %.1f ms frac %.3f | step %.1f ms solve %.1f ms iters %d conv %s bwd %.2e | product max %.3f min %.3f ms %.0f GB/s (heaviest) | e2e %s ms | tiles max/min %.4f' % (
        d['config']['triangles'], d['n_gpus'], o['offdiag_ms']['max'], o['offdiag_ms']['min'], r['frac'],
        d['ms_per_step'], p['solve_ms'], p['cg_iters'], p['cg_converged'], p['cg_backward_error_max_norm'],
        o['gemv_ms']['max'], o['gemv_ms']['min'], p['gemv_gbs_per_gpu'], e.get('ms_per_step'),
        o['stored_bytes']['max'] / max(o['stored_bytes']['min'], 1)))
except Exception as e:
    print('  no bench line:', e)
PY
}
timeout 500 bash -c "$(declare -f run); n=$n; run 29731 tests/dist_worker.py" > $out/${tag}_worker_nccl.log 2>&1
echo "worker (NCCL) rc=$? ok-count $(grep -c ' ok' $out/${tag}_worker_nccl.log)"; grep "ranks=" $out/${tag}_worker_nccl.log | tail -n 30; tail -n 3 $out/${tag}_worker_nccl.log
ICQB_PEER_EXCHANGE=1 timeout 500 bash -c "$(declare -f run); n=$n; run 29732 tests/dist_worker.py" > $out/${tag}_worker_peer.log 2>&1
echo "worker (peer exchange) rc=$? ok-count $(grep -c ' ok' $out/${tag}_worker_peer.log)"; grep "ranks=" $out/${tag}_worker_peer.log | tail -n 4; tail -n 3 $out/${tag}_worker_peer.log
for w in auto "$@"; do
  for mode in nccl peer; do
    if [ $mode = peer ]; then export ICQB_PEER_EXCHANGE=1; else unset ICQB_PEER_EXCHANGE; fi
    timeout 400 bash -c "$(declare -f run); n=$n; run 29733 bench.py --gpus $n --steps 3 --warmup 2 --e2e-steps 1 --workload $w --no-cpu-baseline" \
        > $out/${tag}_bench_n${n}_${w}_${mode}.json 2> $out/${tag}_bench_n${n}_${w}_${mode}.err
    echo "bench $w ($mode) rc=$?"; summ $out/${tag}_bench_n${n}_${w}_${mode}.json; tail -n 2 $out/${tag}_bench_n${n}_${w}_${mode}.err
  done
done
unset ICQB_PEER_EXCHANGE
# the direct solve (csrc/lu.cu) over the ranks
for w in c2 bolt; do
  timeout 300 bash -c "$(declare -f run); n=$n; run 29734 tools/bench_lu.py $w" > $out/${tag}_lu_n${n}_$w.json 2> $out/${tag}_lu_n${n}_$w.err
  echo "LU $w rc=$?"; grep '^{' $out/${tag}_lu_n${n}_$w.json; tail -n 2 $out/${tag}_lu_n${n}_$w.err
done
exit 0
