#!/bin/bash
# Round 2, last multi-GPU visits (final kernels): bench.py on the refined bolt (strong scaling, NCCL
# exchange) as the driver launches it, optionally further workloads and the sharded parity worker.
# usage: tools/gpu_visits/r2_gpu_multi_final.sh <tag> <nranks> [worker] [workloads ...]
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
if [ "$1" = worker ]; then
  shift
  timeout 500 bash -c "$(declare -f run); n=$n; run 29731 tests/dist_worker.py" > $out/${tag}_worker_nccl.log 2>&1
  echo "worker (NCCL) rc=$? ok-count $(grep -c ' ok' $out/${tag}_worker_nccl.log)"; grep "ranks=" $out/${tag}_worker_nccl.log | tail -n 30; tail -n 3 $out/${tag}_worker_nccl.log
fi
for w in auto "$@"; do
  steps=3; warm=2; if [ $w = c5 ]; then steps=2; warm=1; fi
  timeout 400 bash -c "$(declare -f run); n=$n; run 29733 bench.py --gpus $n --steps $steps --warmup $warm --e2e-steps 1 --workload $w --max-iters 800" \
      > $out/${tag}_bench_n${n}_${w}.json 2> $out/${tag}_bench_n${n}_${w}.err
  echo "bench $w rc=$?"; summ $out/${tag}_bench_n${n}_${w}.json; tail -n 2 $out/${tag}_bench_n${n}_${w}.err
done
exit 0
