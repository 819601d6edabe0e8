#!/bin/bash
# Round 2, second GPU visit (one GPU): the suite with the new defaults, the bench contract lines,
# C3-size problems on one GPU, full ncu captures of the two assembly kernels.
tag=${1:-r2b}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 6 $out/${tag}_pytest.log
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    r = d['roofline_assembly']; p = d['phases']; e = d.get('e2e') or {}
    print('  N=%d offdiag %.2f ms %.2f TF/s frac %.3f | step %.1f ms solve %.1f ms iters %d conv %s bwd %.2e gemv %.3f ms %.0f GB/s | e2e %s ms | G-only %.1f ms' % (
        d['config']['triangles'], r['ms_per_launch'], r['achieved'], r['frac'],
        d['ms_per_step'], p['solve_ms'], p['cg_iters'], p['cg_converged'], p['cg_backward_error_max_norm'], p['gemv_ms'], p['gemv_gbs_per_gpu'],
        e.get('ms_per_step'), d['g_only']['ms']))
except Exception as e:
    print('  no bench line:', e)
PY
}
timeout 300 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench default rc=$?"; summ $out/${tag}_bench.json; tail -n 3 $out/${tag}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; echo "bench reference rc=$?"; cat $out/${tag}_bench_reference.json; tail -n 3 $out/${tag}_bench_reference.err
for w in "bolt 8" "bolt 16" "c3 8" "c4 8" "c2 1" "c2 4" "c2 16"; do
  set -- $w
  timeout 300 python bench.py --workload $1 --block-tiles $2 --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline --max-iters 600 \
      > $out/${tag}_bench_$1_bt$2.json 2> $out/${tag}_bench_$1_bt$2.err
  echo "bench $1 bt=$2 rc=$?"; summ $out/${tag}_bench_$1_bt$2.json; tail -n 2 $out/${tag}_bench_$1_bt$2.err
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_offdiag' -s 1 -c 1 \
    -o $out/${tag}_offdiag_v0 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 \
    > $out/${tag}_ncu_offdiag_v0.log 2>&1; echo "ncu offdiag v0 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_offdiag_ff' -s 1 -c 1 \
    -o $out/${tag}_offdiag_v1 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 \
    --assembly-variant 1 > $out/${tag}_ncu_offdiag_v1.log 2>&1; echo "ncu offdiag v1 rc=$?"
exit 0
