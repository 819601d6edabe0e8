#!/bin/bash
# Round 2, third one-GPU visit: reworked far-field kernel and dense-block apply.
tag=${1:-r2d}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q -x > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 $out/${tag}_pytest.log
timeout 200 python tools/bench_kernels.py > $out/${tag}_kernels.json 2> $out/${tag}_kernels.err; echo "kernel timings rc=$?"; cat $out/${tag}_kernels.json
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    r = d['roofline_assembly']; p = d['phases']; e = d.get('e2e') or {}
    print('  N=%d offdiag %.2f ms %.2f TF/s frac %.3f ff %s | step %.1f ms solve %.1f ms iters %d conv %s bwd %.2e gemv %.3f ms %.0f GB/s | e2e %s ms | G-only %.1f ms' % (
        d['config']['triangles'], r['ms_per_launch'], r['achieved'], r['frac'], r.get('far_field_fraction'),
        d['ms_per_step'], p['solve_ms'], p['cg_iters'], p['cg_converged'], p['cg_backward_error_max_norm'], p['gemv_ms'], p['gemv_gbs_per_gpu'],
        e.get('ms_per_step'), d['g_only']['ms']))
except Exception as e:
    print('  no bench line:', e)
PY
}
for w in "c2 0" "c2 1" "bolt 0" "bolt 1"; do
  set -- $w
  timeout 300 python bench.py --workload $1 --assembly-variant $2 --steps 3 --warmup 2 --e2e-steps 1 --no-cpu-baseline \
      > $out/${tag}_bench_$1_v$2.json 2> $out/${tag}_bench_$1_v$2.err
  echo "bench $1 variant $2 rc=$?"; summ $out/${tag}_bench_$1_v$2.json; tail -n 2 $out/${tag}_bench_$1_v$2.err
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > $out/${tag}_ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_offdiag_ff' -s 1 -c 1 \
    -o $out/${tag}_offdiag_v1 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 \
    --assembly-variant 1 > $out/${tag}_ncu_offdiag_v1.log 2>&1; echo "ncu offdiag v1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_dense_apply|k_cg_close|k_cg_vec|k_symv' -s 60 -c 8 \
    -o $out/${tag}_solve -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 \
    > $out/${tag}_ncu_solve.log 2>&1; echo "ncu solve rc=$?"
exit 0
