#!/bin/bash
# Round 2, first GPU visit (one GPU): the experimental solver/assembly paths on hardware,
# then the C3-size problems on ONE GPU with a low iteration cap.
tag=${1:-r2a}
out=gpurun_out
mkdir -p $out
export ICQB_EXPERIMENTAL=1
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
for grp in cap_blocks far_field block_tiles subdivide float32 mixed_precision; do
  timeout 300 python -m pytest tests -m gpu -q -k "$grp" > $out/${tag}_pytest_$grp.log 2>&1
  echo "pytest [$grp] rc=$?"; tail -n 3 $out/${tag}_pytest_$grp.log
done
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    r = d['roofline_assembly']; p = d['phases']
    print('  N=%d offdiag %.2f ms %.2f TF/s frac %.3f ff %s | step %.1f ms solve %.1f ms iters %d conv %s bwd %.2e gemv %.3f ms %.0f GB/s' % (
        d['config']['triangles'], r['ms_per_launch'], r['achieved'], r['frac'], r.get('far_field_fraction'),
        d['ms_per_step'], p['solve_ms'], p['cg_iters'], p['cg_converged'], p['cg_backward_error_max_norm'], p['gemv_ms'], p['gemv_gbs_per_gpu']))
except Exception as e:
    print('  no bench line:', e)
PY
}
for v in 0 1 2 3 4; do
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 --assembly-variant $v \
      > $out/${tag}_bench_ff$v.json 2> $out/${tag}_bench_ff$v.err
  echo "bench variant $v rc=$?"; summ $out/${tag}_bench_ff$v.json
done
ICQB_ASSEMBLY_VARIANT=1 timeout 600 python -m pytest tests -m gpu -q > $out/${tag}_suite_variant1.log 2>&1
echo "parity suite with ICQB_ASSEMBLY_VARIANT=1 rc=$?"; tail -n 4 $out/${tag}_suite_variant1.log
for opt in "--preconditioner block+caps" "--preconditioner block+caps --block-tiles 8"; do
  name=$(echo $opt | tr -d ' +-' | tr -c 'a-zA-Z0-9\n' '_')
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 $opt \
      > $out/${tag}_bench_$name.json 2> $out/${tag}_bench_$name.err
  echo "bench c2 $opt rc=$?"; summ $out/${tag}_bench_$name.json
done
for w in "bolt block" "bolt block+caps" "c3 block+caps" "c3 block"; do
  set -- $w
  timeout 240 python bench.py --workload $1 --preconditioner $2 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --max-iters 600 \
      > $out/${tag}_bench_$1_$2.json 2> $out/${tag}_bench_$1_$2.err
  echo "bench $1 $2 rc=$?"; summ $out/${tag}_bench_$1_$2.json; tail -n 2 $out/${tag}_bench_$1_$2.err
done
timeout 200 python tools/bench_kernels.py > $out/${tag}_kernels.json 2> $out/${tag}_kernels.err
echo "kernel timings rc=$?"; cat $out/${tag}_kernels.json
exit 0
