#!/bin/bash
# Round 2, twelfth one-GPU visit: fused per-iteration update kernel (A/B against the three-kernel
# update), K epilogue as three moment sums.
tag=${1:-r2n}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --durations=4 > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 9 $out/${tag}_pytest.log
timeout 300 python tools/bench_kernels.py > $out/${tag}_kernels.json 2> $out/${tag}_kernels.err; echo "kernel timings rc=$?"; cat $out/${tag}_kernels.json
for mode in fused unfused; do
  if [ $mode = unfused ]; then export ICQB_CG_UNFUSED=1; else unset ICQB_CG_UNFUSED; fi
  timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_$mode.json 2> $out/${tag}_bench_$mode.err; echo "bench $mode rc=$?"
  python - $out/${tag}_bench_$mode.json <<'PYEOF'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
p = d['phases']
print('ms/step %.2f assembly %.2f solve %.2f iters %d gemv %.4f conv %s berr %.2e e2e %.2f launches %d' % (
    d['ms_per_step'], p['assembly_ms'], p['solve_ms'], p['cg_iters'], p['gemv_ms'], p['cg_converged'],
    p['cg_backward_error_max_norm'], d['e2e']['ms_per_step'], d['gpu_launches']))
PYEOF
done
unset ICQB_CG_UNFUSED
timeout 300 python bench.py --workload bolt --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline > $out/${tag}_bench_bolt.json 2> $out/${tag}_bench_bolt.err; echo "bolt rc=$?"
python - $out/${tag}_bench_bolt.json <<'PYEOF'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
p = d['phases']
print('bolt ms/step %.2f assembly %.2f solve %.2f iters %d gemv %.4f conv %s berr %.2e' % (
    d['ms_per_step'], p['assembly_ms'], p['solve_ms'], p['cg_iters'], p['gemv_ms'], p['cg_converged'],
    p['cg_backward_error_max_norm']))
PYEOF
exit 0
