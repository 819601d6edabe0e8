#!/bin/bash
# Round 2, fourteenth visit (one GPU): the suite, C2 and bolt bench lines after single tiles became
# one-tile dense blocks (the bolt's partition has some: it now takes the fused update kernel too).
tag=${1:-r2q}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --durations=4 > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 9 $out/${tag}_pytest.log
for w in c2 bolt; do
  steps=10; warm=3; if [ $w = bolt ]; then steps=3; warm=2; fi
  timeout 400 python bench.py --gpus 1 --workload $w --steps $steps --warmup $warm --e2e-steps 1 > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err; echo "bench $w rc=$?"
  python - $out/${tag}_bench_$w.json <<'PYEOF'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
p = d['phases']
print('ms/step %.2f assembly %.2f solve %.2f iters %d gemv %.4f conv %s berr %.2e e2e %.2f launches %d roofline %.3f' % (
    d['ms_per_step'], p['assembly_ms'], p['solve_ms'], p['cg_iters'], p['gemv_ms'], p['cg_converged'],
    p['cg_backward_error_max_norm'], d['e2e']['ms_per_step'], d['gpu_launches'], d['roofline']['frac']))
PYEOF
  tail -n 2 $out/${tag}_bench_$w.err
done
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 $out/${tag}_smoke.log
exit 0
