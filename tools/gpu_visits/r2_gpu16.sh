#!/bin/bash
# Round 2, sixteenth visit (one GPU): the lines of record with the final code, as the driver runs them.
tag=${1:-r2u}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --durations=4 > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 9 $out/${tag}_pytest.log
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err ) 2>&1 | grep real; echo "bench rc=$?"
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err ) 2>&1 | grep real; echo "reference rc=$?"; grep '^{' $out/${tag}_bench_reference.json | cut -c1-200
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 > $out/${tag}_bench_c4.json 2> $out/${tag}_bench_c4.err; echo "c4 rc=$?"
timeout 300 python bench.py --workload bolt --steps 3 --warmup 2 --e2e-steps 1 > $out/${tag}_bench_bolt.json 2> $out/${tag}_bench_bolt.err; echo "bolt rc=$?"
for w in "" _c4 _bolt; do
python - $out/${tag}_bench$w.json <<'PYEOF'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
p = d['phases']
print('%s ms/step %.2f assembly %.2f solve %.2f iters %d gemv %.4f conv %s berr %.2e e2e %.2f launches %d roofline %.3f gemv frac %.3f' % (
    d['config']['workload'][:24], d['ms_per_step'], p['assembly_ms'], p['solve_ms'], p['cg_iters'], p['gemv_ms'], p['cg_converged'],
    p['cg_backward_error_max_norm'], d['e2e']['ms_per_step'], d['gpu_launches'], d['roofline_assembly']['frac'], d['roofline_gemv']['frac']))
PYEOF
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1; echo "ncu launch list rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 $out/${tag}_smoke.log
exit 0
