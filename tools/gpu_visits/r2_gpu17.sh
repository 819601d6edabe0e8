#!/bin/bash
# Round 2, seventeenth visit (one GPU, the last minutes of the budget): overlapping preconditioner
# blocks (blocks of 4 tiles + 1 tile on either side, the one-rank default): the suite, then the C2
# line of record and the bolt.
tag=${1:-r2x}
out=gpurun_out
mkdir -p $out
export ICQB_BLOCK_OVERLAP=1   # (at the time of the visit the overlapping blocks were opt-in; they became the one-rank default after it passed)
timeout 150 python -m pytest tests -m gpu -q -x > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 $out/${tag}_pytest.log
timeout 60 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
timeout 60 python bench.py --workload bolt --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline > $out/${tag}_bench_bolt.json 2> $out/${tag}_bench_bolt.err; echo "bolt rc=$?"
for w in "" _bolt; do
python - $out/${tag}_bench$w.json <<'PYEOF'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    p = d['phases']
    print('%s ms/step %.2f assembly %.2f solve %.2f iters %d gemv %.4f conv %s berr %.2e e2e %.2f launches %d' % (
        d['config']['workload'][:24], d['ms_per_step'], p['assembly_ms'], p['solve_ms'], p['cg_iters'], p['gemv_ms'],
        p['cg_converged'], p['cg_backward_error_max_norm'], d['e2e']['ms_per_step'], d['gpu_launches']))
except Exception as e:
    print('no line', e)
PYEOF
done
tail -n 3 $out/${tag}_bench.err $out/${tag}_bench_bolt.err
exit 0
