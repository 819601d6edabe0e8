#!/bin/bash
# Round 2, fifteenth visit (one GPU): fused update kernel with 16 rows per CTA and the matrix loads
# issued before the vector is staged.
tag=${1:-r2s}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q -x -k "cap_blocks or block_tiles or c2_full_size or laplace_response or mixed or c3_full" > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $out/${tag}_pytest.log
timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python - $out/${tag}_bench.json <<'PYEOF'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
p = d['phases']
print('ms/step %.2f assembly %.2f solve %.2f iters %d gemv %.4f conv %s berr %.2e e2e %.2f launches %d' % (
    d['ms_per_step'], p['assembly_ms'], p['solve_ms'], p['cg_iters'], p['gemv_ms'], p['cg_converged'],
    p['cg_backward_error_max_norm'], d['e2e']['ms_per_step'], d['gpu_launches']))
PYEOF
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:'k_cg_update_dense|k_symv_reduce' -s 20 -c 6 --csv --log-file $out/${tag}_update.csv \
    python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > $out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
grep -v "^==" $out/${tag}_update.csv | cut -d, -f5,13- | head -14
exit 0
