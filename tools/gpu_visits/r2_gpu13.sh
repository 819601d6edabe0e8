#!/bin/bash
# Round 2, thirteenth one-GPU visit: far-field loop with shared fma(xi, B, A) terms and two
# accumulation chains; bench lines of record, launch lists (C2, bolt with 12 iterations), full ncu
# capture of the G+K assembly launch.
tag=${1:-r2p}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --durations=4 > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 9 $out/${tag}_pytest.log
timeout 300 python tools/bench_kernels.py > $out/${tag}_kernels.json 2> $out/${tag}_kernels.err; echo "kernel timings rc=$?"; cat $out/${tag}_kernels.json
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err ) 2>&1 | grep real; echo "bench rc=$?"
python - $out/${tag}_bench.json <<'PYEOF'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
p = d['phases']
print('ms/step %.2f assembly %.2f solve %.2f iters %d gemv %.4f conv %s berr %.2e e2e %.2f launches %d roofline %.3f' % (
    d['ms_per_step'], p['assembly_ms'], p['solve_ms'], p['cg_iters'], p['gemv_ms'], p['cg_converged'],
    p['cg_backward_error_max_norm'], d['e2e']['ms_per_step'], d['gpu_launches'], d['roofline']['frac']))
PYEOF
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1; echo "ncu launch list rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_launches_bolt.csv \
    python bench.py --workload bolt --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --max-iters 12 > $out/${tag}_ncu_launches_bolt.log 2>&1; echo "ncu launch list bolt rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_offdiag_ff' -s 1 -c 1 \
    -o $out/${tag}_offdiag_ff_gk -f python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > $out/${tag}_ncu_ff.log 2>&1; echo "ncu ff rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_cg_update_dense|k_symv' -s 16 -c 6 \
    -o $out/${tag}_solve -f python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > $out/${tag}_ncu_solve.log 2>&1; echo "ncu solve rc=$?"
exit 0
