#!/bin/bash
# Round 2, ninth one-GPU visit: block-sweep inversion of the preconditioner blocks (register-resident
# pivot tiles), row-coalesced tile stores and the triangular grid of the assembly kernel, the
# software-pipelined variant 2, the block cache behind the contexts' device buffers.
tag=${1:-r2k}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x --durations=6 > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 14 $out/${tag}_pytest.log
timeout 300 python tools/bench_kernels.py > $out/${tag}_kernels.json 2> $out/${tag}_kernels.err; echo "kernel timings rc=$?"; cat $out/${tag}_kernels.json; tail -n 3 $out/${tag}_kernels.err
for v in 1 2; do
  timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --assembly-variant $v --no-cpu-baseline > $out/${tag}_bench_v$v.json 2> $out/${tag}_bench_v$v.err; echo "bench v$v rc=$?"
  python - $out/${tag}_bench_v$v.json <<'EOF'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
p = d['phases']
print('ms/step %.2f assembly %.2f solve %.2f iters %d gemv %.4f conv %s e2e %.2f' % (
    d['ms_per_step'], p['assembly_ms'], p['solve_ms'], p['cg_iters'], p['gemv_ms'], p['cg_converged'],
    d['e2e']['ms_per_step']))
print(json.dumps(d['e2e']['host_phases_ms']))
print('roofline', d['roofline']['frac'], 'gemv', d['roofline_gemv']['frac'])
EOF
  tail -n 3 $out/${tag}_bench_v$v.err
done
# launch list of the default step, and DRAM traffic of the assembly kernel (coalesced stores)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1; echo "ncu launch list rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:'k_offdiag_ff' -c 4 --csv --log-file $out/${tag}_offdiag_traffic.csv \
    python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > $out/${tag}_ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
grep -v "^==" $out/${tag}_offdiag_traffic.csv | cut -d, -f5,12- | head -30
exit 0
