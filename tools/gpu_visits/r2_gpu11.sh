#!/bin/bash
# Round 2, eleventh one-GPU visit: the suite and the bench lines of record with the final kernels of the
# assembly (pipelined far-field loop, coalesced stores) and of the preconditioner set-up (block
# sweep, 512-thread register-resident pivot tiles); launch list and full ncu capture of the G+K
# assembly launch of the timed step.
tag=${1:-r2m}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --durations=6 > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 12 $out/${tag}_pytest.log
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err ) 2>&1 | grep real; echo "bench rc=$?"; grep '^{' $out/${tag}_bench.json | cut -c1-300
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 > $out/${tag}_bench_c4.json 2> $out/${tag}_bench_c4.err; echo "c4 rc=$?"; grep '^{' $out/${tag}_bench_c4.json | cut -c1-300
timeout 300 python bench.py --workload bolt --steps 3 --warmup 2 --e2e-steps 1 > $out/${tag}_bench_bolt.json 2> $out/${tag}_bench_bolt.err; echo "bolt rc=$?"; grep '^{' $out/${tag}_bench_bolt.json | cut -c1-300
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1; echo "ncu launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_offdiag_ff' -s 1 -c 1 \
    -o $out/${tag}_offdiag_ff_gk -f python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > $out/${tag}_ncu_ff.log 2>&1; echo "ncu ff rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_sweep|k_symv' -s 16 -c 8 \
    -o $out/${tag}_solve -f python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > $out/${tag}_ncu_solve.log 2>&1; echo "ncu solve rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 $out/${tag}_smoke.log
exit 0
