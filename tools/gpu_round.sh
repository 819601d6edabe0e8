#!/bin/bash
# One GPU visit: parity tests, bench, sanitizer pass on the smoke case, ncu launch list and
# full captures of the solve kernels.  Outputs under gpurun_out/<tag>_*.
tag=${1:-run}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
tail -5 $out/${tag}_pytest.log
python bench.py --steps 3 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
cat $out/${tag}_bench.json
python bench.py --steps 2 --warmup 3 --preconditioner diagonal --no-cpu-baseline > $out/${tag}_bench_diag.json 2>> $out/${tag}_bench.err; echo "bench diag rc=$?"
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > $out/${tag}_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -4 $out/${tag}_memcheck.log
if [ "$2" != "noncu" ]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > $out/${tag}_ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_symv|k_cg_resid' -s 40 -c 6 -o $out/${tag}_solve -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > $out/${tag}_ncu_solve.log 2>&1; echo "ncu solve rc=$?"
fi
