#!/bin/bash
# Round 2, fourth one-GPU visit: the hand-written LU (direct solve), per-kernel times of the bolt solve.
tag=${1:-r2e}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q -x -k "direct_" > $out/${tag}_pytest_direct.log 2>&1; echo "pytest direct rc=$?"; tail -n 15 $out/${tag}_pytest_direct.log
timeout 600 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 $out/${tag}_pytest.log
for w in small c2 bolt; do
  timeout 400 python tools/bench_lu.py $w > $out/${tag}_lu_$w.json 2> $out/${tag}_lu_$w.err; echo "LU $w rc=$?"; cat $out/${tag}_lu_$w.json; tail -n 3 $out/${tag}_lu_$w.err
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_lu_gemm' -s 40 -c 2 \
    -o $out/${tag}_lu_gemm -f python tools/bench_lu.py c2 > $out/${tag}_ncu_lu_gemm.log 2>&1; echo "ncu lu gemm rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/${tag}_launches_lu.csv python tools/bench_lu.py small > $out/${tag}_ncu_launches_lu.log 2>&1; echo "ncu LU list rc=$?"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bolt.csv python bench.py --workload bolt --steps 1 --warmup 0 --max-iters 12 --no-cpu-baseline --e2e-steps 0 > $out/${tag}_ncu_launches_bolt.log 2>&1; echo "ncu bolt list rc=$?"
exit 0
