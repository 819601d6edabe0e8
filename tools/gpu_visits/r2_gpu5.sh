#!/bin/bash
# Round 2, fifth one-GPU visit: LU after the kernel fixes, new defaults (far-field assembly), pivot checks.
tag=${1:-r2f}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 $out/${tag}_pytest.log
for w in c2 bolt; do
  timeout 400 python tools/bench_lu.py $w > $out/${tag}_lu_$w.json 2> $out/${tag}_lu_$w.err; echo "LU $w rc=$?"; cat $out/${tag}_lu_$w.json; tail -n 3 $out/${tag}_lu_$w.err
done
timeout 300 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench default rc=$?"; cat $out/${tag}_bench.json | cut -c1-600
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_lu_gemm|k_lu_col' -s 300 -c 3 \
    -o $out/${tag}_lu_kernels -f python tools/bench_lu.py c2 > $out/${tag}_ncu_lu.log 2>&1; echo "ncu lu rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/${tag}_launches_lu.csv python tools/bench_lu.py small > $out/${tag}_ncu_launches_lu.log 2>&1; echo "ncu LU list rc=$?"
exit 0
