#!/bin/bash
# One GPU visit: parity tests, bench, ncu launch list and full captures of the kernels.
# Outputs under gpurun_out/<tag>_*.   usage: tools/gpu_round.sh <tag> [noncu]
tag=${1:-run}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
tail -5 $out/${tag}_pytest.log
python bench.py --steps 5 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
cat $out/${tag}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2>> $out/${tag}_bench.err; echo "bench reference rc=$?"
cat $out/${tag}_bench_reference.json
# config C3's own geometry (refined bolt, 84,600 triangles) on this one GPU, validated path
# (G, K and the transposed K part take 86 GB: no end-to-end leg, which would hold a second set)
timeout 300 python bench.py --workload bolt --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline --max-iters 1500 > $out/${tag}_bench_bolt.json 2> $out/${tag}_bench_bolt.err; echo "bench bolt rc=$?"
cat $out/${tag}_bench_bolt.json
if [ "$2" != "noncu" ]; then
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > $out/${tag}_ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_symv|k_cg_resid' -s 40 -c 6 -o $out/${tag}_solve -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > $out/${tag}_ncu_solve.log 2>&1; echo "ncu solve rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_offdiag' -s 1 -c 1 -o $out/${tag}_offdiag -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > $out/${tag}_ncu_offdiag.log 2>&1; echo "ncu offdiag rc=$?"
fi
