#!/bin/bash
# Round 2, sixth one-GPU visit: the suite with the full-size C3 cases, the driver's own bench commands.
tag=${1:-r2i}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --durations=8 > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 16 $out/${tag}_pytest.log
timeout 200 python tools/bench_kernels.py > $out/${tag}_kernels.json 2> $out/${tag}_kernels.err; echo "kernel timings rc=$?"; cat $out/${tag}_kernels.json
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err ) 2>&1 | grep real; echo "bench rc=$?"; grep '^{' $out/${tag}_bench.json | cut -c1-400
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err ) 2>&1 | grep real; echo "reference rc=$?"; grep '^{' $out/${tag}_bench_reference.json | cut -c1-300
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 > $out/${tag}_bench_c4.json 2> $out/${tag}_bench_c4.err; echo "c4 rc=$?"; grep '^{' $out/${tag}_bench_c4.json | cut -c1-300
timeout 300 python bench.py --workload bolt --steps 3 --warmup 2 --e2e-steps 1 > $out/${tag}_bench_bolt.json 2> $out/${tag}_bench_bolt.err; echo "bolt rc=$?"; grep '^{' $out/${tag}_bench_bolt.json | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 $out/${tag}_smoke.log
exit 0
