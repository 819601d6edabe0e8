#!/bin/bash
# First GPU visit of the next round: the experimental paths that were written without a GPU.
#   tools/gpu_experimental.sh <tag> 1        cap blocks + refinement on one GPU
#   tools/gpu_experimental.sh <tag> 2|8      sharded worker with caps (and the peer exchange), then
#                                            the C3-size bench with the cap-block preconditioner
tag=${1:-exp}; n=${2:-1}
out=gpurun_out
mkdir -p $out
export ICQB_EXPERIMENTAL=1
if [ "$n" = "1" ]; then
  # one pytest run per experimental feature: a failure in one must not hide the others
  for grp in subdivide cap_blocks block_tiles direct_ float32 mixed_precision; do
    timeout 400 python -m pytest tests -m gpu -q -k "$grp" > $out/${tag}_exp_pytest_$grp.log 2>&1
    echo "experimental pytest [$grp] rc=$?"; tail -3 $out/${tag}_exp_pytest_$grp.log
  done
  # far-field split of the off-diagonal assembly (assemble_offdiag_ff.cu): parity, then the
  # three variants side by side (same workload, same run), then one full capture of variant 1
  timeout 900 python -m pytest tests -m gpu -q -k "far_field" > $out/${tag}_exp_ff_pytest.log 2>&1
  echo "far-field pytest rc=$?"; tail -5 $out/${tag}_exp_ff_pytest.log
  for v in 0 1 2 3 4; do
    timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --assembly-variant $v \
        > $out/${tag}_exp_bench_ff$v.json 2> $out/${tag}_exp_bench_ff$v.err
    echo "bench variant $v rc=$?"; python - <<PY
import json
try:
    d = json.load(open('$out/${tag}_exp_bench_ff$v.json'))
    r = d['roofline_assembly']
    print('variant $v: offdiag %.2f ms, %.2f TFLOP/s = %.1f %% of peak, far-field fraction %s, step %.1f ms, cg %d' % (
        r['ms_per_launch'], r['achieved'], 100 * r['frac'], r.get('far_field_fraction'), d['ms_per_step'], d['phases']['cg_iters']))
except Exception as e:
    print('no bench line:', e)
PY
  done
  # the WHOLE validated parity suite against the experimental kernels: first the far-field
  # assembly alone (every context starts with variant 1), then all candidates for the next
  # defaults at once (far-field assembly, cap/8-tile blocks, float32 late products)
  ICQB_ASSEMBLY_VARIANT=1 timeout 900 python -m pytest tests -m gpu -q > $out/${tag}_exp_suite_variant1.log 2>&1
  echo "parity suite with ICQB_ASSEMBLY_VARIANT=1 rc=$?"; tail -4 $out/${tag}_exp_suite_variant1.log
  ICQB_DEFAULTS=next timeout 900 python -m pytest tests -m gpu -q > $out/${tag}_exp_suite_next.log 2>&1
  echo "parity suite with ICQB_DEFAULTS=next rc=$?"; tail -4 $out/${tag}_exp_suite_next.log
  # memcheck and racecheck on one small case per new kernel (the emulation cannot see real races)
  for tool in memcheck racecheck; do
    timeout 600 compute-sanitizer --tool $tool python -m pytest -q -m gpu \
        "tests/test_gpu_parity.py::test_far_field_split_every_entry[16-9-1]" \
        "tests/test_gpu_parity.py::test_float32_product_kernel[16-9]" \
        "tests/test_gpu_parity.py::test_cap_blocks_same_answer_on_regular_sphere" \
        > $out/${tag}_exp_sanitizer_$tool.log 2>&1
    echo "compute-sanitizer $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $out/${tag}_exp_sanitizer_$tool.log | tail -4
  done
  for v in 0 1; do
    timeout 300 python tools/bench_c4.py 3 $v > $out/${tag}_exp_c4_variant$v.json 2> $out/${tag}_exp_c4_variant$v.err
    echo "config C4 (Poisson, hollow sphere) variant $v rc=$?"; cat $out/${tag}_exp_c4_variant$v.json
  done
  timeout 300 python tools/bench_kernels.py > $out/${tag}_exp_kernels.json 2> $out/${tag}_exp_kernels.err
  echo "kernel timings rc=$?"; cat $out/${tag}_exp_kernels.json
  # the solver's options one after the other on config C2 (block+caps = the experimental solver)
  for opt in "--preconditioner block+caps" "--preconditioner block+caps --block-tiles 8" \
             "--preconditioner block+caps --block-tiles 8 --mixed-switch 1e-6"; do
    name=$(echo $opt | tr -d ' +-' | tr -c 'a-zA-Z0-9\n' '_')
    timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 $opt \
        > $out/${tag}_exp_bench_$name.json 2> $out/${tag}_exp_bench_$name.err
    echo "bench $opt rc=$?"; python - <<PY
import json
try:
    d = json.load(open('$out/${tag}_exp_bench_$name.json'))
    p = d['phases']
    print('$opt: step %.1f ms, solve %.1f ms, %d iterations (%d float32 products), product %.3f ms, converged %s' % (
        d['ms_per_step'], p['solve_ms'], p['cg_iters'], p.get('cg_float32_products', 0), p['gemv_ms'], p['cg_converged']))
except Exception as e:
    print('no bench line:', e)
PY
  done
  # stream-ordered pool allocation of the context's buffers: end-to-end time with and without
  for pool in 0 1; do
    ICQB_POOL_ALLOC=$pool timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 4 \
        > $out/${tag}_exp_bench_pool$pool.json 2> $out/${tag}_exp_bench_pool$pool.err
    echo "bench ICQB_POOL_ALLOC=$pool rc=$?"; python - <<PY
import json
try:
    d = json.load(open('$out/${tag}_exp_bench_pool$pool.json'))
    print('pool $pool: e2e %.2f ms, step %.2f ms, host phases %s' % (d['e2e']['ms_per_step'], d['ms_per_step'], d['e2e'].get('host_phases_ms')))
except Exception as e:
    print('no bench line:', e)
PY
  done
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_offdiag_ff' -s 1 -c 1 \
      -o $out/${tag}_exp_offdiag_ff1 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 \
      --assembly-variant 1 > $out/${tag}_exp_ncu_offdiag_ff1.log 2>&1; echo "ncu offdiag_ff rc=$?"
  exit 0
fi
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 400 bash -c "$(declare -f tr); n=$n; tr 29731 tests/dist_worker.py" > $out/${tag}_exp_worker_nccl.log 2>&1
echo "worker (caps, NCCL) rc=$?"; grep -c " ok" $out/${tag}_exp_worker_nccl.log
ICQB_PEER_EXCHANGE=1 timeout 400 bash -c "$(declare -f tr); n=$n; tr 29732 tests/dist_worker.py" > $out/${tag}_exp_worker_peer.log 2>&1
echo "worker (caps, peer exchange) rc=$?"; grep -c " ok" $out/${tag}_exp_worker_peer.log
for mode in nccl peer; do
  [ $mode = peer ] && export ICQB_PEER_EXCHANGE=1
  timeout 300 bash -c "$(declare -f tr); n=$n; tr 29733 bench.py --gpus $n --steps 2 --warmup 1 --e2e-steps 1 --preconditioner block+caps" \
      > $out/${tag}_exp_bench_weak_${mode}.json 2> $out/${tag}_exp_bench_weak_${mode}.err
  echo "bench weak ($mode) rc=$?"; cat $out/${tag}_exp_bench_weak_${mode}.json
done
unset ICQB_PEER_EXCHANGE
if [ "$n" = "8" ]; then
  timeout 300 bash -c "$(declare -f tr); n=$n; tr 29734 bench.py --gpus $n --steps 2 --warmup 1 --e2e-steps 1 --workload c3" \
      > $out/${tag}_exp_bench_c3.json 2> $out/${tag}_exp_bench_c3.err
  echo "bench c3 rc=$?"; cat $out/${tag}_exp_bench_c3.json
fi
