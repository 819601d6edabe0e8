#!/bin/bash
# Round 2, tenth one-GPU visit: source-point loop unrolled once / twice (A/B), full ncu captures of
# the assembly kernel (both) and of the new preconditioner set-up kernels.
tag=${1:-r2l}
out=gpurun_out
mkdir -p $out
for u in 1 2; do
  ICQB_FF_UNROLL=$u timeout 300 python tools/bench_kernels.py > $out/${tag}_kernels_u$u.json 2> $out/${tag}_kernels_u$u.err; echo "kernel timings unroll $u rc=$?"; cat $out/${tag}_kernels_u$u.json
done
for u in 1 2; do
  ICQB_FF_UNROLL=$u timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_offdiag_ff' -s 3 -c 1 \
      -o $out/${tag}_offdiag_ff_u$u -f python tools/bench_kernels.py > $out/${tag}_ncu_ff_u$u.log 2>&1; echo "ncu ff unroll $u rc=$?"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_sweep_update|k_sweep_pivot|k_sweep_panel|k_dense_apply|k_cg_close' -s 20 -c 10 \
    -o $out/${tag}_solve_setup -f python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > $out/${tag}_ncu_setup.log 2>&1; echo "ncu setup rc=$?"
ls -la $out/${tag}_*.ncu-rep
exit 0
