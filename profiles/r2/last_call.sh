#!/bin/bash
# The round's last GPU seconds: ncu --set full of the four two-stage strided launches of one RHS (z_inv, y_inv, y_fwd, z_fwd)
# from the native checker, then the timed region of bench.py (no parity block, no end-to-end leg, no CPU baseline).
source profiles/devcheck_env.sh
mkdir -p gpurun_out
timeout 40 ncu --set full --clock-control none --import-source on -k regex:strided_two -c 4 -f -o gpurun_out/prof_r2b_strided_two \
    tests/native/_build/devcheck 0 512 /dev/null quick=1 reps=0 vmask=0 > gpurun_out/ncu_strided_two.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_r2b_strided_two.ncu-rep
timeout 32 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/bench_quick_r2b.json 2> gpurun_out/bench_quick_r2b.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_quick_r2b.json
