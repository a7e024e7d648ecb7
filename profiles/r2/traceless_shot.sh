#!/bin/bash
# If any GPU seconds are left: traceless_flux 0 / 1 side by side at 512^3 (native checker, ~8 s of run time)
source profiles/devcheck_env.sh
mkdir -p gpurun_out
timeout 14 tests/native/_build/devcheck 0 512 gpurun_out/devcheck_traceless.txt quick=1 reps=3 vmask=0 > /dev/null 2>&1
echo "rc=$?"; grep "traceless_flux\|RK4 step, fused\|per-kernel" gpurun_out/devcheck_traceless.txt | cut -c1-400
