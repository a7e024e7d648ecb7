#!/bin/bash
# NOTE: kept as the record of how profiles/r2/devcheck_b200_xpass_*.txt were produced.  x-pass variants 7 and 9-14 were removed
# from the library after this measurement (they run variant 0 now); 6 and 8 remain.
# Second short GPU call: the original input pack in the 3- and 4-warp CTA shapes (variants 13, 14), the fused assembly / stage
# launch variants 0-4, and one ncu --set full capture of the hoisted-load 2-CTA variant (a middle RK4 stage).
#   gpurun --timeout 240 -- 'bash profiles/r2/xpass_shot2.sh'
source profiles/devcheck_env.sh
mkdir -p gpurun_out
D=tests/native/_build/devcheck
O=gpurun_out/devcheck_xpass_shapes.txt
timeout 100 $D 512 512 $O quick=1 reps=5 vmask=0x6001 > gpurun_out/devcheck_xpass_shapes.log 2>&1
echo "rc=$?" >> gpurun_out/devcheck_xpass_shapes.log
tail -22 $O
timeout 120 ncu --set full --clock-control none --import-source on -k regex:items_kernel_b --launch-skip 9 -c 1 -f -o gpurun_out/prof_r2b_assemble_stage \
    $D 0 512 /dev/null quick=1 reps=0 vmask=0 > gpurun_out/ncu_assemble.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_assemble.log; ls -la gpurun_out/*.ncu-rep
