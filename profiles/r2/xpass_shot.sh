#!/bin/bash
# NOTE: kept as the record of how profiles/r2/devcheck_b200_xpass_*.txt were produced.  x-pass variants 7 and 9-14 were removed
# from the library after this measurement (they run variant 0 now); 6 and 8 remain.
# One short GPU call (native checker only, no Python): x-pass input-pack variants 6-12 and the fused assembly / stage launch
# variants against the defaults at 512^3, then the RK4 step with the fastest of each.
#   gpurun --timeout 240 -- 'bash profiles/r2/xpass_shot.sh'
source profiles/devcheck_env.sh
mkdir -p gpurun_out
D=tests/native/_build/devcheck
O=gpurun_out/devcheck_xpass_pack.txt
timeout 120 $D 512 512 $O quick=1 reps=5 vmask=0x1fc9 > gpurun_out/devcheck_xpass_pack.log 2>&1
echo "rc=$?" >> gpurun_out/devcheck_xpass_pack.log
BX=$(grep "ddl_rhs, x-pass variant" $O | sed -E 's/.*variant ([0-9]+):.*"ms": ([0-9.]+)\}.*/\2 \1/' | sort -n | head -1 | awk '{print $2}')
BA=$(grep "RK4 step, assemble_variant" $O | sed -E 's/.*assemble_variant ([0-9]+):.*"ms": ([0-9.]+)\}.*/\2 \1/' | sort -n | head -1 | awk '{print $2}')
echo "fastest: xfused_variant=$BX assemble_variant=$BA" | tee -a gpurun_out/devcheck_xpass_pack.log
if [ -n "$BX" ] && [ -n "$BA" ]; then
  timeout 90 $D 512 512 gpurun_out/devcheck_xpass_best.txt quick=2 reps=5 vmask=$((1 << BX)) xfused_variant=$BX assemble_variant=$BA >> gpurun_out/devcheck_xpass_pack.log 2>&1
fi
tail -45 $O
tail -8 gpurun_out/devcheck_xpass_best.txt
