#!/bin/bash
# Third short GPU call: the two-stage strided pass (csrc/fast_two.cuh; strided_two = 1: twiddles generated, 2: loaded) against
# the three-stage default at 512^3, then the RK4 step with the fastest.
#   gpurun --timeout 200 -- 'bash profiles/r2/xpass_shot3.sh'
source profiles/devcheck_env.sh
mkdir -p gpurun_out
D=tests/native/_build/devcheck
O=gpurun_out/devcheck_strided_two.txt
timeout 90 $D 512 512 $O quick=1 reps=5 vmask=0x1 > gpurun_out/devcheck_strided_two.log 2>&1
echo "rc=$?" >> gpurun_out/devcheck_strided_two.log
BT=$(grep "ddl_rhs, strided_two" $O | sed -E 's/.*strided_two ([0-9]+): ([0-9.]+) ms.*/\2 \1/' | sort -n | head -1 | awk '{print $2}')
echo "fastest: strided_two=$BT" | tee -a gpurun_out/devcheck_strided_two.log
if [ -n "$BT" ] && [ "$BT" != "0" ]; then
  timeout 60 $D 0 512 gpurun_out/devcheck_strided_two_best.txt quick=1 reps=5 vmask=0x1 strided_two=$BT >> gpurun_out/devcheck_strided_two.log 2>&1
fi
grep -v "assemble_variant" $O | cut -c1-420 | tail -14
tail -4 gpurun_out/devcheck_strided_two_best.txt 2>/dev/null | cut -c1-420
