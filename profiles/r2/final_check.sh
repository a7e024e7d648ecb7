#!/bin/bash
# Last GPU call of the round (about 100 s of box time): the tests that touch the two kernels made default in the second
# session, through the Python API against the oracle; then the native checker with every section at 256^3 and smoke().
#   gpurun --timeout 110 -- 'bash profiles/r2/final_check.sh'
source profiles/devcheck_env.sh
mkdir -p gpurun_out
timeout 85 python -m pytest tests/test_gpu_widen.py tests/test_gpu_parity.py -m gpu -x -q \
    -k "two_stage or assemble_stage_launch or opt_in_async or full_size or xfused_launch or generic_and_fast or rk4_properties_256 or rk4_fused_assembly or fused_stage_integrators or mhd_128cubed" \
    > gpurun_out/pytest_gpu_final3.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_final3.log
tail -4 gpurun_out/pytest_gpu_final3.log
timeout 25 tests/native/_build/devcheck 256 256 gpurun_out/devcheck_final_256.txt reps=3 > /dev/null 2>&1
echo "devcheck rc=$?"; grep -c " ok$" gpurun_out/devcheck_final_256.txt; grep -i "fail" gpurun_out/devcheck_final_256.txt | head -5; tail -3 gpurun_out/devcheck_final_256.txt | cut -c1-300
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
