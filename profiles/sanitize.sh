#!/bin/bash
# Device-side sanitizers (run under gpurun; 2 GPUs for the slab part):  bash profiles/sanitize.sh
# compute-sanitizer on the native checker with the kernels that synchronise below CTA level or through mbarriers / cp.async:
#   xfused_kernel (warp-private shared-memory stages, __syncwarp only), xfused_persist_kernel (mbarrier + bulk copies),
#   strided_staged (cp.async staging), and -- with 2 GPUs -- one slab-decomposed RK4 step with the peer-store and push exchanges
# (flag protocol without buffer-free credits, csrc/p2p.cu).  Logs: gpurun_out/sanitizer_*.log
source profiles/devcheck_env.sh
mkdir -p gpurun_out
F="--kernel-regex kns=xfused"
for tool in racecheck memcheck; do
  timeout 600 compute-sanitizer --tool $tool $F --error-exitcode 9 tests/native/_build/devcheck 512 0 /dev/null > gpurun_out/sanitizer_${tool}_xfused512.log 2>&1
  echo "$tool xfused 512: exit $?"; tail -3 gpurun_out/sanitizer_${tool}_xfused512.log
done
timeout 600 compute-sanitizer --tool racecheck --kernel-regex kns=strided_staged --error-exitcode 9 tests/native/_build/devcheck 512 0 /dev/null strided_staged=1 > gpurun_out/sanitizer_racecheck_staged512.log 2>&1
echo "racecheck strided_staged 512: exit $?"; tail -3 gpurun_out/sanitizer_racecheck_staged512.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 tests/native/_build/devcheck 128 0 /dev/null > gpurun_out/sanitizer_racecheck_all128.log 2>&1
echo "racecheck all kernels 128: exit $?"; tail -3 gpurun_out/sanitizer_racecheck_all128.log
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  for ex in peer push; do
    DEDALUS_KY_LAYOUT=cyclic DEDALUS_SLAB_EXCHANGE=$ex timeout 420 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 \
      python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2957$((RANDOM % 10)) \
      tests/slab_asym_worker.py gpurun_out/sanitize_slab_$ex.json > gpurun_out/sanitizer_memcheck_slab2_$ex.log 2>&1
    echo "memcheck 2-GPU slab ($ex): exit $?"; grep -E "ERROR SUMMARY|rel" gpurun_out/sanitizer_memcheck_slab2_$ex.log | tail -4
  done
fi
