#!/bin/bash
# First GPU call of round 2 (run under gpurun, ~10 minutes):  gpurun --timeout 1500 -- 'bash profiles/round2_first_call.sh'
# Everything built after the last full GPU test run of round 1 gets its first device run here; outputs land in gpurun_out/.
#   1. the native checker (3 s): device-only parts, one-to-one kernels, shearing-box transform, timing sweep incl. rhs_plane_chunk
#   2. pytest -m gpu (ordered: what has passed on a B200 first), without -x so that one surprise does not hide the rest
#   3. bench.py (headline line), then the widened-config timings and the 2-D eager / graph step times
set -u
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$(python - <<'PY'
import os
try:
    import nvidia.cuda_runtime as m
    print(os.path.join(list(m.__path__)[0], "lib"))
except Exception:
    print("/usr/local/cuda/lib64")
PY
):/usr/local/cuda/lib64:${LD_LIBRARY_PATH:-}
tests/native/_build/devcheck 128 512 gpurun_out/devcheck_r2.txt reps=5 > gpurun_out/devcheck_r2.log 2>&1
tail -25 gpurun_out/devcheck_r2.txt
# device-side sanitizers on the native checker (the CPU suite runs the same kernel bodies under ASan/UBSan, tests/test_host_sanitizer.py;
# shared-memory hazards of the warp-private x-pass stages and the shuffle epilogues exist only here)
for tool in memcheck racecheck; do
  timeout 240 compute-sanitizer --tool $tool --error-exitcode 9 tests/native/_build/devcheck 64 > gpurun_out/sanitizer_${tool}_r2.log 2>&1
  echo "compute-sanitizer $tool: exit $?"; tail -4 gpurun_out/sanitizer_${tool}_r2.log
done
python -m pytest tests -q -m gpu -p no:cacheprovider --durations=15 > gpurun_out/pytest_gpu_r2.log 2>&1
tail -30 gpurun_out/pytest_gpu_r2.log
python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
tail -c 1500 gpurun_out/bench_r2.json
# end-to-end leg with the full arrays, for comparison with the retained-box transfers the default run uses
python bench.py --e2e-full --steps 4 --warmup 3 > gpurun_out/bench_r2_e2e_full.json 2> gpurun_out/bench_r2_e2e_full.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_r2.json", "gpurun_out/bench_r2_e2e_full.json"):
    try:
        e = json.loads(open(f).read().strip().splitlines()[-1])["e2e"]
        print(f, "e2e ms/step %.1f" % e["ms_per_step"], e["transfer"][:40], e["h2d_bytes_per_step"])
    except Exception as x:
        print(f, "unreadable:", x)
PY
python profiles/widened_configs.py > gpurun_out/widened_r2.jsonl 2>&1
cat gpurun_out/widened_r2.jsonl
python profiles/small_grids.py > gpurun_out/small_grids_r2.log 2>&1
tail -8 gpurun_out/small_grids_r2.log
