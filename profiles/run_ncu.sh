#!/bin/bash
# Run under gpurun.  $1 = tag (e.g. r1_v1).  Produces, in gpurun_out/:
#   launches_<tag>.csv   every launch of one RK4 step with gpu__time_duration (cold cache, serialised)
#   prof_<tag>_{fused,ypass,stage}.ncu-rep   --set full captures of the dominant kernels
TAG=${1:-run}
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off"
$NCU --metrics gpu__time_duration.sum -c 80 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --quick --steps 1 --warmup 1 > gpurun_out/launches_${TAG}.log 2>&1
$NCU --set full --import-source on --kernel-name-base mangled -k regex:Li512ELi3 -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_fused \
    python bench.py --quick --steps 1 --warmup 1 > gpurun_out/prof_${TAG}_fused.log 2>&1
$NCU --set full --import-source on --kernel-name-base mangled -k regex:Li512ELi0ELin1 -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_ypass \
    python bench.py --quick --steps 1 --warmup 1 > gpurun_out/prof_${TAG}_ypass.log 2>&1
$NCU --set full --import-source on --kernel-name-base mangled -k regex:StageF -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_stage \
    python bench.py --quick --steps 1 --warmup 1 > gpurun_out/prof_${TAG}_stage.log 2>&1
ls -la gpurun_out
