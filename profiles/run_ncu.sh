#!/bin/bash
# Run under gpurun.  $1 = tag (e.g. r1_v2).  Produces, in gpurun_out/:
#   launches_<tag>.csv   every launch of one RK4 step with gpu__time_duration (cold cache, serialised)
#   prof_<tag>_<name>.ncu-rep   --set full captures of the kernels named in $2.. (default: fused)
TAG=${1:-run}; shift
WHAT=${@:-fused}
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off"
$NCU --metrics gpu__time_duration.sum -c 80 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --quick --steps 1 --warmup 2 > gpurun_out/launches_${TAG}.log 2>&1
for w in $WHAT; do
  case $w in
    fused) K="regex:xfused_kernel|Li512ELi3"; S=1;;
    ypass) K="regex:strided_fast"; S=7;;
    zpass) K="regex:strided_fast"; S=1;;
    stage) K="regex:StageF"; S=1;;
    astage) K="regex:AssembleStageF"; S=1;;
    assemble) K="regex:AssembleF"; S=1;;
  esac
  $NCU --set full --import-source on --kernel-name-base mangled -k "$K" -s $S -c 1 -f -o gpurun_out/prof_${TAG}_$w \
    python bench.py --quick --steps 1 --warmup 2 > gpurun_out/prof_${TAG}_$w.log 2>&1
done
ls -la gpurun_out
