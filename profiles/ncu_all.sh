#!/bin/bash
# Run under gpurun:  bash profiles/ncu_all.sh <tag> [kernels...]      (1 GPU; ncu replays each kernel ~40x under --set full)
#   gpurun_out/launches_<tag>.csv    every launch of one RK4 step: gpu__time_duration + dram bytes read / written
#   gpurun_out/prof_<tag>_<k>.ncu-rep, gpurun_out/raw_<tag>_<k>.csv   --set full capture of ONE launch of kernel <k>
# kernels: z_inv y_inv x_fused y_fwd z_fwd assemble_stage (default: all six).  bench.py --quick times nothing under ncu.
TAG=${1:-run}; shift
WHAT=${@:-z_inv y_inv x_fused y_fwd z_fwd assemble_stage}
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off"
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --quick --steps 1 --warmup 3 > gpurun_out/launches_${TAG}.log 2>&1
# launch order inside one RHS: z_inv (strided +1), y_inv (strided +1), x_fused, y_fwd (strided -1), z_fwd (strided -1), assemble_stage
for w in $WHAT; do
  case $w in
    z_inv) K="regex:strided_(two|fast).*Li1E"; S=0;;
    y_inv) K="regex:strided_(two|fast).*Li1E"; S=1;;
    x_fused) K="regex:xfused"; S=0;;
    y_fwd) K="regex:strided_(two|fast).*Lin1E"; S=0;;
    z_fwd) K="regex:strided_(two|fast).*Lin1E"; S=1;;
    assemble_stage) K="regex:AssembleStageF"; S=0;;
  esac
  $NCU --set full --import-source on --kernel-name-base mangled -k "$K" -s $S -c 1 -f -o gpurun_out/prof_${TAG}_$w \
    python bench.py --quick --steps 1 --warmup 3 > gpurun_out/prof_${TAG}_$w.log 2>&1
  ncu -i gpurun_out/prof_${TAG}_$w.ncu-rep --page raw --csv > gpurun_out/raw_${TAG}_$w.csv 2>/dev/null
done
ls -la gpurun_out | head -40
