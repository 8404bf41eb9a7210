#!/bin/bash
# ncu --set full captures of the frame-side kernels at the bench batch (240 frames): bash tools/ncu_frames.sh <tag>
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
for k in quadtree_orient blur_kernel match_frame fast_cells pyr_resize describe_kernel; do
  skip=1; [ $k = pyr_resize ] && skip=7
  if [ $k = match_frame ]; then
    ncu --set full --import-source on --clock-control none -k regex:$k -s 1 -c 1 -o $out/${tag}_ncu_$k python bench.py --steps 1 --warmup 1 --rounds 3 --no-cpu-baseline --no-e2e --no-extra --dev-groups 1 > $out/${tag}_ncu_$k.log 2>&1
  else
    ncu --set full --import-source on --clock-control none -k regex:$k -s $skip -c 1 -o $out/${tag}_ncu_$k python tools/stage_times.py > $out/${tag}_ncu_$k.log 2>&1
  fi
  echo "$k done"
done
