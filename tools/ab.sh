#!/bin/sh
# A/B on the same box: libmlmap_b200.so (A) against every lib/v_*.so variant, alternating runs of the timing tools
for i in 1 2; do
  echo "A         : $(python tools/frame_overheads.py | tail -1) | $(python tools/lidar_timing.py)"
  for v in mlmapping_b200/lib/v_*.so; do
    echo "$(basename $v) : $(MLM_LIB_PATH=$v python tools/frame_overheads.py | tail -1) | $(MLM_LIB_PATH=$v python tools/lidar_timing.py)"
  done
done
