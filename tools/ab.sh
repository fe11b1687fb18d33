#!/bin/sh
# A/B on the same box: libmlmap_b200.so (A) vs lib/variant.so (B), alternating runs of the timing tools
for i in 1 2; do
  echo "A: $(python tools/frame_overheads.py | tail -1) | $(python tools/lidar_timing.py)"
  echo "B: $(MLM_LIB_PATH=mlmapping_b200/lib/variant.so python tools/frame_overheads.py | tail -1) | $(MLM_LIB_PATH=mlmapping_b200/lib/variant.so python tools/lidar_timing.py)"
done
