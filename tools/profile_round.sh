# one round of profiling evidence (gpurun brings back at most 64 MiB of gpurun_out/: run PART=1 and PART=2 as separate calls)
set -x
PART=${PART:-1}
TAG=${TAG:-r02}
if [ "$PART" = 1 ]; then
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_${TAG}_c.json 2> gpurun_out/bench_${TAG}_c.err
tail -c 600 gpurun_out/bench_${TAG}_c.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 6 --warmup 3 --no-cpu --no-lidar --no-agents > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_frame --launch-skip 5 -c 1 -o gpurun_out/prof_frame_${TAG} -f python tools/profile_frame.py 8 > gpurun_out/prof_frame_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_get_" --launch-skip 3 -c 3 -o gpurun_out/prof_queries_${TAG} -f python tools/profile_queries.py > gpurun_out/prof_queries_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_frame_explore --launch-skip 14 -c 1 -o gpurun_out/prof_explore_${TAG} -f python tools/explore_timing.py > gpurun_out/prof_explore_${TAG}.log 2>&1
else
ncu --set full --clock-control none --import-source on -k regex:k_frame --launch-skip 5 -c 1 -o gpurun_out/prof_lidar_${TAG} -f python tools/lidar_timing.py > gpurun_out/prof_lidar_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_shard_|k_column|k_fuse" --launch-skip 30 -c 6 -o gpurun_out/prof_shard_${TAG} -f python tools/lidar_shard_timing.py 1 > gpurun_out/prof_shard_${TAG}.log 2>&1
fi
ls -la gpurun_out/*${TAG}*
du -sh gpurun_out
