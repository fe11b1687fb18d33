set -x
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_r02_c.json 2> gpurun_out/bench_r02_c.err
tail -c 600 gpurun_out/bench_r02_c.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 6 --warmup 3 --no-cpu --no-lidar --no-agents > gpurun_out/bench_under_ncu_r02.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_frame --launch-skip 5 -c 1 -o gpurun_out/prof_frame_r02 -f python tools/profile_frame.py 8 > gpurun_out/prof_frame_r02.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_get_" --launch-skip 3 -c 3 -o gpurun_out/prof_queries_r02 -f python tools/profile_queries.py > gpurun_out/prof_queries_r02.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_frame --launch-skip 5 -c 1 -o gpurun_out/prof_lidar_r02 -f python tools/lidar_timing.py > gpurun_out/prof_lidar_r02.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_shard_|k_column|k_fuse" --launch-skip 30 -c 6 -o gpurun_out/prof_shard_r02 -f python tools/lidar_shard_timing.py 1 > gpurun_out/prof_shard_r02.log 2>&1
ls -la gpurun_out/*r02*
