export MLM_SHARD_TIMEOUT_MS=5000
for N in ${NS:-4 2}; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N tests/multi_gpu/sharded_check.py 2>&1 | grep sharded_check | cut -c1-300
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --lidar-only > gpurun_out/lidar_n$N.json 2> gpurun_out/lidar_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/lidar_n$N.json").read().strip().splitlines()[-1])["lidar"]
    print($N, d["parity"], round(d["us_per_scan"],1), "us/scan;", d["sharded"]["stage_us_this_rank"], "e2e", d["sharded"]["e2e"])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/lidar_n$N.err").read()[-1500:])
PY
done
