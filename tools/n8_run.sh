# the driver's N=8 command; the JSON line goes to gpurun_out/r2_scale_n8.json
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 $BK_N8_FLAGS \
  2>gpurun_out/n8.err | python -c '
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])
json.dump(d, open("gpurun_out/r2_scale_n8.json", "w"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["run"]["device_ms_per_step_by_rank"], d["run"]["calls_run_by_rank"], d["sharding_check"]["equal_to_single_gpu_run"], d["clocks"])
for k in ("c5_strong", "c3_sharded"):
    v = d.get(k)
    if v: print(k, round(v["value"]), round(v["e2e"]["value"]), v["equal_to_single_gpu_run"], v["calls_run_by_rank"], v["device_ms_per_pass_by_rank"], v["result_digest"][:12])
print("from_files", d.get("from_files") and round(d["from_files"]["value"]), "dropin", d.get("e2e_dropin") and round(d["e2e_dropin"]["value"]))
'
tail -3 gpurun_out/n8.err
