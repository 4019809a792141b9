# N=2: the host-side work queue over the ranks vs the static partition, with the sharding check (small legs off)
F="--gpus 2 --steps 24 --warmup 3 --no-dropin-leg --no-cpu-baseline"
export BK_BENCH_TEST_STARVE_RANK=1   # rank 1 runs nothing of the last pass: the line must still come out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
S='import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); json.dump(d, open("gpurun_out/n2_%s.json" % sys.argv[1], "w")); print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["run"]["device_ms_per_step_by_rank"], d["run"]["calls_run_by_rank"], d["sharding_check"]["equal_to_single_gpu_run"], {k: (round(v["value"]), round(v["e2e"]["value"]), v["equal_to_single_gpu_run"], v["calls_run_by_rank"], v["device_ms_per_pass_by_rank"]) for k, v in d.items() if k in ("c5_strong", "c3_sharded") and v})'
mkdir -p gpurun_out
$T --master-port 29701 bench.py $F 2>gpurun_out/n2_dynamic.err | python -c "$S" dynamic
tail -5 gpurun_out/n2_dynamic.err
