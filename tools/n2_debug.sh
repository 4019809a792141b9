F="--gpus 2 --steps 48 --warmup 3 --no-ingest-leg --no-ref-cache-leg --no-dropin-leg --no-extra-workloads"
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
S='import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); print(sys.argv[1], round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), d["run"]["device_ms_per_step_by_rank"], d["clocks"])'
$T --master-port 29701 bench.py $F 2>/dev/null | python -c "$S" default
sleep 2
$T --master-port 29712 bench.py $F --replicate 2>/dev/null | python -c "$S" replicate
python bench.py --steps 48 --no-ingest-leg --no-ref-cache-leg --no-dropin-leg --no-extra-workloads --no-cpu-baseline 2>/dev/null | python -c "$S" single
