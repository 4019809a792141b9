F="--steps 24 --no-extra-workloads --no-cpu-baseline --no-ingest-leg --no-ref-cache-leg --no-dropin-leg"
for fl in 4 6 8 12; do
python bench.py --regions 1000 --inflight $fl $F 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('1000 regions inflight $fl', round(d['ms_per_step'],2),'ms/step e2e', round(d['e2e']['ms_per_step'],2), 'seq', round(d['run']['sequential_latency_ms_per_step'],2))"
done
for fl in 4 8; do
python bench.py --inflight $fl $F 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('500 regions inflight $fl', round(d['ms_per_step'],2),'ms/step e2e', round(d['e2e']['ms_per_step'],2))"
done
