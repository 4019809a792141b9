F="--no-extra-workloads --no-cpu-baseline --no-ingest-leg --no-ref-cache-leg --no-dropin-leg"
for reg in 2500 5000 20000; do
python bench.py --workload C5 --regions $reg --inflight 8 --steps 12 $F 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('C5 $reg regions', round(d['ms_per_step'],2),'ms/step =', round(d['ms_per_step']/($reg/2500.0),2), 'ms per 2500; e2e', round(d['e2e']['ms_per_step'],2), 'seq', round(d['run']['sequential_latency_ms_per_step'],2), d['run']['host_resident'], d['run']['batches_in_flight'])"
done
