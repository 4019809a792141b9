F="--no-extra-workloads --no-cpu-baseline --no-ingest-leg --no-ref-cache-leg --no-dropin-leg"
for w in 4 8; do
python bench.py --workload C4 --steps 12 --spec-width $w $F 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('C4 W=$w', round(d['ms_per_step'],2),'ms/step e2e', round(d['e2e']['ms_per_step'],2), 'seq', round(d['run']['sequential_latency_ms_per_step'],2), d['kernel_ms_per_step'])"
done
python bench.py --workload C5 --steps 4 $F 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('C5', round(d['ms_per_step'],2),'ms/step e2e', round(d['e2e']['ms_per_step'],2), 'seq', round(d['run']['sequential_latency_ms_per_step'],2), d['kernel_ms_per_step'])"
