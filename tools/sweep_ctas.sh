F="--no-extra-workloads --no-cpu-baseline --no-ingest-leg --no-ref-cache-leg --no-dropin-leg"
for c in 3 4 5; do for fl in 4 6; do
  BK_ASM_CTAS_PER_SM=$c python bench.py --steps 24 --inflight $fl $F 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('C2x500 W=4 ctas/SM $c inflight $fl', round(d['ms_per_step'],2),'ms/step e2e', round(d['e2e']['ms_per_step'],2), 'seq', round(d['run']['sequential_latency_ms_per_step'],2))"
done; done
