F="--no-extra-workloads --no-cpu-baseline --no-ingest-leg --no-ref-cache-leg --no-dropin-leg"
for w in 1 2 4 8; do
  python bench.py --steps 6 --regions 3000 --inflight 1 --spec-width $w $F 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('packed3000 W=$w', round(d['ms_per_step'],2),'ms/step', d['kernel_ms_per_step']['assemble'])"
done
for w in 2 4 8; do
  python bench.py --steps 24 --inflight 4 --spec-width $w $F 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('C2x500 inflight4 W=$w', round(d['ms_per_step'],2),'ms/step seq', d['run']['sequential_latency_ms_per_step'])"
done
for fl in 2 6 8; do
  python bench.py --steps 24 --inflight $fl $F 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('C2x500 W=4 inflight $fl', round(d['ms_per_step'],2),'ms/step e2e', round(d['e2e']['ms_per_step'],2))"
done
