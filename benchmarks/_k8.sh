mkdir -p gpurun_out/r02b
timeout 1100 python -m pytest tests -q -m gpu > gpurun_out/r02b/gputest_k8b.log 2>&1; tail -15 gpurun_out/r02b/gputest_k8b.log
run() { python bench.py --no-cpu-baseline --no-scale-base --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), round(d['ms_per_step'],2), d['stages_ms'], 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2))"; }
run new
run new
