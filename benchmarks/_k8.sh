mkdir -p gpurun_out/r02b
timeout 1100 python -m pytest tests -q -m gpu > gpurun_out/r02b/gputest_e2e.log 2>&1; tail -15 gpurun_out/r02b/gputest_e2e.log
python benchmarks/e2e_breakdown.py C4 events 2>&1 | tail -2
run() { python bench.py --no-cpu-baseline --no-scale-base --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), round(d['ms_per_step'],2), d['stages_ms'], 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2), d['e2e']['h2d_bytes_per_step'])"; }
run new
