mkdir -p gpurun_out/r02b
cp rl_algo_impls_b200/libb200rl.so /tmp/new.so
run() { python bench.py --no-cpu-baseline --no-scale-base --no-e2e --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), round(d['ms_per_step'],2), d['stages_ms'])"; }
run new
cp build/old/libb200rl_oldk5.so rl_algo_impls_b200/libb200rl.so; run old
cp /tmp/new.so rl_algo_impls_b200/libb200rl.so; run new
cp build/old/libb200rl_oldk5.so rl_algo_impls_b200/libb200rl.so; run old
cp /tmp/new.so rl_algo_impls_b200/libb200rl.so
