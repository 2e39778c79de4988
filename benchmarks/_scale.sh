n=$1
mkdir -p gpurun_out/r02
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/r02/b200_C5_n$n.json 2> gpurun_out/r02/b200_C5_n$n.err
python -c "
import json
d=json.loads(open('gpurun_out/r02/b200_C5_n$n.json').read().strip().splitlines()[-1])
print($n, round(d['value']), round(d['ms_per_step'],1), d['stages_ms'], 'e2e', round(d['e2e']['value']) if d.get('e2e') else None, d['roofline']['launches_timed'])"
