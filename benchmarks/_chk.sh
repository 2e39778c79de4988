python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 2>/dev/null | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --no-scale-base 2>/dev/null | cut -c1-200
