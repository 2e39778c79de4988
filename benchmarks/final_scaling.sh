set -x
mkdir -p gpurun_out/final
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/final/b200_C4_n$n.json 2> gpurun_out/final/n$n.err
done
python bench.py --steps 5 --warmup 3 > gpurun_out/final/b200_C4.json 2> gpurun_out/final/n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final/ref_C4.json 2>/dev/null
python benchmarks/kernels.py sample > gpurun_out/final/kernels_sample.json 2>&1
for f in gpurun_out/final/b200_C4*.json; do cut -c1-160 $f; done
