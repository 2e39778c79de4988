set -x
mkdir -p gpurun_out/r01
for c in C4 C1 C2 C3 C5; do
  timeout 500 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r01/b200_$c.json 2> gpurun_out/r01/b200_$c.err
  timeout 500 python bench.py --impl reference --config $c --steps 2 --warmup 1 > gpurun_out/r01/ref_$c.json 2> gpurun_out/r01/ref_$c.err
done
timeout 400 python benchmarks/kernels.py all --json gpurun_out/r01/kernels_latest.json > /dev/null 2> gpurun_out/r01/kernels.err
python benchmarks/kernels.py sample > gpurun_out/r01/kernels_sample.json 2>&1
python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 3 > gpurun_out/r01/plain.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 40000 -c 14000 --csv --log-file gpurun_out/r01/launches_bench_C4.csv python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 3 > gpurun_out/r01/ncu_launches.log 2>&1
python benchmarks/kernels.py loss_c4 > /dev/null && ncu --set full --clock-control none --import-source on -k regex:gridnet_kernel -s 3 -c 1 -o gpurun_out/r01/loss_c4 -f python benchmarks/kernels.py loss_c4 > gpurun_out/r01/ncu_c4.log 2>&1
python benchmarks/kernels.py loss_c5 > /dev/null && ncu --set full --clock-control none --import-source on -k regex:gridnet -s 6 -c 2 -o gpurun_out/r01/loss_c5 -f python benchmarks/kernels.py loss_c5 > gpurun_out/r01/ncu_c5.log 2>&1
python benchmarks/kernels.py gae_big > /dev/null && ncu --set full --clock-control none --import-source on -k regex:gae_scan -s 3 -c 1 -o gpurun_out/r01/gae_big -f python benchmarks/kernels.py gae_big > gpurun_out/r01/ncu_gae.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gae_scan -s 30 -c 1 -o gpurun_out/r01/gae_v13 -f python benchmarks/kernels.py gae_big > gpurun_out/r01/ncu_gae13.log 2>&1
ls -la gpurun_out/r01
