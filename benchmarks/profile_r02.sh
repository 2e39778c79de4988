# Round-2 measurement pass (run on a GPU box: gpurun -- 'sh benchmarks/profile_r02.sh').  Everything lands in
# gpurun_out/r02/; the summaries that are kept go to profiles/r02/ (benchmarks/ncu_summary.py, launch_summary.py).
set -x
O=gpurun_out/r02
mkdir -p $O
timeout 900 python benchmarks/kernels.py all --json $O/kernels.json > $O/kernels.log 2> $O/kernels.err
python benchmarks/kernels.py sample > $O/kernels_sample.json 2>&1
# ncu: full capture of the learner's fused-loss launch at the C4 minibatch and at one C5 per-GPU minibatch
python benchmarks/kernels.py loss_c4_inplace > /dev/null && ncu --set full --clock-control none --import-source on -k regex:gridnet_kernel -s 5 -c 1 -o $O/loss_c4_inplace -f python benchmarks/kernels.py loss_c4_inplace > $O/ncu_c4.log 2>&1
python benchmarks/kernels.py loss_c5_inplace > /dev/null && ncu --set full --clock-control none --import-source on -k regex:gridnet -s 10 -c 2 -o $O/loss_c5_inplace -f python benchmarks/kernels.py loss_c5_inplace > $O/ncu_c5.log 2>&1
# K8 at the C4 minibatch, encoder level 1
python benchmarks/kernels.py glue > /dev/null && ncu --set full --clock-control none --import-source on -k regex:'pool_fwd|pool_bwd|colsum|bias_relu|relu_bwd' -s 12 -c 7 -o $O/glue_c4 -f python benchmarks/kernels.py glue > $O/ncu_glue.log 2>&1
# launch list of a bench step (its own command first, without ncu)
python bench.py --no-cpu-baseline --no-e2e --no-scale-base --steps 1 --warmup 3 > $O/plain.json 2> $O/plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 12000 --csv --log-file $O/launches_bench_C4.csv python bench.py --no-cpu-baseline --no-e2e --no-scale-base --steps 1 --warmup 3 > $O/ncu_launches.log 2>&1
ls -la $O
