# End-of-round measurement pass: kernels, ncu captures, launch list (profile_r02.sh), then the bench lines.
O=gpurun_out/r02
mkdir -p $O
sh benchmarks/profile_r02.sh > gpurun_out/profile_r02.log 2>&1
python benchmarks/trunk_profile.py C5 > $O/trunk_profile_C5_fused.txt 2>&1
B200RL_FUSED_GLUE=0 python benchmarks/trunk_profile.py C5 > $O/trunk_profile_C5_torch.txt 2>&1
python benchmarks/e2e_breakdown.py C4 events > $O/e2e_timeline_C4.txt 2>&1
python bench.py --steps 20 --warmup 5 > $O/b200_C4.json 2> $O/b200_C4.err
tail -c 600 $O/b200_C4.json
