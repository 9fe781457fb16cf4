python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python tools/one_msm.py 24
python tools/stage_profile.py > gpurun_out/s4_stage_profile_b.log 2>&1; grep -E "^(16|21|24) " gpurun_out/s4_stage_profile_b.log
KF="regex:k_(decompose_b|scatter_staged_b|bucket_hist_b|bucket_scatter_staged_b|accumulate|reduce_items|bucket_reduce|combine_single)"
ncu --set full --clock-control none --import-source on -k "$KF" --launch-skip 10 -c 10 -f -o gpurun_out/r02f_prof python tools/one_msm.py 24 > gpurun_out/r02f_ncu_full.log 2>&1
tail -3 gpurun_out/r02f_ncu_full.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --prove-k 0 --no-skew > gpurun_out/r02f_bench_short.json 2> gpurun_out/r02f_bench_short.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --prove-k 0 --no-skew > gpurun_out/r02f_ncu_list.log 2>&1
tail -2 gpurun_out/r02f_ncu_list.log | cut -c1-300
ls -la gpurun_out/r02f_*
