python tools/stage_profile.py > gpurun_out/s5_stage_profile.log 2>&1; grep -E "^(16|18|20|21|22|24) " gpurun_out/s5_stage_profile.log | sed -e 's/accumulate.*bucket_reduce/... bucket_reduce/'
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
