python tools/stage_profile.py > gpurun_out/s4_stage_profile.log 2>&1; tail -12 gpurun_out/s4_stage_profile.log
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/open_timing.py 24 default; python tools/open_timing.py 20 default
