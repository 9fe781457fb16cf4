python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py > gpurun_out/r02_bench_line_j.json 2> gpurun_out/r02_bench_line_j.err ) 2> gpurun_out/r02_bench_line_j.time
tail -3 gpurun_out/r02_bench_line_j.time; tail -c 300 gpurun_out/r02_bench_line_j.err
( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_line_j.json 2> gpurun_out/r02_bench_reference_j.err ) 2> gpurun_out/r02_bench_reference_j.time
tail -3 gpurun_out/r02_bench_reference_j.time; cut -c1-600 gpurun_out/r02_bench_reference_line_j.json
