( time python bench.py > gpurun_out/s4_bench_1gpu.json 2> gpurun_out/s4_bench_1gpu.err ) 2> gpurun_out/s4_bench_1gpu.time
tail -3 gpurun_out/s4_bench_1gpu.time; tail -c 400 gpurun_out/s4_bench_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s4_bench_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e'].get('staging_rate_gbps_rank0'), 'pinned', d['e2e_pinned']['ms_per_step'], 'plain', d.get('plain_bases'))
print('roofline', {k:d['roofline'][k] for k in ('frac','executed_frac','frac_of_madd_stream','kernel_ms')})
print('prove', d['hyperplonk_prove']['k24']['gpu_ms'], d['hyperplonk_prove']['k20']['gpu_ms'], d['hyperplonk_prove']['k20'].get('cpu_ms'))
print('msmseq', d['hyperplonk_prove_msm']['k24']['gpu_ms'], d['hyperplonk_prove_msm']['k24']['gpu_resident_ms'])
PY
