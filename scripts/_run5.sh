timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_n.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_gputest_n.log; tail -4 gpurun_out/r02_gputest_n.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_b.log 2>&1; echo "smoke exit $?" >> gpurun_out/r02_smoke_b.log; tail -3 gpurun_out/r02_smoke_b.log
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_final.json')); print('bench', round(d['value']), d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('single_call_value'), 'cpu', d['cpu_baseline'])"
bash scripts/profile_r02.sh v3 4096 full 2>&1 | tail -30
bash scripts/profile_r02.sh v4 512 2>&1 | tail -16
