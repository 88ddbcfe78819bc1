timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q -k "vocab or beam" > gpurun_out/r02_t3.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_t3.log; tail -3 gpurun_out/r02_t3.log
run() { # split hints batch
  CARE_B200_VOCAB_SPLIT=$1 CARE_B200_L2_HINTS=$2 python bench.py --batch $3 --steps 5 --warmup 3 --no-latency --no-e2e --no-cpu-baseline 2>/dev/null > gpurun_out/r02_bench_b$3_s$1_h$2.json
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_b$3_s$1_h$2.json')); print('split $1 hints $2 batch $3', round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], [ (r['kernel'][:12], round(r['avg_launch_ms']*1e3,1)) for r in [d['roofline']]+d['roofline_other_kernels'] if 'vocab' in r['kernel']])"
}
run 0 2 512; run 1 2 512; run 0 2 4096; run 1 2 4096; run 0 0 4096
