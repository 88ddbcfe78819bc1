timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q -k "vocab or beam or nar" > gpurun_out/r02_t2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_t2.log; tail -4 gpurun_out/r02_t2.log
CARE_B200_VOCAB_SPLIT=1 timeout 120 python scripts/vb_trace.py > gpurun_out/r02_vb_trace_2560_split.txt 2>&1; grep -E "====|all clusters" gpurun_out/r02_vb_trace_2560_split.txt
run() { # split hints batch
  CARE_B200_VOCAB_SPLIT=$1 CARE_B200_L2_HINTS=$2 python bench.py --batch $3 --steps 5 --warmup 3 --no-latency --no-e2e --no-cpu-baseline 2>/dev/null > gpurun_out/r02_bench_b$3_s$1_h$2.json
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_b$3_s$1_h$2.json')); print('split $1 hints $2 batch $3', round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"
}
run 0 0 512; run 1 0 512; run 1 2 512; run 1 3 512; run 0 0 512
run 0 0 4096; run 1 0 4096; run 1 3 4096; run 0 0 4096; run 1 0 4096
