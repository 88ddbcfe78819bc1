timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "beam or vocab or attention or gemm" > gpurun_out/r02_t1.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_t1.log; tail -4 gpurun_out/r02_t1.log
timeout 120 python scripts/vb_trace.py > gpurun_out/r02_vb_trace_2560.txt 2>&1; tail -5 gpurun_out/r02_vb_trace_2560.txt
for h in 0 3 1; do
  for b in 512 4096; do
    CARE_B200_L2_HINTS=$h python bench.py --batch $b --steps 5 --warmup 3 --no-latency --no-e2e --no-cpu-baseline 2>/dev/null > gpurun_out/r02_bench_b${b}_hint${h}.json
    python -c "
import json; d=json.load(open('gpurun_out/r02_bench_b${b}_hint${h}.json')); print('hints $h batch $b', round(d['value']), d['ms_per_step'], d['clocks']['sm_mhz'])"
  done
done
