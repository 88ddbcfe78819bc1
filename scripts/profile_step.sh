# Launch list + `ncu --set full` capture of one decode step of bench.py (run on the GPU box: gpurun -- bash scripts/profile_step.sh)
set -x
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -q -m gpu -k "self_attention or teacher_forced" -x 2>&1 | tail -3
export CARE_B200_GEMM_2SM=1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_v6.csv python bench.py --steps 2 --warmup 1 --no-latency --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_v6.log 2>&1
python scripts/launch_summary.py gpurun_out/r01_launches_v6.csv > gpurun_out/r01_launches_v6_summary.txt; head -20 gpurun_out/r01_launches_v6_summary.txt
timeout 500 ncu --set full --clock-control none -s 660 -c 30 -o /tmp/step_v6 -f python bench.py --steps 2 --warmup 1 --no-latency --no-e2e --no-cpu-baseline > /dev/null 2>&1
python scripts/step_profile.py /tmp/step_v6.ncu-rep gpurun_out/r01_ncu_full_v6_step.txt gpurun_out/r01_traffic_v6.json "ncu --set full --clock-control none -s 660 -c 30 (CARE_B200_GEMM_2SM=1) on bench.py --batch 4096: the first complete decode step inside the window (t~15), round-1 kernels with the live-slot stream self-attention"
