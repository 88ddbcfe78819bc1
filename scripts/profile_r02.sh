# Round-2 profiles (run on the GPU box: gpurun -- bash scripts/profile_r02.sh TAG [BATCH]).
# 1) unprofiled bench run that records the GEMM variants it picked; 2) ncu launch list replaying those picks;
# 3) (BATCH = 4096 only) `ncu --set full` capture of one decode step.
TAG=${1:-v0}; B=${2:-4096}
mkdir -p gpurun_out
export CARE_B200_GEMM_CHOICE_FILE=$PWD/gpurun_out/gemm_choices_b${B}.txt
rm -f $CARE_B200_GEMM_CHOICE_FILE
python bench.py --batch $B --steps 5 --warmup 3 --no-latency --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_b${B}_${TAG}.json 2>/dev/null
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_b${B}_${TAG}.csv \
  python bench.py --batch $B --steps 2 --warmup 3 --no-latency --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python scripts/launch_summary.py gpurun_out/r02_launches_b${B}_${TAG}.csv > gpurun_out/r02_launches_b${B}_${TAG}_summary.txt
head -24 gpurun_out/r02_launches_b${B}_${TAG}_summary.txt
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_b${B}_${TAG}.json')); print('bench', d['value'], d['ms_per_step'])"
if [ "$B" = "4096" ] && [ "$3" = "full" ]; then
  timeout 600 ncu --set full --clock-control none -s 620 -c 28 -o /tmp/step_${TAG} -f python bench.py --batch $B --steps 2 --warmup 1 --no-latency --no-e2e --no-cpu-baseline > /dev/null 2>&1
  python scripts/step_profile.py /tmp/step_${TAG}.ncu-rep gpurun_out/r02_ncu_full_${TAG}_step.txt gpurun_out/r02_traffic_${TAG}.json "ncu --set full --clock-control none -s 620 -c 28 on bench.py --batch 4096 (GEMM variants replayed from the unprofiled run): the first complete decode step inside the window"
fi
