# `ncu --set full` capture (with source correlation) of one decode step of a 512-video shard (strong-scaling regime).
# Run on the GPU box: gpurun -- bash scripts/profile_b512_full.sh TAG
TAG=${1:-v0}
mkdir -p gpurun_out
export CARE_B200_GEMM_CHOICE_FILE=$PWD/gpurun_out/gemm_choices_b512_${TAG}.txt
rm -f $CARE_B200_GEMM_CHOICE_FILE
python bench.py --batch 512 --steps 5 --warmup 3 --no-latency --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_b512_${TAG}.json 2>/dev/null
timeout 600 ncu --set full --import-source on --clock-control none -s 620 -c 28 -o gpurun_out/step_b512_${TAG} -f python bench.py --batch 512 --steps 2 --warmup 1 --no-latency --no-e2e --no-cpu-baseline > /dev/null 2>&1
python scripts/step_profile.py gpurun_out/step_b512_${TAG}.ncu-rep gpurun_out/r02_ncu_full_b512_${TAG}_step.txt gpurun_out/r02_traffic_b512_${TAG}.json "ncu --set full --clock-control none -s 620 -c 28 on bench.py --batch 512 (GEMM variants replayed from the unprofiled run): the first complete decode step inside the window"
cat gpurun_out/r02_ncu_full_b512_${TAG}_step.txt
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_b512_${TAG}.json')); print('bench', d['value'], d['ms_per_step'])"
ls -la gpurun_out/step_b512_${TAG}.ncu-rep
