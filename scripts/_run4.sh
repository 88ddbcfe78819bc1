run() { # tag batch  (env passed by caller)
  python bench.py --batch $2 --steps 5 --warmup 3 --no-latency --no-e2e --no-cpu-baseline 2>/dev/null > gpurun_out/r02_bench_b$2_$1.json
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_b$2_$1.json')); print('$1 batch $2', round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"
}
run base 512
CARE_B200_FUSE_INFO=1 run finfo 512
CARE_B200_FUSE_NEXT=1 run fnext 512
CARE_B200_FUSE_INFO=1 CARE_B200_FUSE_NEXT=1 run fboth 512
run base 4096
CARE_B200_FUSE_INFO=1 run finfo 4096
CARE_B200_FUSE_INFO=1 CARE_B200_FUSE_NEXT=1 run fboth 4096
timeout 300 python scripts/e2e_probe.py > gpurun_out/r02_e2e_probe.txt 2>&1; tail -30 gpurun_out/r02_e2e_probe.txt
