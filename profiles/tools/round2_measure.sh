#!/bin/bash
# Round-2 first GPU call: everything that was written in round 1 after the GPU budget ran out, measured in one go.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash profiles/tools/round2_measure.sh A'     (parity gate + default + launch shapes, ~35 min)
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash profiles/tools/round2_measure.sh B'     (L = 64 line + ncu captures, ~35 min)
# (no argument = everything, ~80 min of box time: more than one call should carry)
# Outputs land in gpurun_out/r2_*; copy what is to be judged into profiles/ afterwards.
# Every step runs under its own `timeout`, so a hang in one (new, GPU-untested) kernel cannot take the slot.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
STAGE="${1:-AB}"
step() { echo "=== $1" | tee -a $O/r2_steps.log; shift; "$@"; echo "rc=$?" | tee -a $O/r2_steps.log; }
if [[ "$STAGE" == *A* ]]; then

# 1. parity gate: the old GPU tests, then the new ones (beta doubling, 2/4 walkers per warp)
step "pytest gpu (parity, no statistics)" timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_statistics.py \
     > $O/r2_pytest_parity.log 2>&1
# 2. default bench line (BASELINE configs[1], one walker per warp)
step "bench default" timeout 600 python bench.py > $O/r2_bench_default.json 2> $O/r2_bench_default.err
# 3. more walkers per GPU: two waves of the default kernel vs 2 / 4 walkers per warp (all thermalised by beta doubling:
#    the cold start of 8192+ walkers alone would take minutes per run); the first line repeats the default workload with the
#    doubling setup so that its timed state can be compared with step 2
step "bench default workload, doubling setup" timeout 600 python bench.py --beta-doublings 5 --therm 100 --no-cpu --steps 3 > $O/r2_bench_default_doubling.json 2> $O/r2_bench_default_doubling.err
step "bench 8192 walkers, 1 per warp" timeout 900 python bench.py --walkers 8192 --beta-doublings 5 --therm 100 --no-cpu --steps 3 > $O/r2_bench_w8192_k1.json 2> $O/r2_bench_w8192_k1.err
step "bench 8192 walkers, 2 per warp" timeout 900 python bench.py --walkers 8192 --walkers-per-warp 2 --beta-doublings 5 --therm 100 --no-cpu --steps 3 > $O/r2_bench_w8192_k2.json 2> $O/r2_bench_w8192_k2.err
step "bench 11840 walkers, 4 per warp" timeout 900 python bench.py --walkers 11840 --walkers-per-warp 4 --beta-doublings 5 --therm 100 --no-cpu --steps 3 > $O/r2_bench_w11840_k4.json 2> $O/r2_bench_w11840_k4.err
step "bench 4096 walkers, 2 per warp (expected: no gain)" timeout 600 python bench.py --walkers-per-warp 2 --beta-doublings 5 --therm 100 --no-cpu --steps 3 > $O/r2_bench_w4096_k2.json 2> $O/r2_bench_w4096_k2.err
# 3a. other compiled occupancies of the interleaved kernel (SSE_B200_MULTI_MINB)
step "bench 8192 walkers, 2 per warp, 5 CTAs/SM" env SSE_B200_MULTI_MINB=5 timeout 900 python bench.py --walkers 8192 --walkers-per-warp 2 --beta-doublings 5 --therm 100 --no-cpu --steps 3 > $O/r2_bench_w8192_k2_b5.json 2> $O/r2_bench_w8192_k2_b5.err
step "bench 9472 walkers, 4 per warp, 4 CTAs/SM" env SSE_B200_MULTI_MINB=4 timeout 900 python bench.py --walkers 9472 --walkers-per-warp 4 --beta-doublings 5 --therm 100 --no-cpu --steps 3 > $O/r2_bench_w9472_k4_b4.json 2> $O/r2_bench_w9472_k4_b4.err
# 3b. A/B matrix on one thermalised batch (no torch): shapes x {default, no prefetch, no hint pass}
step "quick A/B 8192 walkers" timeout 300 python profiles/tools/quick_multi.py 32 32 8192 5 10 60 8 1,2,4 $O/r2_quick_ab.jsonl 0,2,4 > $O/r2_quick_ab.log 2>&1
fi
if [[ "$STAGE" == *B* ]]; then
# 4. BASELINE configs[2]: L = 64, beta = 64, grown by beta doubling
step "bench L=64 beta=64" timeout 1500 python bench.py --L 64 --beta 64 --walkers 3552 --beta-doublings 6 --therm-per-level 10 --therm 60 \
     --sweeps-per-step 8 --steps 3 --warmup 3 --cpu-therm 60 --cpu-sweeps 60 > $O/r2_bench_L64.json 2> $O/r2_bench_L64.err
# 5. ncu (skip counts: 1 init launch + 300/50 thermalisation launches + 3 warm-up launches precede the timed one)
#    (k_walkers_multi capture: 5 doubling-level launches + 100/50 thermalisation launches + 3 warm-up launches precede it)
#    launch list of the default bench command, one full capture of each kernel in the thermalised state
step "ncu launch list" timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r2_launches.csv \
     python bench.py --steps 2 --warmup 3 --therm 100 --sweeps-per-step 4 --no-cpu > $O/r2_ncu_list.log 2>&1
step "ncu full k_walkers" timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_walkers -s 10 -c 1 -f -o $O/r2_full_k1 \
     python bench.py --steps 1 --warmup 3 --therm 300 --sweeps-per-step 2 --no-cpu > $O/r2_ncu_full_k1.log 2>&1
step "ncu full k_walkers_multi" timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_walkers_multi -s 10 -c 1 -f -o $O/r2_full_k2 \
     python bench.py --walkers 8192 --walkers-per-warp 2 --steps 1 --warmup 3 --beta-doublings 5 --therm 100 --sweeps-per-step 2 --no-cpu > $O/r2_ncu_full_k2.log 2>&1
fi
tail -n 3 $O/r2_pytest_parity.log 2>/dev/null
cat $O/r2_steps.log
