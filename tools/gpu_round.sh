#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list + full capture of the top kernels.
# Usage (under gpurun): [NO_NCU=1] bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L; nproc
timeout -k 10 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $OUT/${TAG}_gputests.log
tail -5 $OUT/${TAG}_gputests.log
timeout -k 10 600 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 3000 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
if [ -z "$NO_NCU" ]; then
# launch list of the bench command itself (cold-cache, serialised: compare SHARES, not absolutes)
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu > $OUT/${TAG}_ncu_bench.log 2>&1
# full-set capture of ONE whole step (16 launches of the six dominant kernels: the two forward kernels run once
# per depth wave) on a 50-row band of the frame (replays are slow on 6 GB); 3 warm-up steps are skipped
timeout -k 10 900 ncu --set full --clock-control none --import-source on \
    -k 'regex:^(k_advect_bwd_h|k_march|k_sample_advect_h|k_density_bwd|k_app_bwd|k_appearance)$' --launch-skip 48 -c 16 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 1 --warmup 3 --no-cpu --rows 50 > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | tail -8
fi
