set -x
timeout -k 5 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02t_bench_2gpu.json 2> gpurun_out/r02t_bench_2gpu.err; tail -c 300 gpurun_out/r02t_bench_2gpu.err
python - <<P
import json
j=json.loads(open("gpurun_out/r02t_bench_2gpu.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","no_early_termination","strong_breakdown","weak","eval_frame","chessboard_eval_t1.0","fan_mask_render"): print(k, j.get(k))
P
