set -x
for m in f16x3; do NVFI_MLP_MODE=$m timeout 200 python __graft_entry__.py smoke 2>&1 | grep -E "train step|PDE"; done
timeout -k 5 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout -k 5 300 python tools/diag_headline_grad.py 2>&1 | tail -30
timeout -k 5 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err
python - <<P
import json
j=json.loads(open("gpurun_out/r02m_bench.json").read().strip().splitlines()[-1])
print(j["value"], j["ms_per_step"])
for k,v in j["kernels"].items():
    if v["share"]>0.005: print(k, round(v["ms_per_launch"],2))
P
