# One gpurun --gpus 8 call (charged 8x: keep it short): parity tests for worlds 2/4/8, the exchange probe, the tile-sharded bench at
# C3 over the fabric and over NCCL, and BASELINE configs[4] (C5, 3D primitive, geometry gradients) at N=8.
# Usage: gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_multi8.sh'
N=8
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/multi8_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/multi8_pytest.log
grep -E "^FAILED|^ERROR|passed|failed|rc=|Error|assert " gpurun_out/multi8_pytest.log | head -20 | cut -c1-300
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/exchange_probe.py 2>&1 | grep -v "OMP\|\*\*\*" | tee gpurun_out/probe8.log
run() {  # name, fabric mode, bench args
  TS2D_FABRIC=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline $3 > gpurun_out/multi8_bench_$1.json 2> gpurun_out/multi8_bench_$1.err
  tail -2 gpurun_out/multi8_bench_$1.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/multi8_bench_$1.json")); print("N=$N $1", round(d["value"],1), "fps", round(d["ms_per_step"],3), "ms  e2e", d.get("e2e",{}).get("value"), d["config"].get("exchange"), "check", d.get("check"), {k:round(v["ms"],3) for k,v in d.get("stages",{}).items()})
except Exception as ex: print("bench FAILED", ex)
PY
}
run C3_fabric auto "--config C3"
run C3_nccl 0 "--config C3 --no-check --no-e2e --no-model-step"
run C5_3D_fabric auto "--config C5 --primitive 3D --no-model-step"
