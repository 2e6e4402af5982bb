# One gpurun --gpus N call (short): multi-GPU parity tests, then the tile-sharded bench over the peer-memory fabric (one line).
# Usage: gpurun --gpus 2 --timeout 500 -- 'bash tools/gpu_multi_quick.sh 2'
N=${1:-2}
mkdir -p gpurun_out/multi
timeout 120 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/multi/pytest_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/multi/pytest_n$N.log
grep -E "^FAILED|^ERROR|passed|failed|rc=|Error|assert " gpurun_out/multi/pytest_n$N.log | head -20 | cut -c1-300
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-model-step > gpurun_out/multi/bench_C3_n$N.json 2> gpurun_out/multi/bench_C3_n$N.err
tail -3 gpurun_out/multi/bench_C3_n$N.err | cut -c1-300
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/multi/bench_C3_n$N.json")); print("N=$N", round(d["value"],1), "fps", round(d["ms_per_step"],3), "ms  e2e", d.get("e2e",{}).get("value"), d["config"].get("exchange"), "check", d.get("check"), {k:round(v["ms"],3) for k,v in d.get("stages",{}).items()})
except Exception as ex: print("bench FAILED", ex)
PY
