# One gpurun --gpus N call: multi-GPU parity tests, then the tile-sharded bench over the peer-memory fabric and over NCCL.
# Usage: gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_multi.sh 2 [C3]'
N=${1:-2}
CFG=${2:-C3}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/multi${N}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/multi${N}_pytest.log
grep -E "^FAILED|^ERROR|passed|failed|rc=|Error|assert " gpurun_out/multi${N}_pytest.log | head -20 | cut -c1-300
for fab in auto 0; do
  TS2D_FABRIC=$fab timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --config $CFG $BENCH_EXTRA > gpurun_out/multi${N}_bench_${CFG}_$fab.json 2> gpurun_out/multi${N}_bench_${CFG}_$fab.err
  tail -3 gpurun_out/multi${N}_bench_${CFG}_$fab.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/multi${N}_bench_${CFG}_$fab.json")); print("N=$N fabric=$fab", round(d["value"],1), "fps", round(d["ms_per_step"],3), "ms  e2e", d.get("e2e",{}).get("value"), d["config"].get("exchange"), "check", d.get("check"), {k:round(v["ms"],3) for k,v in d.get("stages",{}).items()})
except Exception as ex: print("bench FAILED", ex)
PY
done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-model-step --no-check --config $CFG $BENCH_EXTRA > gpurun_out/multi_n1_bench_${CFG}.json 2> gpurun_out/multi_n1_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/multi_n1_bench_${CFG}.json")); print("N=1", round(d["value"],1), "fps", round(d["ms_per_step"],3), "ms", {k:round(v["ms"],3) for k,v in d.get("stages",{}).items()})
except Exception as ex: print("bench FAILED", ex)
PY
