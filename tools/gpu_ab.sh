# One gpurun call: A/B of environment switches on the bench stage table (+ parity of the integer outputs under each setting).
# Usage: gpurun --timeout 1200 -- 'bash tools/gpu_ab.sh'
mkdir -p gpurun_out
run() {  # label, env assignments
  env $2 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-model-step --no-check > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.err; tail -1 gpurun_out/ab_$1.err | cut -c1-200
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_$1.json")); print("$1", round(d["value"],1), "fps", round(d["ms_per_step"],3), "ms", {k:round(v["ms"],3) for k,v in d.get("stages",{}).items()})
except Exception as ex: print("$1 bench FAILED", ex)
PY
}
run bytes_3072 "TS2D_TILE_DIGITS=8 TS2D_RS_BIG=0"
run balanced_3072 "TS2D_TILE_DIGITS=0 TS2D_RS_BIG=0"
run bytes_4096 "TS2D_TILE_DIGITS=8 TS2D_RS_BIG=1"
run balanced_4096 "TS2D_TILE_DIGITS=0 TS2D_RS_BIG=1"
TS2D_RS_BIG=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -k "not C5 and not C4" > gpurun_out/ab_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/ab_pytest.log; grep -E "^FAILED|^ERROR|passed|failed|rc=|AssertionError:" gpurun_out/ab_pytest.log | cut -c1-330
