# One gpurun call: A/B of an environment switch (bench + parity at both settings) and the single-triangle tile audit.
# Usage: gpurun --timeout 1200 -- 'bash tools/gpu_ab.sh'
mkdir -p gpurun_out
rm -f gpurun_out/ab_*.log
timeout 500 python tools/debug_tile.py C3 641034 > gpurun_out/ab_tile.log 2>&1; tail -60 gpurun_out/ab_tile.log | cut -c1-260
for tp in 0 1; do
  TS2D_TWOPHASE=$tp timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-model-step --no-check > gpurun_out/ab_bench_$tp.json 2> gpurun_out/ab_bench_$tp.err; tail -2 gpurun_out/ab_bench_$tp.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_bench_$tp.json")); print("TWOPHASE=$tp bench", round(d["value"],1), "fps", round(d["ms_per_step"],3), "ms", {k:round(v["ms"],3) for k,v in d.get("stages",{}).items()})
except Exception as ex: print("bench FAILED", ex)
PY
done
TS2D_TWOPHASE=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -k "not C5 and not C4" > gpurun_out/ab_pytest_tp1.log 2>&1; echo "rc=$?" >> gpurun_out/ab_pytest_tp1.log; grep -E "^FAILED|^ERROR|passed|failed|rc=|AssertionError:" gpurun_out/ab_pytest_tp1.log | cut -c1-330
