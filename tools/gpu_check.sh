# One gpurun call while developing: parity tests first (hard timeout: a look-back spin that never ends must not hang the box),
# then the full GPU suite, a short bench and the determinism / accuracy report at the benchmarked size.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh'
mkdir -p gpurun_out
rm -f gpurun_out/check_*.log

if true; then
  timeout 900 python -m pytest tests -m gpu -q > gpurun_out/check_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/check_pytest_gpu.log; grep -E "^FAILED|^ERROR|passed|failed|rc=|AssertionError:" gpurun_out/check_pytest_gpu.log | cut -c1-330
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err; tail -3 gpurun_out/check_bench.err
  python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/check_bench.json")); print("bench", round(d["value"],1), "fps", round(d["ms_per_step"],3), "ms  e2e", d.get("e2e",{}).get("value"), "model", (d.get("model_step") or {}).get("value"), {k:round(v["ms"],3) for k,v in d.get("stages",{}).items()})
except Exception as ex: print("bench FAILED", ex)
PY
  rm -f gpurun_out/scale_report.txt gpurun_out/scale_report.json
  timeout 400 python tools/scale_report.py --truth C3 > gpurun_out/check_scale.log 2>&1; grep -E "^==|work|dL_d|integer" gpurun_out/check_scale.log | cut -c1-330
  timeout 300 python tools/debug_outlier.py C3 > gpurun_out/check_outlier.log 2>&1; grep -E "^dL|tri " gpurun_out/check_outlier.log | head -16 | cut -c1-200
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-model-step > gpurun_out/ncu_launch.log 2>&1
  python tools/ncu_launches.py gpurun_out/launches.csv
fi
