# One short gpurun call: the GPU test suite (optionally a -k filter in $1) and a bench line with the stage table.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q ${1:+-k "$1"} > gpurun_out/quick_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/quick_pytest.log; grep -E "^FAILED|^ERROR|passed|failed|rc=|AssertionError:" gpurun_out/quick_pytest.log | cut -c1-330
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err; tail -3 gpurun_out/quick_bench.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/quick_bench.json")); print("bench", round(d["value"],1), "fps", round(d["ms_per_step"],3), "ms  e2e", d.get("e2e",{}).get("value"), "model", (d.get("model_step") or {}).get("value"), {k:round(v["ms"],3) for k,v in d.get("stages",{}).items()})
except Exception as ex: print("bench FAILED", ex)
PY
