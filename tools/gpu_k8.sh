# One gpurun call: 2D parity (golden scenes, live reference, truth, benchmarked sizes C2 / C3) + the bench stage table.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -k "not 3d and not 3D and not C4 and not C5" > gpurun_out/k8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/k8_pytest.log; grep -E "^FAILED|^ERROR|passed|failed|rc=|AssertionError:" gpurun_out/k8_pytest.log | cut -c1-330
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-model-step --no-e2e > gpurun_out/k8_bench.json 2> gpurun_out/k8_bench.err; tail -3 gpurun_out/k8_bench.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/k8_bench.json")); print("bench", round(d["value"],1), "fps", round(d["ms_per_step"],3), "ms", {k:round(v["ms"],3) for k,v in d.get("stages",{}).items()})
except Exception as ex: print("bench FAILED", ex)
PY
