# One gpurun call: the two host layers (compiled pybind11 module / ctypes) over the same library -- parity subset under each,
# host time per step under each, and a bench line.
mkdir -p gpurun_out
for h in native ctypes; do
  TS2D_HOST=$h timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -k "golden or autograd or one_enqueue or error or empty or C2" > gpurun_out/host_pytest_$h.log 2>&1; echo "rc=$?" >> gpurun_out/host_pytest_$h.log
  echo "== $h"; grep -E "^FAILED|^ERROR|passed|failed|rc=|Error" gpurun_out/host_pytest_$h.log | cut -c1-300
  TS2D_HOST=$h timeout 120 python tools/host_overhead.py 2>&1 | tail -2
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-model-step > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err; tail -3 gpurun_out/quick_bench.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/quick_bench.json")); print("bench", round(d["value"],1), "fps", round(d["ms_per_step"],3), "ms  e2e", d.get("e2e",{}).get("value"), {k:round(v["ms"],3) for k,v in d.get("stages",{}).items()})
except Exception as ex: print("bench FAILED", ex)
PY
TS2D_HOST=ctypes timeout 300 python bench.py --config C4 --primitive 3D --steps 20 --warmup 5 --no-cpu-baseline --no-model-step --no-check 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C4-3D ctypes', round(d['value'],1), 'e2e', d['e2e']['value'])"
TS2D_HOST=native timeout 300 python bench.py --config C4 --primitive 3D --steps 20 --warmup 5 --no-cpu-baseline --no-model-step --no-check 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C4-3D native', round(d['value'],1), 'e2e', d['e2e']['value'])"
