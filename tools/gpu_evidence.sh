# One gpurun call that produces everything under profiles/: GPU tests, both bench arms (2D headline + 3D primitive), the ncu launch
# list of the bench command and one `ncu --set full` capture of each of our kernels.  Usage: gpurun --timeout 900 -- 'bash tools/gpu_evidence.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
timeout 200 python bench.py --impl reference --primitive 3D --steps 10 --warmup 3 --no-model-step > gpurun_out/bench_ref_3d_C3.json 2> gpurun_out/bench_ref_3d_C3.err
timeout 200 python bench.py --primitive 3D --steps 10 --warmup 3 --no-cpu-baseline --no-model-step > gpurun_out/bench_ours_3d_C3.json 2> gpurun_out/bench_ours_3d_C3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-model-step > gpurun_out/ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_render|k_emit|k_preprocess|k_ranges" -c 6 -o gpurun_out/prof_all python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-model-step > gpurun_out/ncu_full.log 2>&1
python - <<'PY'
import json
for n in ("bench_ref","bench_ours","bench_ref_3d_C3","bench_ours_3d_C3"):
    try:
        d=json.load(open(f"gpurun_out/{n}.json")); print(n, round(d["value"],1), round(d["ms_per_step"],3), "e2e", d.get("e2e",{}).get("value"), "model", (d.get("model_step") or {}).get("value"), {k:round(v["ms"],3) for k,v in d.get("stages",{}).items()}, d["clocks"].get("sm_mhz"))
    except Exception as ex: print(n, "FAILED", ex)
PY
