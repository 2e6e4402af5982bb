# One gpurun call that produces the round-2 evidence under gpurun_out/ev/ (copied to profiles/r02/ afterwards):
# GPU tests, both bench arms, the other configurations, the parity + accuracy report at the benchmarked sizes, the ncu launch list
# of the bench command, one `ncu --set full` pass over our kernels, the host-side profile.
# Usage: gpurun --timeout 2400 -- 'bash tools/gpu_evidence_r02.sh'
E=gpurun_out/ev
mkdir -p $E
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $E/smoke.log 2>&1; tail -1 $E/smoke.log
timeout 900 python -m pytest tests -m gpu -q > $E/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $E/pytest_gpu.log; tail -3 $E/pytest_gpu.log
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $E/bench_reference_n1.json 2> $E/bench_reference_n1.err
timeout 400 python bench.py --steps 20 --warmup 5 > $E/bench_ours_n1.json 2> $E/bench_ours_n1.err
timeout 200 python bench.py --primitive 3D --steps 10 --warmup 3 --no-cpu-baseline --no-model-step > $E/bench_ours_C3_3D.json 2> $E/bench_ours_C3_3D.err
timeout 200 python bench.py --config C4 --primitive 3D --steps 20 --warmup 5 --no-cpu-baseline --no-model-step > $E/bench_ours_C4_3D.json 2> $E/bench_ours_C4_3D.err
timeout 200 python bench.py --config C2 --steps 20 --warmup 5 --no-cpu-baseline > $E/bench_ours_C2.json 2> $E/bench_ours_C2.err
timeout 300 python bench.py --config C5 --primitive 3D --steps 10 --warmup 3 --no-cpu-baseline --no-model-step > $E/bench_ours_C5_3D.json 2> $E/bench_ours_C5_3D.err
python - <<'PY'
import json
for n in ("bench_reference_n1","bench_ours_n1","bench_ours_C3_3D","bench_ours_C4_3D","bench_ours_C2","bench_ours_C5_3D"):
    try:
        d=json.load(open(f"gpurun_out/ev/{n}.json")); print(n, round(d["value"],1), round(d["ms_per_step"],3), "e2e", d.get("e2e",{}).get("value"), "model", (d.get("model_step") or {}).get("value"), {k:round(v["ms"],3) for k,v in d.get("stages",{}).items()}, d["clocks"].get("sm_mhz"), d.get("roofline"))
    except Exception as ex: print(n, "FAILED", ex)
PY
rm -f gpurun_out/scale_report.txt gpurun_out/scale_report.json
timeout 900 python tools/scale_report.py --truth C2 C3 C4_3D C5_3D > $E/scale_report.log 2>&1; cp gpurun_out/scale_report.txt gpurun_out/scale_report.json $E/ 2>/dev/null; grep -E "^==|integer" $E/scale_report.log | cut -c1-250
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $E/ncu_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-model-step --no-check > $E/ncu_launch.log 2>&1
python tools/ncu_launches.py $E/ncu_launches.csv > $E/ncu_launches_summary.txt; head -22 $E/ncu_launches_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_render|k_emit|k_preprocess|k_radix|k_bwd_rows|k_tile_tables|k_scan" -c 40 -o $E/prof_r02 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-model-step --no-check > $E/ncu_full.log 2>&1
ncu -i $E/prof_r02.ncu-rep --page raw --csv > $E/ncu_full_raw.csv 2>/dev/null; wc -l $E/ncu_full_raw.csv
python -m cProfile -s tottime tools/host_overhead.py 2>&1 | head -45 > $E/host_profile.txt; head -3 $E/host_profile.txt
