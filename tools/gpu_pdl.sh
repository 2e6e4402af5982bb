# One gpurun call: the GPU suite with programmatic dependent launch on (the default) -- and once more with TS2D_PDL=0 if anything fails --,
# then the A/B of the frame time with the launch attribute on the short-kernel chains (default) / off (TS2D_PDL=0) / on every launch
# (TS2D_PDL=2) at C3 and C4-3D, the other configurations, the reference arm, and the ncu
# launch list + one full capture of the two composite kernels at HEAD.
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_pdl.sh'
E=gpurun_out/pdl
mkdir -p $E
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $E/smoke.log 2>&1; tail -1 $E/smoke.log
timeout 400 python -m pytest tests -m gpu -q -x > $E/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc" >> $E/pytest_gpu.log; grep -E "^FAILED|^ERROR|passed|failed|rc=|AssertionError:" $E/pytest_gpu.log | cut -c1-300
if [ $rc -ne 0 ]; then
  TS2D_PDL=0 timeout 400 python -m pytest tests -m gpu -q -x > $E/pytest_gpu_pdl0.log 2>&1; echo "pytest(PDL=0) rc=$?" >> $E/pytest_gpu_pdl0.log; grep -E "^FAILED|^ERROR|passed|failed|rc=|AssertionError:" $E/pytest_gpu_pdl0.log | cut -c1-300
fi
show() {
python - "$@" <<'PY'
import json,sys
for n in sys.argv[1:]:
    try:
        d=json.load(open(f"gpurun_out/pdl/{n}.json")); print(n, round(d["value"],1), "fps", round(d["ms_per_step"],4), "ms", "e2e", (d.get("e2e") or {}).get("value"), "model", (d.get("model_step") or {}).get("value"), {k:round(v["ms"],3) for k,v in d.get("stages",{}).items()}, d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"))
    except Exception as ex: print(n, "FAILED", ex)
PY
}
timeout 300 python bench.py --steps 20 --warmup 5 > $E/bench_ours_n1.json 2> $E/bench_ours_n1.err
TS2D_PDL=0 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-model-step --no-check > $E/bench_ours_n1_pdl0.json 2> $E/bench_ours_n1_pdl0.err
TS2D_PDL=2 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-model-step --no-check > $E/bench_ours_n1_pdl2.json 2> $E/bench_ours_n1_pdl2.err
show bench_ours_n1 bench_ours_n1_pdl0 bench_ours_n1_pdl2
timeout 200 python bench.py --config C4 --primitive 3D --steps 40 --warmup 10 --no-cpu-baseline --no-model-step --no-check > $E/bench_ours_C4_3D.json 2> $E/bench_ours_C4_3D.err
TS2D_PDL=0 timeout 200 python bench.py --config C4 --primitive 3D --steps 40 --warmup 10 --no-cpu-baseline --no-model-step --no-check > $E/bench_ours_C4_3D_pdl0.json 2> $E/bench_ours_C4_3D_pdl0.err
TS2D_PDL=2 timeout 200 python bench.py --config C4 --primitive 3D --steps 40 --warmup 10 --no-cpu-baseline --no-model-step --no-check > $E/bench_ours_C4_3D_pdl2.json 2> $E/bench_ours_C4_3D_pdl2.err
show bench_ours_C4_3D bench_ours_C4_3D_pdl0 bench_ours_C4_3D_pdl2
if [ -z "$SKIP_REF" ]; then
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $E/bench_reference_n1.json 2> $E/bench_reference_n1.err
show bench_reference_n1
fi
timeout 200 python bench.py --config C2 --steps 20 --warmup 5 --no-cpu-baseline --no-check > $E/bench_ours_C2.json 2> $E/bench_ours_C2.err
timeout 200 python bench.py --primitive 3D --steps 10 --warmup 3 --no-cpu-baseline --no-model-step --no-check > $E/bench_ours_C3_3D.json 2> $E/bench_ours_C3_3D.err
timeout 300 python bench.py --config C5 --primitive 3D --steps 10 --warmup 3 --no-cpu-baseline --no-model-step --no-check > $E/bench_ours_C5_3D.json 2> $E/bench_ours_C5_3D.err
show bench_ours_C2 bench_ours_C3_3D bench_ours_C5_3D
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $E/ncu_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-model-step --no-check > $E/ncu_launch.log 2>&1
python tools/ncu_launches.py $E/ncu_launches.csv > $E/ncu_launches_summary.txt; head -12 $E/ncu_launches_summary.txt
if [ -z "$SKIP_FULL" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_render_fwd_fast|k_render_bwd_fast" -c 2 -o $E/prof_k78 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-model-step --no-check > $E/ncu_full.log 2>&1
ncu -i $E/prof_k78.ncu-rep --page raw --csv > $E/ncu_full_raw_k78.csv 2>/dev/null; wc -l $E/ncu_full_raw_k78.csv
fi
