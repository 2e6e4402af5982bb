# One short gpurun call: GPU tests of the loss front-end and the timing of both arms' loss terms.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_geometry_loss.py tests/test_gpu_loss.py -m gpu -q > gpurun_out/loss_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/loss_pytest.log
grep -E "^FAILED|^ERROR|passed|failed|rc=|AssertionError|Error" gpurun_out/loss_pytest.log | head -30 | cut -c1-400
python - <<'PY' 2>&1 | tail -6
import json, torch, sys
sys.path.insert(0, ".")
import bench
from triangle_splatting_b200.scenes import make_config
dev = torch.device("cuda:0")
sc = make_config("C3", P=1000)
out = {"ours": bench.geometry_loss_ms(sc, dev, fused=True), "reference": bench.geometry_loss_ms(sc, dev, fused=False),
       "image_ours": bench.image_loss_ms(sc, dev, fused=True), "image_reference": bench.image_loss_ms(sc, dev, fused=False)}
json.dump(out, open("gpurun_out/loss_bench.json", "w"), indent=1)
for k, v in out.items(): print(k, round(v["ms"], 4), "ms")
PY
