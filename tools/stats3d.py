import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch, harness
from harness import rel_err, frac_above
from triangle_splatting_b200 import _C
from triangle_splatting_b200.scenes import make_config
dev = torch.device("cuda:0")
for kw in (dict(), dict(gamma=7.0, opacity_ste=0.3, sh_degree=0)):
    sc = make_config("C2", **kw)
    _C.set_exact(True); a = harness.run_ours(sc, dev, primitive="3D")
    a2 = harness.run_ours(sc, dev, primitive="3D")
    _C.set_exact(False); b = harness.run_ours(sc, dev, primitive="3D")
    print("==", kw)
    for k in harness.GRAD_KEYS:
        for (x, y, nm) in ((a2[k], a[k], "exact~exact"), (b[k], a[k], "fast~exact")):
            print(f"  {k:13s} {nm:12s} max(floor1e-3) {rel_err(x,y,1e-3):.2e} max(1e-2) {rel_err(x,y,1e-2):.2e} max(1e-1) {rel_err(x,y,1e-1):.2e} "
                  f"frac>1e-4 {frac_above(x,y,1e-4,1e-2):.2e} frac>1e-3 {frac_above(x,y,1e-3,1e-2):.2e} frac>1e-2 {frac_above(x,y,1e-2,1e-2):.2e}")
