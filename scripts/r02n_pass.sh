#!/bin/bash
# whole-image loop: which march build / walker count overlaps best with the fused kernel of the previous chunk
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; TAG=${TAG:-r02n}
cat > /tmp/wi.py <<'PY'
import json, os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, bench
from types import SimpleNamespace
from f2nerf_b200 import RenderWholeImage
prob = bench.build_problem(0, 1, SimpleNamespace(config="wanjinyou", rays=0), torch.device("cuda", 0))
o, d, cam, gt = prob["host"]
big_o, big_d = torch.from_numpy(o).cuda().repeat(16, 1), torch.from_numpy(d).cuda().repeat(16, 1)
r = prob["renderer"]
out = {}
for chunk in (8192, 16384):
    for _ in range(2):
        RenderWholeImage(r, big_o, big_d, ray_batch_size=chunk)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        RenderWholeImage(r, big_o, big_d, ray_batch_size=chunk)
    torch.cuda.synchronize(); w = (time.perf_counter() - t0) / 5
    out[f"chunk{chunk}"] = {"ms_per_image": round(w * 1e3, 2), "rays_per_s": round(big_o.shape[0] / w)}
print(json.dumps({"env": {k: os.environ.get(k) for k in ("F2B_EVAL_MARCH_BG", "F2B_FUSED_WALKERS")}, "rays": int(big_o.shape[0]), **out}))
PY
for V in "0 1" "1 1" "0 2" "1 2"; do set -- $V
  F2B_EVAL_MARCH_BG=$1 F2B_FUSED_WALKERS=$2 timeout 300 python /tmp/wi.py 2>/dev/null | tail -n 1
done
