"""Where the encode's and the scatter's time goes, level by level — the data behind DESIGN.md's answer to "TMA-staged hash-table
tiles" (BASELINE.json north_star).  On the headline batch (bench.build_problem, the reference's ngp_fox scene, 4096 rays):

  * time of f2b_hash_fwd_levels / f2b_hash_bwd_levels for each group of 4 levels (CUDA events, 10 launches);
  * for each level, the number of DISTINCT grid cells per 128-sample tile (= the tile a CTA of the fused field kernel encodes)
    and per 32-sample warp — what a shared-memory staging of "the tile's volume" would have to hold, and the share of
    gathers it could serve (1 - distinct/128).

    python scripts/level_probe.py > gpurun_out/level_probe.json
"""
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def timed(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    import bench
    from f2nerf_b200 import VALIDATE, ops
    from f2nerf_b200._lib import call, stream
    dev = torch.device("cuda", 0)
    prob = bench.build_problem(0, 1, SimpleNamespace(config="wanjinyou", rays=0), dev)
    field, sampler, gdp = prob["field"], prob["sampler"], prob["gdp"]
    o, d, cam, gt = (torch.from_numpy(x).to(dev) for x in prob["host"])
    gdp.mode_ = VALIDATE
    s = sampler.GetSamples(o, d)
    pts, anc = s.pts.contiguous(), s.anchors.contiguous()
    n = pts.shape[0]
    table16 = field.table_f16()
    args = (field.prim_pool_, field.bias_pool_, int(field.n_volumes_), int(field.local_size_))
    out = torch.zeros((n, 32), dtype=torch.float16, device=dev)
    dfeat = (torch.randn((n, 32), device=dev) * 1e-3).half()
    d_table = torch.zeros_like(field.feat_pool_)
    res = {"n_samples": int(n), "n_rays": int(o.shape[0]), "groups": []}
    res["encode_all_ms"] = timed(lambda: call("f2b_hash_fwd", table16, *args, pts, anc, 3, n, out, stream()))
    res["scatter_all_ms"] = timed(lambda: call("f2b_hash_bwd", *args, pts, anc, 3, n, dfeat, 1, 1.0, d_table, stream()))
    for lo in (0, 4, 8, 12):
        f = timed(lambda: call("f2b_hash_fwd_levels", table16, *args, pts, anc, 3, n, lo, 4, out, stream()))
        b = timed(lambda: call("f2b_hash_bwd_levels", *args, pts, anc, 3, n, dfeat, 1, 1.0, d_table, lo, 4, stream()))
        res["groups"].append({"levels": [lo, lo + 3], "encode_ms": f, "scatter_ms": b})
    # locality: distinct (volume, cell) per 128-sample tile / 32-sample warp, per level (host side, exact integer cells)
    scales = ops.hash_level_scales().numpy()
    x = ((pts.cpu().numpy() + 1.0) * 0.5).astype(np.float32)
    vol = anc[:, 0].cpu().numpy().astype(np.int64)
    bias = field.bias_pool_.cpu().numpy()
    bounds = s.pts_idx_bounds.cpu().numpy()
    take = np.concatenate([np.arange(a, a + ((b - a) // 128) * 128) for a, b in bounds[:512]])        # whole tiles of 512 rays
    loc = []
    for l in range(16):
        p = x[take] * scales[l] + bias[l * field.n_volumes_ + vol[take]]
        cell = np.floor(p).astype(np.int64)
        key = ((vol[take] * 4096 + cell[:, 0] % 4096) * 4096 + cell[:, 1] % 4096) * 4096 + cell[:, 2] % 4096
        per_tile = np.array([len(np.unique(k)) for k in key.reshape(-1, 128)])
        per_warp = np.array([len(np.unique(k)) for k in key.reshape(-1, 32)])
        loc.append({"level": l, "scale": float(scales[l]), "distinct_cells_per_128_tile": float(per_tile.mean()),
                    "distinct_cells_per_32_warp": float(per_warp.mean())})
    res["locality"] = loc
    print(json.dumps(res))


if __name__ == "__main__":
    main()
