#!/usr/bin/env python
"""Generate the committed golden fixture tests/golden/ref_ngp_fox.npz by running the UNMODIFIED reference
(oracle/_ref/ref_driver, built by oracle/Makefile.ref) on a B200:

    gpurun -- python oracle/make_golden.py            # writes gpurun_out/golden/ref_ngp_fox.npz
    mv gpurun_out/golden/ref_ngp_fox.npz tests/golden/

The fixture holds the reference's own ngp_fox octree / warp blobs, a small seeded ray batch and every
boundary tensor of PersSampler::GetSamples, Hash3DAnchored::AnchoredQuery, SHShader::Query and
Renderer::Render for it (VALIDATE mode, plus the seeded TRAIN-mode sampler outputs and octree statistics).
tests/test_golden_ref.py pins oracle/f2_oracle.c against it on the CPU.  The 64 MB hash table is not stored:
both sides regenerate it from the CPU generator seed 1234 (see ref_driver.cpp).
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_RAYS = 40


def main():
    drv = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    out = "/tmp/f2b_golden_dump"
    r = subprocess.run([drv, os.path.join(ROOT, "oracle", "ref_config_ngp_fox.yaml"), out, str(N_RAYS), "0", "1"], cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-4000:])
        sys.exit(1)
    keep = ["tree_nodes", "pers_trans", "edge_pool", "search_order", "prim_pool", "bias_pool", "field_mlp_params",
            "shader_mlp_params", "app_emb", "scalars", "rays_o", "rays_d", "rays_d_normed", "gt_colors", "emb_idx",
            "val_pts", "val_dirs", "val_dt", "val_t", "val_anchors", "val_bounds", "val_first_oct_dis", "val_scene_feat",
            "val_rgb", "val_sh", "val_colors", "val_disparity", "val_depth", "val_weights", "val_idx_start_end",
            "train_noise", "train_bg", "train_edge_idx", "train_edge_coord", "train_pts", "train_dt", "train_t", "train_anchors",
            "train_bounds", "train_colors", "train_disparity", "train_depth", "train_weights", "train_idx_start_end",
            "train_first_oct_dis", "train_tree_nodes_after", "train_weight_stats_after", "train_alpha_stats_after",
            "train_visit_cnt_after", "train_loss", "grad_field_mlp", "grad_shader_mlp", "grad_app_emb",
            "edge_idx", "edge_coord", "edge_pts", "edge_anchors", "train_edge_feats",
            "ds_poses", "ds_intri", "ds_dist_params", "ray_ij", "ds_hw", "ds_bounds", "ds_train_set", "ray_bounds",
            "oct_intri", "oct_w2c", "oct_bound", "backward_nan",
            "oct_nodes_in", "oct_w_in", "oct_a_in", "oct_visit_in", "oct_nodes_sub", "oct_w_sub", "oct_a_sub",
            "oct_nodes_invis", "oct_nodes_final", "oct_w_final", "oct_a_final"]
    data = {k: np.load(os.path.join(out, k + ".npy")) for k in keep}
    for k in ("edge_idx", "edge_coord", "edge_pts", "edge_anchors"):         # 2048 of the 8192 draws are plenty
        data[k] = np.ascontiguousarray(data[k][:2048])
    # the reference's hash-table gradient (64 MB dense) as a seeded 2^18-element subsample of its live prefix + per-slab norms
    g = np.load(os.path.join(out, "grad_feat_pool.npy")).reshape(-1)
    local = ((int(data["scalars"][5]) // 16) >> 4) << 4
    sub = np.sort(np.random.default_rng(0).choice(17 * local, 1 << 18, replace=False)).astype(np.int64)
    data["grad_feat_pool_sub_idx"], data["grad_feat_pool_sub_val"] = sub, g[sub].astype(np.float32)
    data["grad_feat_pool_slab_norm"] = np.array([np.linalg.norm(g[l * local:(l + 1) * local].astype(np.float64)) for l in range(17)], np.float32)
    assert not g[17 * local:].any()
    data["train_edge_feats"] = np.ascontiguousarray(data["train_edge_feats"][:2048]).astype(np.float16)   # fp16 values
    # the level scales as the device computes them (MUFU.EX2): the oracle takes them as an input
    sys.path.insert(0, ROOT)
    from f2nerf_b200 import ops
    data["level_scales"] = ops.hash_level_scales().numpy()
    data["val_scene_feat"] = data["val_scene_feat"].astype(np.float16)       # tcnn outputs are fp16 values
    data["train_edge_coord"] = data["train_edge_coord"].astype(np.float32)
    os.makedirs(os.path.join(ROOT, "gpurun_out", "golden"), exist_ok=True)
    dst = os.path.join(ROOT, "gpurun_out", "golden", "ref_ngp_fox.npz")
    np.savez_compressed(dst, **data)
    print("wrote", dst, os.path.getsize(dst) / 1e6, "MB;", {k: v.shape for k, v in data.items() if k.startswith(("tree", "pers", "val_pts"))})


if __name__ == "__main__":
    main()
