#!/usr/bin/env python
"""bench.py — training rays/s of the F2-NeRF per-ray rendering hot path on B200.

One "step" = one pass of the hot path over one batch: Renderer.Render (perspective-warp ray march,
early-stop pass, hash encode, density + colour MLPs, composite) + the trainer's loss + backward down
to the parameter gradients (+ the NCCL gradient / octree-vote all-reduce when N > 1).  The optimizer
step is outside the path.  Workload: BASELINE.json configs[1] shape — 4096 rays x
<= 1024 samples per ray per GPU, log2_table_size 19, wanjinyou.yaml sampler settings — on a synthetic
scene (the reference dataset does not travel to the GPU box), random-init table / MLPs.

  python bench.py --gpus N --steps K --warmup W            (torchrun launches N > 1)
  python bench.py --impl reference ...                      CPU port of the path (oracle), rank 0 only
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

N_RAYS = 4096
SAMPLE_L = 1.0 / 256
NEAR = 0.01
SCALE_BY_DIS = True
LOG2_TABLE = 19


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tf=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tf=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed region, sampled in-process through NVML (no nvidia-smi
    subprocesses: they stall kernel launches for tens of ms each)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.err, self.h = index, [], False, None, None
        try:                                     # nvmlInit takes ~0.3 s and a driver lock: do it before the timed region
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[index]) if vis and vis.split(",")[0].isdigit() else index
            self.h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.err = str(e)[:120]

    def run(self):
        if self.h is None:
            return
        import pynvml as nv
        try:
            while not self.stop_flag:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetCurrentClocksEventReasons(self.h),
                                  nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
                time.sleep(0.2)
        except Exception as e:  # noqa: BLE001
            self.err = str(e)[:120]

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {self.err}"]}
        import pynvml as nv
        sm = sorted(r[0] for r in self.rows)
        bits = 0
        for r in self.rows:
            bits |= r[1]
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        return {"sm_mhz": float(sm[len(sm) // 2]), "sm_max_mhz": float(self.max_mhz), "reasons": [k for k, v in names.items() if bits & v],
                "power_w_max": max(r[2] for r in self.rows), "samples": len(self.rows)}


def build_problem(rank, n_rays, log2_table, device):
    import torch
    from f2nerf_b200 import TRAIN, GlobalDataPool, Hash3DAnchored, PersSampler, Renderer, SHShader
    from f2nerf_b200.scene import SyntheticScene
    sc = SyntheticScene(n_cams=24, seed=0)                       # identical octree on every rank
    nodes, trans, edges = sc.blobs()
    gdp = GlobalDataPool()
    torch.manual_seed(2022)                                       # replicated parameters
    sampler = PersSampler(gdp, nodes, trans, edges, near=NEAR, sample_l=SAMPLE_L, scale_by_dis=SCALE_BY_DIS, device=device)
    field = Hash3DAnchored(gdp, log2_table_size=log2_table, device=device)
    field.Reset()
    shader = SHShader(gdp, device=device)
    renderer = Renderer(gdp, sampler, field, shader, n_images=len(sc.c2w), use_app_emb=True, device=device)
    gdp.mode_ = TRAIN
    o, d, cam = sc.rays(n_rays, seed=1234 + rank)                 # rank-sharded rays (weak scaling)
    rng = np.random.default_rng(99 + rank)
    gt = rng.random((n_rays, 3), dtype=np.float32)
    return dict(scene=sc, gdp=gdp, sampler=sampler, field=field, shader=shader, renderer=renderer,
                host=(o, d, cam, gt), blobs=(nodes, trans, edges))


def train_step(prob, rays_o, rays_d, emb_idx, gt, dist_sync=None):
    """ExpRunner::Train's use of the path (src/ExpRunner.cpp:93-130) minus the optimizer step."""
    import torch
    from f2nerf_b200 import CustomOps
    r = prob["renderer"]
    for p in (prob["field"].feat_pool_, prob["field"].mlp_.params_, prob["shader"].mlp_.params_, r.app_emb_):
        p.grad = None
    res = r.Render(rays_o, rays_d, None, emb_idx)
    color_loss = torch.sqrt((res.colors - gt) ** 2 + 1e-4).mean()
    var_loss = torch.sqrt(CustomOps.WeightVar(res.weights, res.idx_start_end) + 1e-2).mean()
    tv_loss = ((res.edge_feats[:, 0] - res.edge_feats[:, 1]) ** 2).mean()
    loss = color_loss + var_loss * 1e-2 + tv_loss * 1e-1
    loss.backward()
    if dist_sync is not None:
        dist_sync(prob)
    return loss, res


ALGO_BYTES = {  # algorithmic bytes per unit (sample) at the operator boundary — DESIGN.md "roofline accounting"
    "f2b_field_fwd": 16 + 512 + 64, "f2b_field_fwd_slots": 16 + 512 + 64, "f2b_sampler_march": 28, "f2b_compact_slots": 28 + 64 + 44 + 64,
    "f2b_hash_fwd": 16 + 512 + 64, "f2b_hash_bwd": 16 + 64 + 512, "f2b_sampler_fill": 44, "f2b_sampler_count": 0,
    "f2b_composite_fwd": 28, "f2b_composite_bwd": 24 + 16 + 16, "f2b_early_stop": 8 + 9, "f2b_compact_samples": 88,
    "f2b_shader_prep": 64 + 12 + 64, "f2b_shader_act": 32 + 12, "f2b_cast_f16_to_f32": 6, "f2b_cast_f32_to_f16": 6,
}


NCU_KERNEL = {"f2b_hash_bwd": "hash_bwd_kernel<1>", "f2b_sampler_march": "march16_kernel<2>",
              "f2b_field_fwd_slots": "field_fwd_kernel<1, 4>", "f2b_field_fwd": "field_fwd_kernel<1, 4>",
              "f2b_mlp_bwd2": "mlp_bwd_tc_kernel<1>",
              "f2b_compact_slots": "compact_slots_kernel", "f2b_composite_fwd": "composite_fwd_kernel",
              "f2b_composite_bwd": "composite_bwd_kernel<0>", "f2b_composite_act_bwd": "composite_bwd_kernel<1>"}


def ncu_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel behind C-ABI call ``name``, from the
    committed `ncu --set full` capture of this same command (profiles/*_traffic.json, newest round); None if absent."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if not files or name not in NCU_KERNEL:
        return None
    try:
        k = json.load(open(files[-1]))["kernels"].get(NCU_KERNEL[name])
        return None if k is None else {"dram_bytes_per_launch": k["dram_bytes_per_launch"], "source": os.path.basename(files[-1])}
    except Exception:  # noqa: BLE001
        return None


def unit_count(name, ints):
    """number of samples (units) a traced call processed, from its integer arguments."""
    pos = {"f2b_field_fwd": 3, "f2b_hash_fwd": 3, "f2b_hash_bwd": 3, "f2b_mlp_fwd": 1, "f2b_mlp_bwd": 1, "f2b_shader_prep": 0,
           "f2b_shader_act": 0, "f2b_shader_act_bwd": 0, "f2b_shader_prep_bwd": 0, "f2b_cast_f16_to_f32": 0,
           "f2b_cast_f32_to_f16": 0, "f2b_table_to_half": 0}
    if name in pos and len(ints) > pos[name]:
        return ints[pos[name]]
    return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from f2nerf_b200 import _lib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist_sync = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        from f2nerf_b200.dist import allreduce_step
        dist_sync = allreduce_step
    prob = build_problem(rank, args.rays, args.log2_table, device)
    if world > 1:
        from f2nerf_b200.dist import install_vote_sync
        install_vote_sync(prob["sampler"])
    o, d, cam, gt = prob["host"]
    pin = lambda a: torch.from_numpy(a).pin_memory()
    h_o, h_d, h_cam, h_gt = pin(o), pin(d), pin(cam), pin(gt)
    d_o, d_d, d_cam, d_gt = (x.to(device) for x in (h_o, h_d, h_cam, h_gt))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value): K steps bracketed by barrier + synchronize, CUDA events --------
    clocks = ClockSampler(local)                 # polling starts before the warm-up: the first NVML queries of a
    clocks.start()                               # process stall kernel submission for 100s of ms on this driver
    for _ in range(args.warmup):                 # same object lifetimes as the timed loop (the caching allocator must have
        loss, res = train_step(prob, d_o, d_d, d_cam, d_gt, dist_sync)   # seen the steady-state peak before timing starts)
    barrier()
    clocks.rows.clear()                          # keep only samples taken under the timed regions
    def timed_loop():
        import gc
        gc.collect()
        gc.disable()                              # like timeit: no cyclic-GC pause inside the timed region
        try:
            return _timed_loop()
        finally:
            gc.enable()

    def _timed_loop():
        barrier()
        l0 = _lib.LAUNCHES
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        ns = nk = 0
        w = []
        for _ in range(args.steps):
            w0 = time.perf_counter()
            loss, res = train_step(prob, d_o, d_d, d_cam, d_gt, dist_sync)
            ns += prob["renderer"].n_sampled_pts_
            nk += res.weights.shape[0]
            w.append(round((time.perf_counter() - w0) * 1e3, 2))
        eb.record()
        barrier()
        return ea.elapsed_time(eb), w, ns, nk, _lib.LAUNCHES - l0

    # A step whose host wall time is far off the median (seen on fresh boxes: one NVML poll of the clock sampler
    # holding the driver lock for ~40 ms while the main thread launches) makes the whole K-step number a
    # measurement of that stall: such a run is rejected and the K steps are timed ONCE more; both are reported.
    attempts = []
    for _ in range(2):
        ms, walls, n_samples, n_kept, launches = timed_loop()
        med = sorted(walls)[len(walls) // 2]
        outlier = max(walls[1:] or walls) > 1.5 * med and max(walls[1:] or walls) - med > 2.5
        attempts.append({"ms_per_step": ms / args.steps, "host_wall_ms_per_step": walls, "rejected": bool(outlier)})
        if not outlier:
            break
        if world > 1:
            break                                 # ranks must take the same branch: no re-measure under torchrun
    # ---- per-kernel CUDA-event trace over the same steps (separate loop: event pairs around every C-ABI call) --
    _lib.TRACE = []
    for _ in range(args.steps):
        train_step(prob, d_o, d_d, d_cam, d_gt, dist_sync)
    barrier()
    trace, _lib.TRACE = _lib.TRACE, None
    # ---- end-to-end timing: pinned host rays -> device, loss -> host, every step ----------------
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        ro, rd = h_o.to(device, non_blocking=True), h_d.to(device, non_blocking=True)
        rc, rg = h_cam.to(device, non_blocking=True), h_gt.to(device, non_blocking=True)
        loss, res = train_step(prob, ro, rd, rc, rg, dist_sync)
        loss_host = float(loss.item())
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    clocks.stop_flag = True
    clocks.join(timeout=2)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        cnt = torch.tensor([n_samples, n_kept], device=device, dtype=torch.float64)
        dist.all_reduce(cnt)
        n_samples, n_kept = float(cnt[0]), float(cnt[1])
    # ---- per-kernel breakdown from the traced events --------------------------------------------
    agg = {}
    for name, a, b, ints in trace:
        u = unit_count(name, ints)
        rec = agg.setdefault(name, dict(ms=0.0, calls=0, units=0))
        rec["ms"] += a.elapsed_time(b); rec["calls"] += 1; rec["units"] += (u or 0)
    peaks = load_peaks()
    total_traced = sum(v["ms"] for v in agg.values())
    top = max(agg.items(), key=lambda kv: kv[1]["ms"])
    name, rec = top
    per_launch_ms = rec["ms"] / rec["calls"]
    if name in ("f2b_mlp_fwd", "f2b_mlp_bwd"):
        flops = {"f2b_mlp_fwd": 2 * 3072, "f2b_mlp_bwd": 4 * 3072}[name] * (rec["units"] / rec["calls"])
        roof = dict(bound="tensor", achieved=flops / (per_launch_ms * 1e-3) / 1e12, peak=peaks["tf"], unit="TFLOP/s")
    else:
        units = rec["units"] / rec["calls"] if rec["units"] else n_samples / args.steps / max(world, 1)
        byts = ALGO_BYTES.get(name, 0) * units
        roof = dict(bound="hbm", achieved=byts / (per_launch_ms * 1e-3) / 1e9, peak=peaks["hbm"], unit="GB/s")
    roof.update(frac=roof["achieved"] / roof["peak"], traffic=ncu_traffic(name), kernel=name, ms_per_launch=per_launch_ms,
                share_of_step=rec["ms"] / max(total_traced, 1e-9), peak_source=peaks["src"])
    rays_total = args.rays * world * args.steps
    line = {
        "metric": "training rays/sec", "value": rays_total / (ms * 1e-3), "unit": "rays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 march/composite, f16 hash table + MLP operands (f32 accumulate)",
        "data": "synthetic", "samples_per_sec": n_samples / (ms * 1e-3),
        "config": {"workload": "ngp_fox-shaped batch (BASELINE configs[1]): 4096 rays x <=1024 samples per GPU, "
                               "wanjinyou.yaml sampler (near 0.01, scale_by_dis, sample_l 1/256, fineness 1), log2_table_size "
                               f"{args.log2_table}, synthetic 24-camera scene, random-init table/MLPs, Render+loss+backward",
                   "rays_per_gpu": args.rays, "samples_per_ray": n_samples / args.steps / world / args.rays,
                   "kept_per_ray": n_kept / args.steps / world / args.rays, "parallelism": f"dp{world} (rays sharded, params replicated)",
                   "l2": "per-step working set (>=184 MB of samples + 64 MB table) exceeds the 126 MB L2; no explicit flush",
                   "mlp_impl": int(_lib.lib.f2b_get_mlp_impl())},
        "e2e": {"value": rays_total / (ms_e2e * 1e-3), "unit": "rays/s",
                "h2d_bytes_per_step": int(sum(x.numel() * x.element_size() for x in (h_o, h_d, h_cam, h_gt))) * world,
                "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "host_wall_ms_per_step": walls, "timing_attempts": attempts,
        "clocks": clocks.summary(),
        "roofline": roof,
        "kernels": {k: {"ms_per_step": v["ms"] / args.steps, "calls_per_step": v["calls"] / args.steps} for k, v in
                    sorted(agg.items(), key=lambda kv: -kv[1]["ms"])},
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_port(prob, args, budget_s=20.0)
        if world == 1:
            line["optimizer_step"] = optimizer_timing(prob)
            line["forward_only"] = forward_only_timing(prob, d_o, d_d, args)
            line["ray_generation"] = ray_generation_timing(prob, args)
        ref_gpu = reference_gpu_timing(args)
        if ref_gpu is not None:
            line["reference_gpu"] = ref_gpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def forward_only_timing(prob, d_o, d_d, args, iters=20):
    """Forward-only rays/s (the path ExpRunner::RenderWholeImage drives, ExpRunner.cpp:257-293): VALIDATE mode
    (noise == 1, bg 0.5, no octree votes / edge samples), no autograd, same ray batch as the headline."""
    import torch
    from f2nerf_b200 import TRAIN, VALIDATE
    gdp, r = prob["gdp"], prob["renderer"]
    gdp.mode_ = VALIDATE
    try:
        with torch.no_grad():
            for _ in range(3):
                r.Render(d_o, d_d, None, None)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                r.Render(d_o, d_d, None, None)
            e1.record()
            torch.cuda.synchronize()
    finally:
        gdp.mode_ = TRAIN
    ms = e0.elapsed_time(e1) / iters
    return {"value": args.rays / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms, "mode": "VALIDATE, no_grad",
            "note": "the reference's number is reference_gpu.ms_validate_median on the same ray count"}


def ray_generation_timing(prob, args, iters=20):
    """SURVEY 8f N3 (reported separately): one training batch of rays + ground-truth colours.  `ours` =
    RayGenerator.RandRaysData (CPU index draws, ONE 48 KB H2D copy, two kernels on HBM-resident images);
    `reference_style` = the same draws followed by what Dataset::RandRaysData does (Dataset.cpp:287-296): CPU gather from
    a CPU image tensor, three H2D copies, ray kernel.  ngp_fox geometry: 50 images of 960 x 540."""
    import torch
    from f2nerf_b200 import RayGenerator
    sc = prob["scene"]
    n_img, H, W = 50, 960, 540
    g = torch.Generator().manual_seed(0)
    images = torch.rand((n_img, H, W, 3), generator=g)
    c2w = np.asarray(sc.c2w, np.float32)
    poses = np.tile(c2w[:1, :3, :4], (n_img, 1, 1)).astype(np.float32)
    poses[:len(c2w)] = c2w[:n_img, :3, :4]
    intri = np.tile(np.array([[687.6, 0, 270.], [0, 687.2, 480.], [0, 0, 1]], np.float32), (n_img, 1, 1))
    dist = np.tile(np.array([0.057, -0.0787, -0.0019, -0.0025], np.float32), (n_img, 1))
    bounds = np.tile(np.array([0.1, 10.], np.float32), (n_img, 1))
    gen = RayGenerator(poses, intri, dist, bounds, images=images)
    dev = gen.poses_.device
    flat = images.view(-1, 3)

    def ours():
        gen.RandRaysData(args.rays)

    def ref_style():
        cam = torch.randint(n_img, (args.rays,), dtype=torch.int64)
        i = torch.randint(0, H, (args.rays,), dtype=torch.int64)
        j = torch.randint(0, W, (args.rays,), dtype=torch.int64)
        ij = torch.stack([i, j], -1).to(dev).contiguous()
        gt = flat[cam * H * W + i * W + j].to(dev).contiguous()
        cam_d = cam.to(dev)
        gen.Img2WorldRayFlex(cam_d.to(torch.int32), ij.to(torch.int32))
        return gt, gen.bounds_[cam_d].contiguous()

    out = {}
    for name, fn in (("ours_ms", ours), ("reference_style_ms", ref_style)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
        out[name] = (time.perf_counter() - t0) / iters * 1e3
    out["note"] = "host wall time per 4096-ray batch incl. the CPU index draws; images 50x960x540x3 fp32"
    return out


def optimizer_timing(prob, iters=20):
    """SURVEY 8f N1, reported separately from the headline (the metric excludes the optimizer step): the hash
    table's Adam update as the reference runs it (torch::optim::Adam's ATen sequence, ExpRunner.cpp:136, plus the
    fp32->fp16 table copy the next forward makes) vs f2b_adam_step (one pass over the live prefix + fp16 shadow)."""
    import torch
    from f2nerf_b200 import FusedAdam, ops
    field = prob["field"]
    p = field.feat_pool_
    g = p.grad if p.grad is not None else torch.zeros_like(p)
    pa, m, v = p.detach().clone(), torch.zeros_like(p), torch.zeros_like(p)

    def aten(step):
        bc1, bc2 = 1 - 0.9 ** step, 1 - 0.99 ** step
        m.mul_(0.9).add_(g, alpha=0.1)
        v.mul_(0.99).addcmul_(g, g, value=0.01)
        pa.addcdiv_(m, (v.sqrt() / (bc2 ** 0.5)).add_(1e-15), value=-(1e-2 / bc1))
        ops.table_to_half(pa)

    saved = p.detach().clone()
    opt = FusedAdam([dict(params=[p], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)], table_field=field)
    out = {}
    for name, fn in (("aten_sequence_ms", aten), ("fused_ms", lambda step: opt.step())):
        for i in range(3):
            fn(i + 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(i + 4)
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / iters
    p.data.copy_(saved)                                            # leave the problem as it was
    field.manage_shadow(False)
    out.update(elements=int(p.numel()), live_elements=int(17 * field.local_size_),
               note="hash-table group only; ATen side includes the fp16 table copy of the next forward")
    return out


def reference_gpu_timing(args):
    """Informational: the compiled, unmodified reference (oracle/_ref/ref_driver) timed on the same box."""
    drv = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if args.no_ref_gpu or not os.path.exists(drv) or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return None
    try:
        out_dir = "/tmp/f2b_ref_bench"
        subprocess.run([drv, os.path.join(ROOT, "oracle", "ref_config_ngp_fox.yaml"), out_dir, str(args.rays), "20"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
        return json.load(open(os.path.join(out_dir, "ref_timing.json")))
    except Exception as e:  # noqa: BLE001
        return {"unavailable": str(e)[:200]}


def cpu_port(prob, args, budget_s=20.0):
    """The CPU restatement (oracle) of the same step on a bounded sample of the same workload."""
    import oracle_lib as O
    import oracle_pipeline as OP
    import torch
    nodes, trans, edges = prob["blobs"]
    o, d, cam, gt = prob["host"]
    field, shader, renderer = prob["field"], prob["shader"], prob["renderer"]
    sc = dict(nodes=nodes, trans=trans, edges=edges, near=NEAR, sample_l=SAMPLE_L, scale_by_dis=SCALE_BY_DIS, max_hits=1024)
    fld = dict(table16=field.feat_pool_.detach().cpu().numpy().astype(np.float16), prim=field.prim_pool_.cpu().numpy(),
               bias=field.bias_pool_.cpu().numpy(), V=field.n_volumes_, local_size=field.local_size_,
               mlp_params=field.mlp_.params_.detach().cpu().numpy())
    sp = shader.mlp_.params_.detach().cpu().numpy()
    emb = renderer.app_emb_.detach().cpu().numpy()
    n = 32
    rng = np.random.default_rng(0)
    n_edges = edges.size // 64

    def one(n):
        dn = (torch.from_numpy(d[:n]) / torch.linalg.norm(torch.from_numpy(d[:n]), 2, -1, True)).numpy()
        noise = (rng.random(1024 + n + 10, dtype=np.float32) + .5).astype(np.float32)
        bg = rng.random((n, 3), dtype=np.float32)
        edge = (rng.integers(0, n_edges, 8192).astype(np.int32), (rng.random((8192, 2), dtype=np.float32) * 2 - 1))
        t0 = time.time()
        OP.render_train(sc, o[:n], dn, noise, bg, fld, sp, emb, cam[:n], edge, gt[:n])
        return time.time() - t0
    t = one(n)
    while t < budget_s / 4 and n < args.rays:
        n = min(n * 2, args.rays)
        t = one(n)
    return {"value": n / t, "unit": "rays/s", "cores": O.num_threads(), "kind": "port",
            "sample": f"{n} of the {args.rays} rays of the same batch (all stages incl. backward), {t:.1f} s wall; "
                      "all stages OpenMP-parallel (oracle/f2_oracle.c)"}


def run_reference(args):
    """--impl reference: the reference has no CPU path and cannot be pip-installed (C++/CUDA executable);
    this arm times the CPU port of its algorithm (oracle/) with all host threads on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if "F2B_REF_THREADS" in os.environ or os.environ.get("OMP_NUM_THREADS") == "1":
        # torchrun pins OMP_NUM_THREADS=1 for its workers; the CPU arm is meant to use every host thread it can
        os.environ["OMP_NUM_THREADS"] = os.environ.get("F2B_REF_THREADS", str(os.cpu_count()))
    import torch  # noqa: F401
    # the CPU arm needs the same parameters; build them on the CPU without touching the GPU library
    import oracle_lib as O
    import oracle_pipeline as OP
    from f2nerf_b200.scene import SyntheticScene
    sc0 = SyntheticScene(n_cams=24, seed=0)
    nodes, trans, edges = sc0.blobs()
    V = trans.size // 544
    rng = np.random.default_rng(2022)
    pool = (1 << args.log2_table) * 16
    fld = dict(table16=(rng.random((pool, 2), dtype=np.float32) * 0.02 - 0.01).astype(np.float16),
               prim=(rng.integers(1 << 28, 1 << 30, size=(16, V, 3)).astype(np.int32) | 1),
               bias=(rng.random((16 * V, 3), dtype=np.float32) * 1000 + 100), V=V, local_size=1 << args.log2_table,
               mlp_params=O.mlp_init(32, 0))
    sp = O.mlp_init(32, 1)
    emb = (rng.standard_normal((24, 16)) * .1).astype(np.float32)
    sc = dict(nodes=nodes, trans=trans, edges=edges, near=NEAR, sample_l=SAMPLE_L, scale_by_dis=SCALE_BY_DIS, max_hits=1024)
    o, d, cam = sc0.rays(args.rays, seed=1234)
    gt = rng.random((args.rays, 3), dtype=np.float32)
    n = min(args.rays, args.ref_rays)
    dn = (d[:n] / np.linalg.norm(d[:n], axis=-1, keepdims=True)).astype(np.float32)
    n_edges = edges.size // 64

    def step():
        noise = (rng.random(1024 + n + 10, dtype=np.float32) + .5).astype(np.float32)
        bg = rng.random((n, 3), dtype=np.float32)
        edge = (rng.integers(0, n_edges, 8192).astype(np.int32), (rng.random((8192, 2), dtype=np.float32) * 2 - 1))
        OP.render_train(sc, o[:n], dn, noise, bg, fld, sp, emb, cam[:n], edge, gt[:n])
    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    dt = time.time() - t0
    v = n * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": "training rays/sec", "value": v, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (CPU port)", "data": "synthetic",
        "config": {"workload": f"same batch shape as the product arm; each step = {n} of the {args.rays} rays (bounded sample)",
                   "rays_per_step": n},
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": O.num_threads(), "kind": "port",
                         "sample": f"{n} rays x <=1024 samples per step, all stages incl. backward"},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=N_RAYS)
    ap.add_argument("--log2-table", dest="log2_table", type=int, default=LOG2_TABLE)
    ap.add_argument("--ref-rays", dest="ref_rays", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
