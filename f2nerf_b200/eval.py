"""Whole-image evaluation — SURVEY §8(f) N4: ``ExpRunner::RenderWholeImage`` (src/ExpRunner.cpp:257-293).

The reference renders an image in fixed chunks of 8192 rays: every chunk is copied host->device, rendered
(``Renderer::Render`` in VALIDATE mode under a NoGradGuard), and three results are copied device->host and written
into CPU tensors with ``index_put_`` — 3 blocking H2D + 3 blocking D2H copies and several ``.item()``-style syncs per
chunk.  Same chunking and the same returned tensors here (CPU ``pred_colors [N,3]``, ``first_oct_disp [N,1]``,
``pred_disp [N,1]`` with the reference's normalisations), but the rays are uploaded once (pinned, asynchronous), the
per-chunk results are written straight into device-resident images, and there is ONE device->host copy at the end.
Each chunk is the march + ONE fused kernel (``Renderer.render_forward`` -> ``f2b_render_fwd_fused``, csrc/fused_fwd.cu) writing its
rows of the image; chunks alternate between two streams / scratch sets so a chunk's march runs under its predecessor's fused kernel.
"""
import torch

from .sampler import VALIDATE

RAY_BATCH_SIZE = 8192          # ExpRunner.cpp:268


@torch.no_grad()
def RenderWholeImage(renderer, rays_o, rays_d, bounds=None, ray_batch_size=RAY_BATCH_SIZE, device=None):
    """-> (pred_colors [N,3], first_oct_disp [N,1], pred_disp [N,1]) CPU float32, like the reference.
    ``rays_o`` / ``rays_d``: [N,3] CPU or CUDA tensors; ``bounds`` is accepted and ignored (the sampler marches
    [global_near_, 1e8], PersSampler.cu:322-323)."""
    gdp = renderer.global_data_pool_
    dev = torch.device(device) if device is not None else renderer.app_emb_.device
    n_rays = rays_d.shape[0]

    def to_dev(x):
        x = x.to(torch.float32)
        if x.is_cuda:
            return x.contiguous()
        return (x.contiguous().pin_memory() if torch.cuda.is_available() else x).to(dev, non_blocking=True)

    o, d = to_dev(rays_o), to_dev(rays_d)
    colors = torch.zeros((n_rays, 3), dtype=torch.float32, device=dev)
    first = torch.full((n_rays, 1), 1., dtype=torch.float32, device=dev)
    disp = torch.zeros((n_rays, 1), dtype=torch.float32, device=dev)
    prev_mode = gdp.mode_
    gdp.mode_ = VALIDATE                                         # the reference's callers set it (ExpRunner.cpp:323,342)
    try:
        if renderer._fused_forward_ok(n_rays):
            # Tiled renderer: every chunk is march + ONE fused kernel (Renderer.render_forward) that writes its rows of the
            # device-resident image directly; no host sync, no per-chunk copies.  Two chunks are in flight on two streams with
            # two march scratch sets, so chunk k+1's march (per-ray latency bound, ~20 % occupancy) runs under chunk k's
            # fused kernel (gather bound).
            main = torch.cuda.current_stream(dev)
            lanes = [renderer._side_stream(dev, 3), renderer._side_stream(dev, 4)]
            ready = torch.cuda.Event()
            ready.record(main)
            disp_flat, depth = disp.view(-1), torch.empty((n_rays,), dtype=torch.float32, device=dev)
            alive = []
            for k, i in enumerate(range(0, n_rays, ray_batch_size)):
                j = min(i + ray_batch_size, n_rays)
                st = lanes[k % 2]
                if k < 2:
                    st.wait_event(ready)
                with torch.cuda.stream(st):
                    r = renderer.render_forward(o[i:j], d[i:j], out=(colors[i:j], disp_flat[i:j], depth[i:j]), lane=1 + k % 2)
                    first[i:j] = r.first_oct_dis.reshape(-1, 1)
                alive.append(r)                                  # allocations of a lane stay referenced until the join below
            for st in lanes:
                main.wait_stream(st)
            del alive
        else:
            for i in range(0, n_rays, ray_batch_size):
                j = min(i + ray_batch_size, n_rays)
                r = renderer.Render(o[i:j], d[i:j], None, None)
                colors[i:j] = r.colors
                disp[i:j, 0] = r.disparity.reshape(-1)
                if r.first_oct_dis is not None and r.first_oct_dis.numel() == (j - i):
                    first[i:j] = r.first_oct_dis.reshape(-1, 1)
    finally:
        gdp.mode_ = prev_mode
    disp = disp / disp.max()                                     # ExpRunner.cpp:289-290
    first = first.min() / first
    out = torch.cat([colors, first, disp], 1).cpu()              # the one device->host copy
    return out[:, :3].contiguous(), out[:, 3:4].contiguous(), out[:, 4:5].contiguous()
