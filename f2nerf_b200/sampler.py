"""PersSampler — host-side mirror of the reference's perspective-warp octree sampler operator
(``src/PtsSampler/PersSampler.h:75-96``, ``PersSampler.cu:317-615``) over the C ABI.

Same method names and argument meaning as the reference class (``GetSamples``, ``GetEdgeSamples``,
``UpdateOctNodes``, ``States``, ``LoadStates``); tensors are CUDA, contiguous, f32 / i32 like the
reference's ``CUDAFloat`` / ``CUDAInt``.  RNG draws use the same torch calls in the same order as the
reference so that a shared ``torch.manual_seed`` reproduces its noise (``PersSampler.cu:377,456-457``).
Octree (re)construction / compaction (``PersOctree::ProcOctree``) is out of scope:
blobs come from a reference checkpoint / dump or from :mod:`f2nerf_b200.scene`.
"""
from dataclasses import dataclass

import torch

from . import ops
from ._lib import call, stream

MAX_SAMPLE_PER_RAY = 1024
INIT_NODE_STAT = 1000
TRAIN, VALIDATE = 0, 1


@dataclass
class SampleResultFlex:
    """``struct SampleResultFlex`` (src/PtsSampler/PtsSampler.h:13-22)."""
    pts: torch.Tensor              # [n_all_pts, 3] warped coordinates
    dirs: torch.Tensor             # [n_all_pts, 3]
    dt: torch.Tensor               # [n_all_pts]
    t: torch.Tensor                # [n_all_pts]
    anchors: torch.Tensor          # [n_all_pts, 3] i32: trans_idx, node idx, 0
    pts_idx_bounds: torch.Tensor   # [n_rays, 2] i32 start, end
    first_oct_dis: torch.Tensor    # [n_rays, 1]


class GlobalDataPool:
    """The cross-module scalars of ``GlobalDataPool`` (src/Utils/GlobalDataPool.h:10-32) the path reads/writes."""

    def __init__(self):
        self.mode_ = TRAIN
        self.n_volumes_ = 1
        self.iter_step_ = 0
        self.sampled_oct_per_ray_ = 16.0
        self.sampled_pts_per_ray_ = 512.0
        self.meaningful_sampled_pts_per_ray_ = 512.0
        self.learning_rate_ = 1.0
        self.ray_march_fineness_ = 1.0
        self.gradient_scaling_progress_ = 1.0
        self.backward_nan_ = False


class PersSampler:
    def __init__(self, global_data_pool, tree_nodes, pers_trans, edge_pool=None, *, near=0.05, sample_l=1.0 / 256,
                 scale_by_dis=False, max_oct_intersect_per_ray=1024, device="cuda"):
        self.global_data_pool_ = global_data_pool
        dev = torch.device(device)
        as_u8 = lambda b: torch.as_tensor(b, dtype=torch.uint8).reshape(-1).to(dev).contiguous()
        self.tree_nodes_gpu_ = as_u8(tree_nodes)
        self.pers_trans_gpu_ = as_u8(pers_trans)
        self.edge_pool_gpu_ = as_u8(edge_pool) if edge_pool is not None else torch.zeros(0, dtype=torch.uint8, device=dev)
        if self.tree_nodes_gpu_.numel() % 64 or self.pers_trans_gpu_.numel() % 544 or self.edge_pool_gpu_.numel() % 64:
            raise ValueError("PersSampler: blob sizes must be multiples of sizeof(TreeNode)=64 / TransInfo=544 / EdgePool=64")
        n_nodes = self.tree_nodes_gpu_.numel() // 64
        self.tree_weight_stats_ = torch.full((n_nodes,), INIT_NODE_STAT, dtype=torch.int32, device=dev)
        self.tree_alpha_stats_ = torch.full((n_nodes,), INIT_NODE_STAT, dtype=torch.int32, device=dev)
        self.tree_visit_cnt_ = torch.zeros((n_nodes,), dtype=torch.int32, device=dev)
        self.global_near_, self.sample_l_, self.scale_by_dis_ = float(near), float(sample_l), bool(scale_by_dis)
        self.max_oct_intersect_per_ray_ = int(max_oct_intersect_per_ray)
        self.sub_div_milestones_ = []
        self.exact_oct_stat_ = False         # True: run the traversal to exhaustion for the exact "OctSamples" log EMA
        self.vote_allreduce_ = None          # set by f2nerf_b200.dist.install_vote_sync under data parallelism
        global_data_pool.n_volumes_ = self.pers_trans_gpu_.numel() // 544

    @property
    def n_nodes(self):
        return self.tree_nodes_gpu_.numel() // 64

    @property
    def n_edges(self):
        return self.edge_pool_gpu_.numel() // 64

    def make_noise(self, n_rays, device):
        gdp = self.global_data_pool_
        n = MAX_SAMPLE_PER_RAY + n_rays + 10
        if gdp.mode_ == VALIDATE:
            noise = torch.ones(n, dtype=torch.float32, device=device)
        else:
            noise = ((torch.rand(n, dtype=torch.float32, device=device) - .5) + 1.).contiguous()
        return noise.mul_(gdp.ray_march_fineness_)

    def GetSamples(self, rays_o_raw, rays_d_raw, bounds_raw=None, rays_noise=None):
        """PersSampler::GetSamples (PersSampler.cu:317-434).  ``bounds_raw`` is ignored like in the
        reference (it marches [global_near_, 1e8]).  One host sync (sample total) instead of two."""
        rays_o = rays_o_raw.contiguous()
        rays_d = (rays_d_raw / torch.linalg.norm(rays_d_raw, 2, -1, True)).contiguous()
        n_rays = rays_o.shape[0]
        gdp = self.global_data_pool_
        if rays_noise is None:
            rays_noise = self.make_noise(n_rays, rays_o.device)
        args = (self.tree_nodes_gpu_, self.pers_trans_gpu_, rays_o, rays_d, rays_noise, self.global_near_, 1e8,
                self.sample_l_, self.scale_by_dis_, self.max_oct_intersect_per_ray_)
        scratch = self._scratch(n_rays, rays_o.device)
        bounds, totals, first = ops.sampler_march(*args, scratch, count_all_hits=self.exact_oct_stat_)
        n_all_pts, n_all_oct = (int(v) for v in totals.tolist())            # the one sync
        if gdp.mode_ != VALIDATE and n_rays > 0:
            gdp.sampled_oct_per_ray_ = gdp.sampled_oct_per_ray_ * .9 + (n_all_oct / n_rays) * .1
        pts, dirs, dt, t, anchors = ops.sampler_gather(rays_d, bounds, n_all_pts, scratch)
        return SampleResultFlex(pts, dirs, dt, t, anchors, bounds, first)

    def _scratch(self, n_rays, dev):
        """Persistent one-pass march scratch: a 1024-sample slot per ray (28 B/sample), grown on demand."""
        cap = max(n_rays, 1) * MAX_SAMPLE_PER_RAY
        if getattr(self, "_scratch_cap", 0) < cap or self._scratch_bufs[0].device != dev:
            f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
            self._scratch_bufs = (f(cap, 3), f(cap), f(cap), torch.empty((cap, 2), dtype=torch.int32, device=dev))
            self._scratch_cap = cap
        return self._scratch_bufs

    def GetEdgeSamples(self, n_pts):
        """PersSampler::GetEdgeSamples (PersSampler.cu:454-473): (out_pts [n,2,3], out_idx [n,2])."""
        if self.n_edges <= 0:
            raise RuntimeError("GetEdgeSamples: empty edge pool (needs >= 2 face-adjacent valid leaves)")
        dev = self.tree_nodes_gpu_.device
        edge_idx = torch.randint(0, self.n_edges, (n_pts,), dtype=torch.int32, device=dev).contiguous()
        edge_coord = (torch.rand((n_pts, 2), dtype=torch.float32, device=dev) * 2. - 1.).contiguous()
        return ops.edge_samples(self.edge_pool_gpu_, self.pers_trans_gpu_, edge_idx, edge_coord)

    def UpdateOctNodes(self, sample_result, sampled_weight, sampled_alpha):
        """PersSampler::UpdateOctNodes (PersSampler.cu:536-603) without the milestone/compaction calls."""
        n_nodes, n_rays = self.n_nodes, sample_result.pts_idx_bounds.shape[0]
        dev = self.tree_nodes_gpu_.device
        n_pts = sample_result.anchors.shape[0]
        if sampled_weight.shape[0] != n_pts or sampled_alpha.shape[0] != n_pts:
            raise ValueError("UpdateOctNodes: weight/alpha length must equal the number of samples")
        vote_w = torch.full((n_nodes,), -1, dtype=torch.int32, device=dev)
        vote_a = torch.full((n_nodes,), -1, dtype=torch.int32, device=dev)
        mark = torch.zeros((n_nodes,), dtype=torch.int32, device=dev)
        oct_idx = sample_result.anchors.reshape(-1)[1:]                          # anchors[:,1] with stride 3
        call("f2b_oct_mark_visit", sample_result.pts_idx_bounds, n_rays, oct_idx.data_ptr(), 3,
             sampled_weight.contiguous(), sampled_alpha.contiguous(), vote_w, vote_a, mark, self.tree_visit_cnt_, stream())
        if self.vote_allreduce_ is not None:                                     # DP: identical pruning on every rank
            self.vote_allreduce_(vote_w, vote_a, mark, self.tree_visit_cnt_)
        self.last_votes_ = (vote_w, vote_a, mark)
        self.apply_votes(vote_w, vote_a, mark)

    def apply_votes(self, vote_w, vote_a, mark):
        call("f2b_oct_update_stats", vote_w, vote_a, mark, self.tree_weight_stats_, self.tree_alpha_stats_,
             self.tree_nodes_gpu_, self.n_nodes, stream())

    def States(self):
        """PersSampler::States (PersSampler.cpp:692-701): same order / dtypes as the reference checkpoint."""
        ms = torch.tensor(self.sub_div_milestones_, dtype=torch.int32, device=self.tree_nodes_gpu_.device)
        return [self.tree_nodes_gpu_, self.pers_trans_gpu_, self.tree_visit_cnt_, ms]

    def LoadStates(self, states, idx):
        dev = self.tree_nodes_gpu_.device
        self.tree_nodes_gpu_ = states[idx].clone().to(dev).contiguous(); idx += 1
        self.pers_trans_gpu_ = states[idx].clone().to(dev).contiguous(); idx += 1
        self.tree_visit_cnt_ = states[idx].clone().to(dev).contiguous(); idx += 1
        self.sub_div_milestones_ = states[idx].cpu().tolist(); idx += 1
        n_nodes = self.n_nodes
        self.tree_weight_stats_ = torch.full((n_nodes,), INIT_NODE_STAT, dtype=torch.int32, device=dev)
        self.tree_alpha_stats_ = torch.full((n_nodes,), INIT_NODE_STAT, dtype=torch.int32, device=dev)
        self.global_data_pool_.n_volumes_ = self.pers_trans_gpu_.numel() // 544
        return idx
