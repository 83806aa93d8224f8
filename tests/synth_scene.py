"""Octree / perspective-warp blobs in the reference's byte layout, and a synthetic scene builder.

The ray-march kernels consume the reference's raw struct bytes (``PersSampler.h:15-37``):
``TreeNode`` 64 B, ``TransInfo`` 544 B, ``EdgePool`` 64 B.  This module defines matching numpy
structured dtypes, (de)serialises them, and builds a *synthetic* scene (cameras on a sphere
looking inwards) with the same construction rules as ``PersOctree::ConstructTreeNode`` /
``ConstructTrans`` / ``ConstructEdgePool`` (``src/PtsSampler/PersSampler.cpp:359-660``) so that
benchmarks and tests have valid octree blobs on a box where the reference data is absent.
Octree construction is scene setup (SURVEY.md §8f "next" row N2), not the hot path: it runs in
numpy on the host.
"""
import numpy as np

TREE_NODE = np.dtype({
    "names": ["center", "side_len", "parent", "childs", "is_leaf", "trans_idx"],
    "formats": [("<f4", 3), "<f4", "<i4", ("<i4", 8), "u1", "<i4"],
    "offsets": [0, 12, 16, 20, 52, 56],
    "itemsize": 64,
})
TRANS_INFO = np.dtype({
    "names": ["w2xz", "weight", "center", "dis_summary"],
    "formats": [("<f4", (12, 2, 4)), ("<f4", (3, 12)), ("<f4", 3), "<f4"],
    "offsets": [0, 384, 528, 540],
    "itemsize": 544,
})
EDGE_POOL = np.dtype({
    "names": ["t_idx_a", "t_idx_b", "center", "dir_0", "dir_1"],
    "formats": ["<i4", "<i4", ("<f4", 3), ("<f4", 3), ("<f4", 3)],
    "offsets": [0, 4, 8, 20, 32],
    "itemsize": 64,
})
N_PROS = 12


def view_nodes(blob):
    return np.ascontiguousarray(blob).view(np.uint8).reshape(-1).view(TREE_NODE)


def view_trans(blob):
    return np.ascontiguousarray(blob).view(np.uint8).reshape(-1).view(TRANS_INFO)


def view_edges(blob):
    return np.ascontiguousarray(blob).view(np.uint8).reshape(-1).view(EDGE_POOL)


def to_bytes(arr):
    return np.ascontiguousarray(arr).view(np.uint8).reshape(-1).copy()


def look_at_cameras(n_cams, radius=1.0, seed=0, jitter=0.15):
    """c2w [n,3,4] (OpenGL convention: camera looks down -z) on a sphere, looking at the origin."""
    rng = np.random.default_rng(seed)
    k = np.arange(n_cams) + 0.5
    phi = np.arccos(1 - 2 * k / n_cams * 0.6 - 0.2)        # band around the equator
    theta = np.pi * (1 + 5 ** 0.5) * k
    pos = np.stack([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)], -1)
    pos = pos * radius * (1 + jitter * (rng.random((n_cams, 1)) - .5))
    c2w = np.zeros((n_cams, 3, 4), np.float32)
    for i in range(n_cams):
        z = pos[i] / np.linalg.norm(pos[i])               # camera z axis points away from the target
        up = np.array([0., 0., 1.]) if abs(z[2]) < 0.95 else np.array([0., 1., 0.])
        x = np.cross(up, z); x /= np.linalg.norm(x)
        y = np.cross(z, x)
        c2w[i, :, 0], c2w[i, :, 1], c2w[i, :, 2], c2w[i, :, 3] = x, y, z, pos[i]
    return c2w


def _distance_summary(dis):
    """PersSampler.cpp:16-25: exp(mean(log d | log d < 25 % quantile))."""
    if dis.size == 0:
        return 1e8
    ld = np.log(dis.astype(np.float32))
    thr = np.quantile(ld, 0.25).astype(np.float32)
    m = ld < thr
    return float(np.exp(ld.mean())) if m.sum() < 1e-3 else float(np.exp(ld[m].mean()))


class SyntheticScene:
    """Builds TreeNode / TransInfo / EdgePool blobs from inward-looking cameras."""

    def __init__(self, n_cams=24, focal_over_half_w=2.55, aspect=0.5625, near=0.05, far_bound=6.0,
                 bbox_levels=10, max_level=16, split_dist_thres=1.5, seed=0, max_nodes=200000):
        self.rng = np.random.default_rng(seed)
        self.c2w = look_at_cameras(n_cams, seed=seed)
        self.focal = float(focal_over_half_w)            # fx / cx
        self.half_w, self.half_h = 1.0, float(aspect)
        self.bound = np.array([near, far_bound], np.float32)
        self.max_level, self.split = int(max_level), float(split_dist_thres)
        self.max_nodes = max_nodes
        self.nodes, self.trans = [], []
        side = float(1 << (bbox_levels - 1))
        self._new_node(-1)
        self._build(0, 0, np.zeros(3, np.float32), side)
        self.edges = self._edge_pool()

    # -- visibility: a coarse ray grid per camera against the cube (PersSampler.cpp:27-66) ---------
    def _visible_cams(self, center, side):
        gw, gh = 16, 9
        j = (np.arange(gw) + .5) / gw * 2 - 1
        i = (np.arange(gh) + .5) / gh * 2 - 1
        jj, ii = np.meshgrid(j, i)
        cam = np.stack([jj.ravel() * self.half_w / self.focal, -ii.ravel() * self.half_h / self.focal,
                        -np.ones(gw * gh)], -1)                                     # [n_pix,3]
        d = np.einsum("cij,pj->cpi", self.c2w[:, :, :3], cam)                        # [n_cams,n_pix,3]
        o = self.c2w[:, None, :, 3]
        with np.errstate(divide="ignore", invalid="ignore"):
            a = ((center - side * .5)[None, None] - o) / d
            b = ((center + side * .5)[None, None] - o) / d
        a = np.nan_to_num(a, nan=0., posinf=1e6, neginf=-1e6); b = np.nan_to_num(b, nan=0., posinf=1e6, neginf=-1e6)
        far = np.minimum(np.maximum(a, b).min(-1), self.bound[1])
        near = np.maximum(np.minimum(a, b).max(-1), self.bound[0])
        return np.nonzero((far > near).sum(-1) > 0)[0]

    def _new_node(self, parent):
        n = np.zeros((), TREE_NODE)
        n["parent"], n["childs"], n["trans_idx"] = parent, -1, -1
        self.nodes.append(n)
        return len(self.nodes) - 1

    def _build(self, u, depth, center, side):
        nd = self.nodes[u]
        nd["center"], nd["side_len"], nd["is_leaf"], nd["trans_idx"] = center, side, 0, -1
        if depth > self.max_level or len(self.nodes) > self.max_nodes:
            nd["is_leaf"] = 1
            return
        visi = self._visible_cams(center, side)
        cam_dis = np.linalg.norm(self.c2w[:, :, 3] - center[None], axis=-1)
        dsum = _distance_summary(cam_dis[visi])
        if len(visi) >= N_PROS // 2 and dsum < side * self.split:
            for st in range(8):
                v = self._new_node(u)
                off = np.array([((st >> 2) & 1) - .5, ((st >> 1) & 1) - .5, (st & 1) - .5], np.float32)
                self.nodes[u]["childs"][st] = v
                self._build(v, depth + 1, (center + side * .5 * off).astype(np.float32), side * .5)
        elif len(visi) < N_PROS // 2:
            nd["is_leaf"] = 1
        else:
            nd["is_leaf"] = 1
            nd["trans_idx"] = len(self.trans)
            self.trans.append(self._construct_trans(center, side, visi))

    # -- PersOctree::ConstructTrans (PersSampler.cpp:439-620) in numpy ---------------------------------
    def _construct_trans(self, center, side, visi):
        n_virt = N_PROS // 2
        c2w = self.c2w[visi].astype(np.float64)
        cam_pos = c2w[:, :, 3]
        cam_axes = np.linalg.inv(c2w[:, :, :3])
        rel = cam_pos - center[None]
        dis = np.linalg.norm(rel, axis=-1)
        dsum = _distance_summary(dis)
        normed = rel / dis[:, None]
        pair = np.linalg.norm(normed[None] - normed[:, None], axis=-1)
        good = [int(self.rng.integers(len(visi)))]
        marks = np.zeros(len(visi), bool); marks[good[0]] = True
        while len(good) < min(n_virt, len(visi)):                                   # farthest-point pick
            dmin = np.where(marks[None, :], pair, 1e8).min(-1)
            dmin[marks] = -1
            c = int(np.argmax(dmin)); marks[c] = True; good.append(c)
        i = 0
        while len(good) < n_virt:
            good.append(good[i]); i += 1
        cam_scale = np.clip(dis / dsum, 1., 1e9)
        rel_clip = normed * np.clip(dis, dsum, 1e9)[:, None]
        g_pos = rel_clip[good] + center[None]
        g_axis = cam_axes[good].copy()
        g_scale = cam_scale[good]
        expect_z = rel_clip[good] / np.linalg.norm(rel_clip[good], axis=-1, keepdims=True)
        for k in range(n_virt):                                                     # rotate z axis onto expect_z
            fz, tz = g_axis[k, 2], expect_z[k]
            cr = np.cross(fz, tz); s = np.linalg.norm(cr); c = float(np.dot(fz, tz))
            ang = np.arcsin(min(s, 1.0))
            if c < 0: ang = np.pi - ang
            if s < 1e-12:
                R = np.eye(3)
            else:
                ax = cr / s
                K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
                R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
            g_axis[k] = g_axis[k] @ R.T
        x_axis = g_axis[:, 0] * self.focal * g_scale[:, None]
        y_axis = g_axis[:, 1] * self.focal * g_scale[:, None]
        z_axis = g_axis[:, 2]
        xa = np.concatenate([x_axis, y_axis], 0); za = np.concatenate([z_axis, z_axis], 0)
        wp = np.concatenate([g_pos, g_pos], 0)
        frame = np.zeros((N_PROS, 2, 4))
        frame[:, 0, :3], frame[:, 1, :3] = xa, za
        frame[:, 0, 3], frame[:, 1, 3] = -(xa * wp).sum(-1), -(za * wp).sum(-1)
        pts = (self.rng.random((4096, 3)) - .5) * side + center[None]
        tp = np.einsum("nrk,pk->pnr", frame[:, :, :3], pts) + frame[None, :, :, 3]   # [P,12,2]
        dv_da = 1. / tp[..., 1]
        dv_db = tp[..., 0] / -(tp[..., 1] ** 2)
        dv_dxyz = dv_da[..., None] * frame[None, :, 0, :3] + dv_db[..., None] * frame[None, :, 1, :3]
        vals = tp[..., 0] / tp[..., 1]
        mean = vals.mean(0, keepdims=True)
        cov = ((vals - mean)[:, :, None] * (vals - mean)[:, None, :]).mean(0)
        L, V = np.linalg.eigh(cov)
        V = V[:, np.argsort(-L)][:, :3].T                                            # [3,12]
        jac = np.einsum("rn,pnk->prk", V, dv_dxyz)
        jw2i = np.einsum("pnk,pkr->pnr", dv_dxyz, np.linalg.inv(jac))
        exp_step = 1. / np.abs(jw2i).max(1)
        V = V / exp_step.mean(0)[:, None]
        t = np.zeros((), TRANS_INFO)
        t["w2xz"], t["weight"], t["center"], t["dis_summary"] = frame.astype(np.float32), V.astype(np.float32), center, dsum
        return t

    # -- PersOctree::ConstructEdgePool (PersSampler.cpp:622-660): faces shared by two valid leaves ----
    def _edge_pool(self):
        nodes = self.nodes_array()
        valid = np.nonzero(nodes["trans_idx"] >= 0)[0]
        out = []
        for ai in range(len(valid)):
            for bi in range(ai + 1, len(valid)):
                a, b = valid[ai], valid[bi]
                u, v = (a, b) if nodes[a]["side_len"] <= nodes[b]["side_len"] else (b, a)
                lu = nodes[u]["side_len"] * .5
                cu, cv, sv = nodes[u]["center"], nodes[v]["center"], nodes[v]["side_len"]
                if np.abs(cu - cv).max() > lu + sv * .5 + 1e-3:
                    continue
                for ax in range(3):
                    for sgn in (1., -1.):
                        p = cu.copy(); p[ax] += sgn * lu
                        if np.abs((p - cv) / sv * 2).max() < 1 + 1e-4:
                            o0, o1 = [k for k in range(3) if k != ax]            # the two in-plane axes
                            d0, d1 = np.zeros(3, np.float32), np.zeros(3, np.float32)
                            d0[o0], d1[o1] = lu, lu
                            e = np.zeros((), EDGE_POOL)
                            e["t_idx_a"], e["t_idx_b"] = nodes[a]["trans_idx"], nodes[b]["trans_idx"]
                            e["center"], e["dir_0"], e["dir_1"] = p, d0, d1
                            out.append(e)
        return out

    def nodes_array(self):
        return np.array(self.nodes, dtype=TREE_NODE)

    def blobs(self):
        """(tree_nodes u8[N*64], pers_trans u8[V*544], edge_pool u8[E*64])."""
        nodes = to_bytes(self.nodes_array())
        trans = to_bytes(np.array(self.trans, dtype=TRANS_INFO)) if self.trans else np.zeros(0, np.uint8)
        edges = to_bytes(np.array(self.edges, dtype=EDGE_POOL)) if self.edges else np.zeros(0, np.uint8)
        return nodes, trans, edges

    def rays(self, n_rays, seed=1234):
        """Random pixels of random cameras -> (rays_o, rays_d) float32, un-normalised like the dataset."""
        rng = np.random.default_rng(seed)
        cam = rng.integers(len(self.c2w), size=n_rays)
        u = (rng.random(n_rays) * 2 - 1) * self.half_w / self.focal
        v = (rng.random(n_rays) * 2 - 1) * self.half_h / self.focal
        dcam = np.stack([u, -v, -np.ones(n_rays)], -1)
        d = np.einsum("nij,nj->ni", self.c2w[cam][:, :, :3], dcam)
        return self.c2w[cam][:, :, 3].astype(np.float32).copy(), d.astype(np.float32), cam.astype(np.int32)
