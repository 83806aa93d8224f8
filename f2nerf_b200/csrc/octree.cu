// octree.cu — SURVEY §8(f) N2: octree maintenance on the device.
//
// Replaces PersOctree::ProcOctree (src/PtsSampler/PersSampler.cpp:120-330: three blob copies device->host, a
// sequential host pass over all nodes — prune dead leaves, collapse single-child chains, renumber, optionally split
// every visited leaf into 8 — and three copies back, a full pipeline stall every `compact_freq` iterations and at the
// subdivision milestones, PersSampler.cu:604-614) and PersOctree::MarkInvisibleNodes (PersSampler.cu:616-680).
// Everything stays in HBM; the host reads back 4 bytes (the new node count) to size its tensors.
//
// The reference's passes are sequential but their results are order-free, which is what makes them parallel:
//   prune      dead leaf -> cleared in its parent; childless internal node -> dead leaf; iterate (<= tree depth rounds)
//   collapse   R = non-root nodes with exactly one child; every other live node re-attaches to its nearest
//              ancestor outside R, in the child slot the chain hung from
//   renumber   compaction keeps index order (exclusive scan of the keep flags); subdivision renumbers in DFS
//              pre-order over child slots 0..7, the 8 new children right behind their parent: index(u) = sum over
//              the ancestors of (1 + sizes of the earlier siblings), with subtree sizes accumulated bottom-up
// so node numbering, parent / child links, centres and statistics are identical to the reference's.
#include "common.cuh"

namespace f2b {

constexpr int kInitNodeStat = 1000;            // INIT_NODE_STAT, PersSampler.h:10
constexpr int kPruneRounds = 32;               // >= tree depth (max_level 16 + root), PersSampler.cpp:359-421

__device__ __forceinline__ bool dead_leaf(const TreeNode& n) { return n.is_leaf_node && n.trans_idx < 0; }

// round (a): dead leaves unhook themselves from their parent (PersSampler.cpp:140-152)
__global__ void prune_unhook_kernel(TreeNode* __restrict__ nb, int n) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n) return;
  const TreeNode& me = nb[u];
  if (!dead_leaf(me) || me.parent < 0) return;
  TreeNode& p = nb[me.parent];
#pragma unroll
  for (int st = 0; st < 8; st++)
    if (p.childs[st] == u) p.childs[st] = -1;
}
// round (b): internal nodes that lost every child become (dead) leaves (PersSampler.cpp:154-173); the root never does
__global__ void prune_leafify_kernel(TreeNode* __restrict__ nb, int n) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u < 1 || u >= n) return;
  TreeNode& me = nb[u];
  bool any = false;
#pragma unroll
  for (int st = 0; st < 8; st++) any |= me.childs[st] >= 0;
  if (!any) me.is_leaf_node = 1;
}

// collapse: flag the single-child non-root internal nodes (PersSampler.cpp:181-214)
__global__ void collapse_flag_kernel(const TreeNode* __restrict__ nb, int n, int* __restrict__ removed) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n) return;
  int cnt = 0;
#pragma unroll
  for (int st = 0; st < 8; st++) cnt += nb[u].childs[st] >= 0;
  removed[u] = (nb[u].parent >= 0 && cnt == 1) ? 1 : 0;
}
__global__ void collapse_reparent_kernel(TreeNode* __restrict__ nb, int n, const int* __restrict__ removed) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n || removed[u] || dead_leaf(nb[u])) return;
  int prev = u, v = nb[u].parent;
  while (v >= 0 && removed[v]) { prev = v; v = nb[v].parent; }
  if (prev == u) return;                        // parent is not collapsed (or u is the root)
  // v >= 0: removed nodes are never the root, so the walk ends on a kept ancestor
#pragma unroll
  for (int st = 0; st < 8; st++)
    if (nb[v].childs[st] == prev) nb[v].childs[st] = u;
  nb[u].parent = v;
}

// keep flags (PersSampler.cpp:216-224) + per-node subdivision decision and own contribution to subtree sizes
__global__ void keep_flags_kernel(const TreeNode* __restrict__ nb, int n, const int* __restrict__ removed,
                                  const int* __restrict__ visit_cnt, int subdivide, int brute_force,
                                  int* __restrict__ keep, int* __restrict__ sub, int* __restrict__ size) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n) return;
  const bool k = !removed[u] && !dead_leaf(nb[u]);
  keep[u] = k ? 1 : 0;
  const bool s = k && subdivide && nb[u].is_leaf_node && (brute_force || visit_cnt[u] > 4);
  sub[u] = s ? 1 : 0;
  size[u] = 0;
}
__global__ void subtree_size_kernel(const TreeNode* __restrict__ nb, int n, const int* __restrict__ keep,
                                    const int* __restrict__ sub, int* __restrict__ size) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n || !keep[u]) return;
  const int c = sub[u] ? 9 : 1;
  for (int a = u; a >= 0; a = nb[a].parent) atomicAdd(size + a, c);
}
// DFS pre-order index of a kept node in the subdivided tree
__global__ void preorder_kernel(const TreeNode* __restrict__ nb, int n, const int* __restrict__ keep,
                                const int* __restrict__ size, int* __restrict__ order) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n || !keep[u]) return;
  int idx = 0;
  for (int cur = u, p = nb[u].parent; p >= 0; cur = p, p = nb[p].parent) {
    idx += 1;
    for (int st = 0; st < 8; st++) {
      const int c = nb[p].childs[st];
      if (c == cur) break;
      if (c >= 0) idx += size[c];
    }
  }
  order[u] = idx;
}

// write the new tree (PersSampler.cpp:226-250 without / :252-313 with subdivision)
__global__ void emit_kernel(const TreeNode* __restrict__ nb, int n, const int* __restrict__ keep,
                            const int* __restrict__ order, const int* __restrict__ sub,
                            const int* __restrict__ wstat, const int* __restrict__ astat,
                            TreeNode* __restrict__ out, int* __restrict__ wstat_o, int* __restrict__ astat_o) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n || !keep[u]) return;
  const int o = order[u];
  TreeNode me = nb[u];
  if (me.parent >= 0) me.parent = order[me.parent];
#pragma unroll
  for (int st = 0; st < 8; st++)
    if (me.childs[st] >= 0) me.childs[st] = order[me.childs[st]];
  if (sub[u]) {
    const float half = fmul(me.side_len, .5f);
    for (int st = 0; st < 8; st++) {
      TreeNode ch;
      ch.center[0] = fadd(me.center[0], fmul(half, float((st >> 2) & 1) - .5f));
      ch.center[1] = fadd(me.center[1], fmul(half, float((st >> 1) & 1) - .5f));
      ch.center[2] = fadd(me.center[2], fmul(half, float(st & 1) - .5f));
      ch.side_len = half;
      ch.parent = o;
#pragma unroll
      for (int k = 0; k < 8; k++) ch.childs[k] = -1;
      ch.is_leaf_node = 1;
      ch._pad0[0] = ch._pad0[1] = ch._pad0[2] = 0;
      ch.trans_idx = me.trans_idx;
      ch._pad = 0;
      out[o + 1 + st] = ch;
      wstat_o[o + 1 + st] = wstat[u];
      astat_o[o + 1 + st] = astat[u];
      me.childs[st] = o + 1 + st;
    }
    me.is_leaf_node = 0;
    me.trans_idx = -1;
    wstat_o[o] = kInitNodeStat;
    astat_o[o] = kInitNodeStat;
  } else {
    wstat_o[o] = wstat[u];
    astat_o[o] = astat[u];
  }
  out[o] = me;
}

__global__ void finish_kernel(const int* __restrict__ keep_bounds_total, const int* __restrict__ size, int subdivide,
                              int* __restrict__ n_out) {
  n_out[0] = subdivide ? size[0] : keep_bounds_total[0];
}

// exclusive scan of the keep flags (single block; n is 1e3..1e5) -> order[u] for kept nodes, total
__global__ void __launch_bounds__(1024) keep_scan_kernel(const int* __restrict__ keep, int n, int* __restrict__ order,
                                                         int* __restrict__ total) {
  __shared__ int s_warp[32];
  const int tid = threadIdx.x;
  const int per = (n + 1023) / 1024;
  const int beg = min(tid * per, n), end = min(beg + per, n);
  int local = 0;
  for (int i = beg; i < end; i++) local += keep[i];
  int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if ((tid & 31) >= o) incl += v;
  }
  if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    int w = s_warp[tid];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, w, o);
      if (tid >= o) w += v;
    }
    s_warp[tid] = w;
  }
  __syncthreads();
  int run = incl - local + ((tid >> 5) ? s_warp[(tid >> 5) - 1] : 0);
  for (int i = beg; i < end; i++) {
    if (keep[i]) order[i] = run;
    run += keep[i];
  }
  if (tid == 1023) total[0] = s_warp[31];
}

// MarkInvisibleNodesKernel + CheckVisible (PersSampler.cu:616-666)
__global__ void mark_invisible_kernel(int n_nodes, int n_cams, TreeNode* __restrict__ nodes, const float* __restrict__ intri,
                                      const float* __restrict__ w2c, const float* __restrict__ bound) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_nodes) return;
  const float cx0 = nodes[u].center[0], cy0 = nodes[u].center[1], cz0 = nodes[u].center[2];
  const float radius = (float)((double)nodes[u].side_len * 0.707);
  int visible = 0;
  for (int c = 0; c < n_cams && !visible; c++) {
    const float* W = w2c + size_t(c) * 12;
    const float* K = intri + size_t(c) * 9;
    const float x = fadd(ffma(W[0], cx0, ffma(W[1], cy0, fmul(W[2], cz0))), W[3]);
    const float y = fadd(ffma(W[4], cx0, ffma(W[5], cy0, fmul(W[6], cz0))), W[7]);
    const float z = fadd(ffma(W[8], cx0, ffma(W[9], cy0, fmul(W[10], cz0))), W[11]);
    const float nz = -z;
    if (nz < fsub(bound[2 * c], radius) || nz > fadd(bound[2 * c + 1], radius)) continue;
    if (fsqrt(ffma(x, x, ffma(y, y, fmul(z, z)))) < radius) { visible = 1; break; }
    const float kcx = K[2], kcy = K[5], fx = K[0], fy = K[4];
    const float bx = fmul(fdiv(radius, nz), fx), by = fmul(fdiv(radius, nz), fy);
    const float px = fmul(fdiv(x, nz), fx), py = fmul(fdiv(y, nz), fy);
    if (fadd(px, bx) < -kcx || px > fadd(kcx, bx) || fadd(py, by) < -kcy || py > fadd(kcy, by)) continue;
    visible = 1;
  }
  if (!visible) nodes[u].trans_idx = -1;
}

}  // namespace f2b

using namespace f2b;

extern "C" int f2b_octree_proc(const void* tree_nodes, const int* weight_stats, const int* alpha_stats, const int* visit_cnt,
                               int n_nodes, int subdivide, int brute_force, void* work_nodes /* [n_nodes] */,
                               int* work_i32 /* [5 * n_nodes + 2] */, void* nodes_out /* capacity 9 * n_nodes */,
                               int* weight_stats_out, int* alpha_stats_out, int* n_nodes_out /* device [1] */, void* stream) {
  F2B_REQUIRE(n_nodes > 0, "f2b_octree_proc: empty tree");
  F2B_REQUIRE(tree_nodes && weight_stats && alpha_stats && visit_cnt && work_nodes && work_i32 && nodes_out && weight_stats_out &&
                  alpha_stats_out && n_nodes_out, "f2b_octree_proc: null pointer");
  cudaStream_t st = as_stream(stream);
  TreeNode* nb = (TreeNode*)work_nodes;
  int* removed = work_i32;
  int* keep = work_i32 + n_nodes;
  int* sub = work_i32 + 2 * size_t(n_nodes);
  int* size = work_i32 + 3 * size_t(n_nodes);
  int* order = work_i32 + 4 * size_t(n_nodes);
  int* total = work_i32 + 5 * size_t(n_nodes);
  const int g = div_up(n_nodes, 256);
  cudaMemcpyAsync(nb, tree_nodes, size_t(n_nodes) * sizeof(TreeNode), cudaMemcpyDeviceToDevice, st);
  for (int r = 0; r < kPruneRounds; r++) {
    prune_unhook_kernel<<<g, 256, 0, st>>>(nb, n_nodes);
    prune_leafify_kernel<<<g, 256, 0, st>>>(nb, n_nodes);
  }
  prune_unhook_kernel<<<g, 256, 0, st>>>(nb, n_nodes);
  collapse_flag_kernel<<<g, 256, 0, st>>>(nb, n_nodes, removed);
  collapse_reparent_kernel<<<g, 256, 0, st>>>(nb, n_nodes, removed);
  keep_flags_kernel<<<g, 256, 0, st>>>(nb, n_nodes, removed, visit_cnt, subdivide, brute_force, keep, sub, size);
  if (subdivide) {
    subtree_size_kernel<<<g, 256, 0, st>>>(nb, n_nodes, keep, sub, size);
    preorder_kernel<<<g, 256, 0, st>>>(nb, n_nodes, keep, size, order);
    cudaMemsetAsync(total, 0, sizeof(int), st);
  } else {
    keep_scan_kernel<<<1, 1024, 0, st>>>(keep, n_nodes, order, total);
  }
  emit_kernel<<<g, 256, 0, st>>>(nb, n_nodes, keep, order, sub, weight_stats, alpha_stats, (TreeNode*)nodes_out, weight_stats_out,
                                 alpha_stats_out);
  finish_kernel<<<1, 1, 0, st>>>(total, size, subdivide, n_nodes_out);
  return check_launch("f2b_octree_proc");
}

extern "C" int f2b_octree_mark_invisible(void* tree_nodes, int n_nodes, const float* intri, const float* w2c,
                                         const float* bounds, int n_cams, void* stream) {
  if (n_nodes <= 0) return F2B_OK;
  F2B_REQUIRE(tree_nodes && intri && w2c && bounds && n_cams >= 0, "f2b_octree_mark_invisible: bad argument");
  mark_invisible_kernel<<<div_up(n_nodes, 128), 128, 0, as_stream(stream)>>>(n_nodes, n_cams, (TreeNode*)tree_nodes, intri, w2c, bounds);
  return check_launch("f2b_octree_mark_invisible");
}
