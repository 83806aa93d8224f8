// ref_driver.cpp — harness around the UNMODIFIED reference operators (Totoro97/f2-nerf @98f0daa),
// linked from the reference's own sources by oracle/Makefile.ref into oracle/_ref/ref_driver.
//
// TEST INFRASTRUCTURE ONLY (see oracle/f2_oracle.c header).  It constructs the reference's
// GlobalDataPool / Dataset / Renderer exactly like ExpRunner::ExpRunner does (src/ExpRunner.cpp:17-63),
// then (1) dumps the octree / warp / hash blobs, (2) runs PersSampler::GetSamples,
// Hash3DAnchored::AnchoredQuery, SHShader::Query and Renderer::Render (+ backward) on fixed seeded
// inputs and dumps every boundary tensor as .npy, (3) times Render + backward with CUDA events.
// The dumps pin oracle/f2_oracle.c and the CUDA kernels (tests/golden/, tests/test_ref_parity.py).
//
//   ref_driver <runtime_config.yaml> <out_dir> <n_rays> [n_time_iters=0] [dump_full_grads=0]
//   ref_driver --train <runtime_config.yaml>      what the reference's main.cpp does: ExpRunner(conf).Execute() — the
//                                                 unmodified trainer (train to end_iter, then TestImages -> mean PSNR)
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <iostream>
#include <experimental/filesystem>
#include <torch/torch.h>
#include <cuda_runtime.h>
#ifdef F2B_WITH_SHIM
#include "../f2nerf_b200/shim/B200Ops.h"
#include "../f2nerf_b200/shim/B200Renderer.h"
#endif
#include "ExpRunner.h"
#include "Common.h"
#include "Utils/GlobalDataPool.h"
#include "Utils/cnpy.h"
#include "Utils/CustomOps/CustomOps.h"
#include "Dataset/Dataset.h"
#include "Renderer/Renderer.h"
#include "PtsSampler/PersSampler.h"
#include "Field/Hash3DAnchored.h"
#include "Shader/SHShader.h"

namespace fs = std::experimental::filesystem::v1;
using Tensor = torch::Tensor;

static std::string g_out;

static void dump(const std::string& name, const Tensor& t_in) {
  Tensor t = t_in.detach().to(torch::kCPU).contiguous();
  std::vector<size_t> shape;
  for (auto s : t.sizes()) shape.push_back((size_t)s);
  if (shape.empty()) shape.push_back(1);
  const std::string path = g_out + "/" + name + ".npy";
  switch (t.scalar_type()) {
    case torch::kFloat32: cnpy::npy_save(path, t.data_ptr<float>(), shape); break;
    case torch::kInt32: cnpy::npy_save(path, t.data_ptr<int>(), shape); break;
    case torch::kInt64: cnpy::npy_save(path, t.data_ptr<int64_t>(), shape); break;
    case torch::kUInt8: cnpy::npy_save(path, t.data_ptr<uint8_t>(), shape); break;
    case torch::kBool: { Tensor u = t.to(torch::kUInt8); cnpy::npy_save(path, u.data_ptr<uint8_t>(), shape); break; }
    default: std::cerr << "dump: unsupported dtype for " << name << std::endl; std::exit(2);
  }
}
static void dump_scalar(const std::string& name, float v) { dump(name, torch::full({1}, v, CPUFloat)); }

int main(int argc, char** argv) {
  if (argc >= 3 && std::string(argv[1]) == "--train") {     // main.cpp:6-13 verbatim in behaviour
    torch::manual_seed(2022);
    auto exp_runner = std::make_unique<ExpRunner>(argv[2]);
    exp_runner->Execute();
    return 0;
  }
  if (argc < 4) { std::fprintf(stderr, "usage: ref_driver <config.yaml> <out_dir> <n_rays> [n_time_iters] [dump_full_grads]\n"); return 1; }
  const std::string conf_path = argv[1];
  g_out = argv[2];
  const int n_rays = std::atoi(argv[3]);
  const int n_time = argc > 4 ? std::atoi(argv[4]) : 0;
  const int full_grads = argc > 5 ? std::atoi(argv[5]) : 0;
  fs::create_directories(g_out);

  torch::manual_seed(2022);   // main.cpp:8
  auto gdp = std::make_unique<GlobalDataPool>(conf_path);
  gdp->base_exp_dir_ = gdp->config_["base_exp_dir"].as<std::string>();
  fs::create_directories(gdp->base_exp_dir_);
  gdp->learning_rate_ = 1e-2f;
  auto dataset = std::make_unique<Dataset>(gdp.get());
  auto renderer = std::make_unique<Renderer>(gdp.get(), dataset->n_images_);
  auto* sampler = dynamic_cast<PersSampler*>(renderer->pts_sampler_.get());
  auto* field = dynamic_cast<Hash3DAnchored*>(renderer->scene_field_.get());
  auto* shader = dynamic_cast<SHShader*>(renderer->shader_.get());
  CHECK(sampler && field && shader);

  // ---- deterministic, non-trivial parameters (reproducible from the CPU generator in Python) ----
  {
    auto g = at::detail::createCPUGenerator(1234);
    Tensor feat = (torch::rand({field->pool_size_, 2}, g, CPUFloat) * 2.f - 1.f);      // U(-1,1)
    field->feat_pool_.data().copy_(feat.to(torch::kCUDA));
    field->mlp_->params_.data().mul_(4.f);          // wider density range so early stop triggers
    renderer->app_emb_.data().copy_((torch::rand({dataset->n_images_, 16}, g, CPUFloat) * .2f - .1f).to(torch::kCUDA));
  }

#ifdef F2B_WITH_SHIM
  // ---- ref_driver_b200 only.  Default (F2B_SHIM unset or "fused"): the body of Renderer::Render is the fused B200 host
  // (f2nerf_b200/shim/B200Renderer.cpp) working on the reference's own operator objects.  F2B_SHIM=ops: the operator-level
  // drop-in — the reference's OWN Renderer::Render / autograd / loss drive the B200 subclasses: same octree
  // object, same parameters.  Either way every dump below is directly comparable with the pure-reference run of
  // oracle/_ref/ref_driver.  Never built into ref_driver (the reference arm stays unmodified).
  const char* shim_env = std::getenv("F2B_SHIM");
  const bool shim_ops = shim_env && std::string(shim_env) == "ops";
  if (shim_ops) {
    f2b_render_use_reference(true);
    auto b_sampler = std::make_unique<B200Sampler>(gdp.get());
    b_sampler->pers_octree_ = std::move(sampler->pers_octree_);          // the very same octree / warps / edge pool
    auto b_field = std::make_unique<B200HashField>(gdp.get());
    b_field->LoadStates(field->States(), 0);
    auto b_shader = std::make_unique<B200Shader>(gdp.get());
    b_shader->LoadStates(shader->States(), 0);
    renderer->pts_sampler_ = std::move(b_sampler);
    renderer->scene_field_ = std::move(b_field);
    renderer->shader_ = std::move(b_shader);
    sampler = dynamic_cast<PersSampler*>(renderer->pts_sampler_.get());
    field = dynamic_cast<Hash3DAnchored*>(renderer->scene_field_.get());
    shader = dynamic_cast<SHShader*>(renderer->shader_.get());
    CHECK(sampler && field && shader);
    std::printf("ref_driver_b200: PersSampler / Hash3DAnchored / SHShader replaced by the B200 subclasses\n");
  } else {
    f2b_render_keep_samples(true);                                       // the dumps read renderer->sample_result_
    std::printf("ref_driver_b200: Renderer::Render replaced by the fused B200 host\n");
  }
#endif

  // ---- scene blobs & scalars --------------------------------------------------------------------
  dump("tree_nodes", sampler->pers_octree_->tree_nodes_gpu_);
  dump("pers_trans", sampler->pers_octree_->pers_trans_gpu_);
  dump("edge_pool", sampler->pers_octree_->edge_pool_gpu_);
  dump("search_order", sampler->pers_octree_->node_search_order_);
  dump("prim_pool", field->prim_pool_);
  dump("bias_pool", field->bias_pool_);
  dump("field_mlp_params", field->mlp_->params_);
  dump("shader_mlp_params", shader->mlp_->params_);
  dump("app_emb", renderer->app_emb_);
  {
    Tensor sc = torch::zeros({8}, CPUFloat);
    sc[0] = sampler->global_near_; sc[1] = sampler->sample_l_; sc[2] = sampler->scale_by_dis_ ? 1.f : 0.f;
    sc[3] = float(sampler->max_oct_intersect_per_ray_); sc[4] = float(field->n_volumes_);
    sc[5] = float(field->pool_size_); sc[6] = float(dataset->n_images_); sc[7] = float(n_rays);
    dump("scalars", sc);
  }

  // ---- rays ---------------------------------------------------------------------------------------
  torch::manual_seed(2023);
  auto [rays, gt_colors, emb_idx] = dataset->RandRaysData(n_rays, DATA_TRAIN_SET);
  Tensor rays_o = rays.origins.contiguous(), rays_d = rays.dirs.contiguous(), bounds = rays.bounds.contiguous();
  dump("rays_o", rays_o); dump("rays_d", rays_d); dump("gt_colors", gt_colors); dump("emb_idx", emb_idx);
  dump("rays_d_normed", (rays_d / torch::linalg_norm(rays_d, 2, -1, true)).contiguous());
  {   // ray generation: the camera tables and a replay of RandRaysData's CPU draws (Dataset.cpp:287-289)
    dump("ds_poses", dataset->poses_.reshape({-1, 12})); dump("ds_intri", dataset->intri_.reshape({-1, 9}));
    dump("ds_dist_params", dataset->dist_params_); dump("ds_bounds", dataset->bounds_); dump("ray_bounds", bounds);
    Tensor hw = torch::zeros({2}, CUDAFloat); hw[0] = float(dataset->height_); hw[1] = float(dataset->width_);
    dump("ds_hw", hw);
    Tensor tset = torch::from_blob(dataset->train_set_.data(), {int(dataset->train_set_.size())}, CPUInt).clone().to(torch::kCUDA);
    dump("ds_train_set", tset);
    torch::manual_seed(2023);
    Tensor cam = torch::randint(int(dataset->train_set_.size()), {n_rays}, CPULong);
    Tensor ri = torch::randint(0, dataset->height_, n_rays, CPULong);
    Tensor rj = torch::randint(0, dataset->width_, n_rays, CPULong);
    dump("ray_ij", torch::stack({ri, rj}, -1).to(torch::kInt32).to(torch::kCUDA).contiguous());
    dump("ray_cam_draw", cam.to(torch::kInt32).to(torch::kCUDA).contiguous());
  }

  // ---- edge samples: the operator alone, with its RNG draws replayed next to it --------------------
  {
    torch::manual_seed(4242);
    auto [edge_pts, edge_anchors] = sampler->GetEdgeSamples(8192);
    dump("edge_pts", edge_pts); dump("edge_anchors", edge_anchors);
    torch::manual_seed(4242);
    int n_edges = sampler->pers_octree_->edge_pool_.size();
    Tensor edge_idx = torch::randint(0, n_edges, {8192}, CUDAInt);                                 // PersSampler.cu:456
    Tensor edge_coord = torch::rand({8192, 2}, CUDAFloat) * 2.f - 1.f;                             // :457
    dump("edge_idx", edge_idx); dump("edge_coord", edge_coord);
  }

  // ---- sampler, VALIDATE mode (noise == 1) --------------------------------------------------------
  gdp->iter_step_ = 1;
  gdp->ray_march_fineness_ = 1.f;
  gdp->gradient_scaling_progress_ = 0.25f;   // exercise GradientScaling
  gdp->mode_ = RunningMode::VALIDATE;
  {
    torch::NoGradGuard ng;
    auto s = sampler->GetSamples(rays_o, rays_d, bounds);
    dump("val_pts", s.pts); dump("val_dirs", s.dirs); dump("val_dt", s.dt); dump("val_t", s.t);
    dump("val_anchors", s.anchors.index({Slc(), Slc(0, 2)}).contiguous());
    dump("val_bounds", s.pts_idx_bounds); dump("val_first_oct_dis", s.first_oct_dis);
    Tensor feat = field->AnchoredQuery(s.pts, s.anchors.index({"...", 0}).contiguous());
    dump("val_scene_feat", feat);
    Tensor shading = torch::cat({torch::ones_like(feat.index({Slc(), Slc(0, 1)})), feat.index({Slc(), Slc(1, None)})}, 1);
    dump("val_rgb", shader->Query(shading, s.dirs));
    dump("val_sh", shader->SHEncode(s.dirs));
    auto r = renderer->Render(rays_o, rays_d, bounds, Tensor());
    dump("val_colors", r.colors); dump("val_disparity", r.disparity); dump("val_depth", r.depth);
    dump("val_weights", r.weights); dump("val_idx_start_end", r.idx_start_end);
  }

  // ---- TRAIN mode: seeded RNG, replayed once to expose the internal draws -------------------------
  gdp->mode_ = RunningMode::TRAIN;
  const int64_t seed = 777;
  Tensor stats_w0 = sampler->pers_octree_->tree_weight_stats_.clone();
  {
    torch::manual_seed(seed);
    auto r = renderer->Render(rays_o, rays_d, bounds, emb_idx);
    dump("train_colors", r.colors); dump("train_disparity", r.disparity); dump("train_depth", r.depth);
    dump("train_weights", r.weights); dump("train_idx_start_end", r.idx_start_end);
    dump("train_first_oct_dis", r.first_oct_dis); dump("train_edge_feats", r.edge_feats);
    const auto& s = renderer->sample_result_;
    dump("train_pts", s.pts); dump("train_dt", s.dt); dump("train_t", s.t);
    dump("train_anchors", s.anchors.index({Slc(), Slc(0, 2)}).contiguous()); dump("train_bounds", s.pts_idx_bounds);
    dump("train_tree_nodes_after", sampler->pers_octree_->tree_nodes_gpu_);
    dump("train_weight_stats_after", sampler->pers_octree_->tree_weight_stats_);
    dump("train_alpha_stats_after", sampler->pers_octree_->tree_alpha_stats_);
    dump("train_visit_cnt_after", sampler->pers_octree_->tree_visit_cnt_);

    // losses as ExpRunner::Train (src/ExpRunner.cpp:94-118), fixed weights
    Tensor color_loss = torch::sqrt((r.colors - gt_colors).square() + 1e-4f).mean();
    Tensor disparity_loss = r.disparity.square().mean();
    Tensor tv_loss = (r.edge_feats.index({Slc(), 0}) - r.edge_feats.index({Slc(), 1})).square().mean();
    Tensor var_loss = (CustomOps::WeightVar(r.weights, r.idx_start_end) + 1e-2).sqrt().mean();
    Tensor loss = color_loss + var_loss * 1e-2f + disparity_loss * 1e-2f + tv_loss * 1e-1f;
    dump("train_loss", loss.reshape({1}));
    loss.backward();
    dump("grad_field_mlp", field->mlp_->params_.grad());
    dump("grad_shader_mlp", shader->mlp_->params_.grad());
    if (renderer->app_emb_.grad().defined()) dump("grad_app_emb", renderer->app_emb_.grad());   // undefined when use_app_emb is false
    Tensor g = field->feat_pool_.grad().reshape({-1});
    if (full_grads) dump("grad_feat_pool", g);
    Tensor nz = torch::nonzero(g).reshape({-1});
    dump("grad_feat_pool_nz_idx", nz.to(torch::kInt32));
    dump("grad_feat_pool_nz_val", g.index({nz}));
    dump_scalar("backward_nan", gdp->backward_nan_ ? 1.f : 0.f);
    // ---- the optimizer step the trainer takes next (ExpRunner.cpp:54,136): torch::optim::Adam over the field's groups
    if (full_grads) {
      dump("feat_pool_before_adam", field->feat_pool_.detach().reshape({-1}));
      dump("field_mlp_before_adam", field->mlp_->params_.detach().reshape({-1}));
      torch::optim::Adam opt(field->OptimParamGroups());
      for (int it = 0; it < 2; it++) opt.step();                 // same gradient twice: exercises step 1 and step 2 bias corrections
      dump("feat_pool_after_adam", field->feat_pool_.detach().reshape({-1}));
      dump("field_mlp_after_adam", field->mlp_->params_.detach().reshape({-1}));
      dump_scalar("adam_lr", gdp->learning_rate_);
    }
  }
  {   // replay of the RNG draws Render made, in its order.  Every TCNNWP forward allocates its output with
      // torch::rand (TCNNWP.cpp:143), so the early-stop MLP call sits between the background and the edge draws.
    const int64_t n_all = renderer->sample_result_.pts.size(0);
    torch::manual_seed(seed);
    Tensor noise = ((torch::rand({1024 + n_rays + 10}, CUDAFloat) - .5f) + 1.f).contiguous();   // PersSampler.cu:377
    noise.mul_(gdp->ray_march_fineness_);
    Tensor bg = torch::rand({n_rays, 3}, CUDAFloat);                                               // Renderer.cpp:73
    Tensor burn = torch::rand({(n_all + 127) / 128 * 128, 16}, torch::TensorOptions().dtype(torch::kFloat16).device(torch::kCUDA));
    int n_edges = sampler->pers_octree_->edge_pool_.size();
    Tensor edge_idx = torch::randint(0, n_edges, {8192}, CUDAInt);                                 // PersSampler.cu:456
    Tensor edge_coord = torch::rand({8192, 2}, CUDAFloat) * 2.f - 1.f;                             // :457
    dump("train_noise", noise); dump("train_bg", bg); dump("train_edge_idx", edge_idx); dump("train_edge_coord", edge_coord);
  }

  // ---- timing: Render + backward, CUDA events on the default stream ------------------------------
  if (n_time > 0) {
#ifdef F2B_WITH_SHIM
    f2b_render_keep_samples(false);
#endif
    std::vector<float> ms_fwd, ms_all;
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    double samples = 0, kept = 0;
    for (int it = 0; it < n_time + 5; it++) {
      gdp->iter_step_ = 1 + it * 2 + 1;      // never a multiple of compact_freq / milestone
      field->feat_pool_.mutable_grad() = Tensor(); field->mlp_->params_.mutable_grad() = Tensor();
      shader->mlp_->params_.mutable_grad() = Tensor(); renderer->app_emb_.mutable_grad() = Tensor();
      torch::cuda::synchronize();
      cudaEventRecord(e0, 0);
      auto r = renderer->Render(rays_o, rays_d, bounds, emb_idx);
      cudaEventRecord(e1, 0);
      Tensor color_loss = torch::sqrt((r.colors - gt_colors).square() + 1e-4f).mean();
      Tensor tv_loss = (r.edge_feats.index({Slc(), 0}) - r.edge_feats.index({Slc(), 1})).square().mean();
      Tensor var_loss = (CustomOps::WeightVar(r.weights, r.idx_start_end) + 1e-2).sqrt().mean();
      Tensor loss = color_loss + var_loss * 1e-2f + tv_loss * 1e-1f;
      loss.backward();
      cudaEventRecord(e2, 0);
      torch::cuda::synchronize();
      float a, b; cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e0, e2);
      if (it >= 5) { ms_fwd.push_back(a); ms_all.push_back(b); samples += renderer->sample_result_.pts.size(0); kept += r.weights.size(0); }
    }
    std::sort(ms_fwd.begin(), ms_fwd.end()); std::sort(ms_all.begin(), ms_all.end());
    const float mf = ms_fwd[ms_fwd.size() / 2], ma = ms_all[ms_all.size() / 2];
    // forward only, as ExpRunner::RenderWholeImage drives it (VALIDATE mode, NoGradGuard; ExpRunner.cpp:257-293)
    std::vector<float> ms_val;
    {
      torch::NoGradGuard ng;
      gdp->mode_ = RunningMode::VALIDATE;
      for (int it = 0; it < n_time + 5; it++) {
        torch::cuda::synchronize();
        cudaEventRecord(e0, 0);
        auto r = renderer->Render(rays_o, rays_d, bounds, Tensor());
        cudaEventRecord(e1, 0);
        torch::cuda::synchronize();
        float a; cudaEventElapsedTime(&a, e0, e1);
        if (it >= 5) ms_val.push_back(a);
      }
      gdp->mode_ = RunningMode::TRAIN;
    }
    std::sort(ms_val.begin(), ms_val.end());
    const float mv = ms_val[ms_val.size() / 2];
    std::printf("{\"ref_timing\": {\"n_rays\": %d, \"iters\": %d, \"ms_fwd_median\": %.4f, \"ms_fwd_bwd_median\": %.4f, "
                "\"samples_per_ray\": %.2f, \"kept_per_ray\": %.2f, \"rays_per_s\": %.1f, \"ms_validate_median\": %.4f}}\n",
                n_rays, n_time, mf, ma, samples / n_time / n_rays, kept / n_time / n_rays, n_rays / (ma * 1e-3), mv);
    FILE* f = std::fopen((g_out + "/ref_timing.json").c_str(), "w");
    std::fprintf(f, "{\"n_rays\": %d, \"iters\": %d, \"ms_fwd_median\": %.4f, \"ms_fwd_bwd_median\": %.4f, "
                 "\"samples_per_ray\": %.2f, \"kept_per_ray\": %.2f, \"rays_per_s\": %.1f, \"ms_validate_median\": %.4f}\n",
                 n_rays, n_time, mf, ma, samples / n_time / n_rays, kept / n_time / n_rays, n_rays / (ma * 1e-3), mv);
    std::fclose(f);
  }
  // ---- octree maintenance: ProcOctree / MarkInvisibleNodes on a deterministically damaged tree -------
  if (full_grads) {
    auto& oct = *sampler->pers_octree_;
    const int n = (int)oct.tree_nodes_.size();
    Tensor nodes_cpu = oct.tree_nodes_gpu_.to(torch::kCPU).clone();
    TreeNode* tn = reinterpret_cast<TreeNode*>(nodes_cpu.data_ptr());
    int k = 0;
    for (int u = 0; u < n; u++)
      if (tn[u].is_leaf_node && tn[u].trans_idx >= 0 && (k++ % 5 == 0)) tn[u].trans_idx = -1;     // as MarkInvalidNodes would
    oct.tree_nodes_gpu_ = nodes_cpu.to(torch::kCUDA).contiguous();
    oct.tree_visit_cnt_ = ((torch::arange(n, CUDAInt) * 7) % 11).to(torch::kInt32).contiguous();
    dump("oct_nodes_in", oct.tree_nodes_gpu_); dump("oct_w_in", oct.tree_weight_stats_); dump("oct_a_in", oct.tree_alpha_stats_);
    dump("oct_visit_in", oct.tree_visit_cnt_);
    dump("oct_intri", oct.intri_.reshape({-1, 9})); dump("oct_w2c", oct.w2c_.reshape({-1, 12})); dump("oct_bound", oct.bound_);
    oct.ProcOctree(true, true, false);
    dump("oct_nodes_sub", oct.tree_nodes_gpu_); dump("oct_w_sub", oct.tree_weight_stats_); dump("oct_a_sub", oct.tree_alpha_stats_);
    oct.MarkInvisibleNodes();
    dump("oct_nodes_invis", oct.tree_nodes_gpu_);
    oct.ProcOctree(true, false, false);
    dump("oct_nodes_final", oct.tree_nodes_gpu_); dump("oct_w_final", oct.tree_weight_stats_); dump("oct_a_final", oct.tree_alpha_stats_);
  }
  std::printf("ref_driver: done, dumps in %s\n", g_out.c_str());
  return 0;
}
