// B200Ops.h — C++/LibTorch host-side drop-ins for the reference's operator classes.
//
// Each class derives from the reference class it replaces (so octree construction, States/LoadStates,
// OptimParamGroups, Reset and every member the trainer touches are inherited unchanged) and overrides
// only the hot-path virtuals, forwarding them to the flat C ABI of libf2nerf_b200.so
// (include/f2nerf_b200.h).  Compiled INSIDE the reference tree (include path = <reference>/src); see
// INTEGRATION.md for the three factory lines and the CMake stanza.  No reference source is copied.
//
//   B200Sampler   : PersSampler      GetSamples / GetEdgeSamples / UpdateOctNodes  (PersSampler.h:75-96)
//   B200HashField : Hash3DAnchored   AnchoredQuery                                  (Hash3DAnchored.h:22-51)
//   B200Shader    : SHShader         Query                                          (SHShader.h)
#pragma once
#include <torch/torch.h>
#include "PtsSampler/PersSampler.h"
#include "Field/Hash3DAnchored.h"
#include "Shader/SHShader.h"

class B200Sampler : public PersSampler {
  using Tensor = torch::Tensor;
public:
  explicit B200Sampler(GlobalDataPool* global_data_pool) : PersSampler(global_data_pool) {}
  SampleResultFlex GetSamples(const Tensor& rays_o, const Tensor& rays_d, const Tensor& bounds) override;
  std::tuple<Tensor, Tensor> GetEdgeSamples(int n_pts) override;
  void UpdateOctNodes(const SampleResultFlex& sample_result, const Tensor& sampled_weights,
                      const Tensor& sampled_alpha) override;
};

class B200HashField : public Hash3DAnchored {
  using Tensor = torch::Tensor;
public:
  explicit B200HashField(GlobalDataPool* global_data_pool) : Hash3DAnchored(global_data_pool) {}
  Tensor AnchoredQuery(const Tensor& points, const Tensor& anchors) override;
};

class B200Shader : public SHShader {
  using Tensor = torch::Tensor;
public:
  explicit B200Shader(GlobalDataPool* global_data_pool) : SHShader(global_data_pool) {}
  Tensor Query(const Tensor& feats, const Tensor& dirs) override;
};
