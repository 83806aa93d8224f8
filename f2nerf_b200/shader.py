"""SHShader — host-side mirror of ``src/Shader/SHShader.{h,cpp,cu}`` over the C ABI."""
import torch

from . import ops
from .field import TCNNWP


class SHShader:
    def __init__(self, global_data_pool, d_in=32, d_out=3, d_hidden=64, n_hiddens=2, degree=4, device="cuda"):
        if degree != 4 or d_in != 32:
            raise NotImplementedError("SHShader(b200): degree 4 / d_in 32 (confs/shader/sh_shader.yaml) is what is built")
        self.global_data_pool_ = global_data_pool
        self.d_in_, self.d_out_, self.degree_, self.d_hidden_, self.n_hiddens_ = d_in, d_out, degree, d_hidden, n_hiddens
        self.mlp_ = TCNNWP(global_data_pool, d_in, d_out, d_hidden, n_hiddens, device)

    def SHEncode(self, dirs):
        return ops.sh_encode(dirs.contiguous(), self.degree_)

    def Query(self, feats, dirs):
        """SHShader::Query (SHShader.cpp:23-29): feats [n,16] fp32 (already holding the constant-1 channel
        and the appearance embedding), dirs [n,3] -> rgb [n,3]; differentiable w.r.t. feats and mlp params."""
        x = torch.cat([feats, self.SHEncode(dirs)], -1)
        out = self.mlp_.Query(x)
        eps = 1e-3
        return (1. + 2. * eps) / (1. + torch.exp(-out)) - eps

    def States(self):
        return [self.mlp_.params_.data]

    def LoadStates(self, states, idx):
        self.mlp_.params_.data.copy_(states[idx])
        return idx + 1

    def OptimParamGroups(self):
        lr = self.global_data_pool_.learning_rate_
        return [dict(params=[self.mlp_.params_], lr=lr, betas=(0.9, 0.99), eps=1e-15, weight_decay=1e-6)]

    def Reset(self):
        self.mlp_.InitParams()
