"""RNG-stream compatibility with the reference.

The reference allocates every tiny-cuda-nn output with ``torch::rand`` (TCNNWP.cpp:143-144): each MLP forward
consumes the CUDA generator's Philox stream for a ``[ceil128(batch), 16]`` fp16 tensor whose values are then
overwritten.  So the draws that FOLLOW an MLP call — the TV-loss edge samples inside ``Renderer::Render``
(PersSampler.cu:456-457), the next iteration's ray noise and background — depend on it.  GradientScaling's
backward does the same with an unused ``rand_like`` (CustomOps.cu:154).  A drop-in must leave the generator in
the same state, but has no reason to write 134 MB of random numbers to do so: these helpers advance the
generator's Philox offset by exactly what ATen's ``distribution_nullary_kernel`` would have consumed
(ATen/native/cuda/DistributionTemplates.h: calc_execution_policy) without launching anything.
"""
import torch

_BLOCK = 256          # block_size_bound
_UNROLL = 4           # sizeof(float4) / sizeof(float): fp16 and fp32 uniforms both draw curand_uniform4
_PER_CALL = 4         # max_generator_offsets_per_curand_call
_props = {}


def _device_index(device):
    device = torch.device(device)
    return torch.cuda.current_device() if device.index is None else device.index


def rand_philox_offset(numel, device):
    """Philox offset increment of one ``torch.rand`` / ``rand_like`` call producing ``numel`` fp16/fp32 elements."""
    if numel <= 0:
        return 0
    idx = _device_index(device)
    if idx not in _props:
        p = torch.cuda.get_device_properties(idx)
        _props[idx] = p.multi_processor_count * (p.max_threads_per_multi_processor // _BLOCK)
    grid = min(_props[idx], (numel + _BLOCK - 1) // _BLOCK)
    return ((numel - 1) // (_BLOCK * grid * _UNROLL) + 1) * _PER_CALL


def burn_rand(numel, device):
    """Advance the default CUDA generator as if ``torch.rand(numel)`` had run on ``device`` (no kernel launch)."""
    inc = rand_philox_offset(int(numel), device)
    if inc:
        gen = torch.cuda.default_generators[_device_index(device)]
        gen.set_offset(gen.get_offset() + inc)


def burn_mlp_output(batch, device):
    """The reference's TCNNWP::Query on ``batch`` rows: torch::rand({ceil128(batch), 16}, fp16)."""
    if batch > 0:
        burn_rand(((int(batch) + 127) // 128 * 128) * 16, device)
