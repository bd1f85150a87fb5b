"""VAE reparameterisation (reference models.py:89-92, row a6) as one fused kernel.

    std = exp(0.5*logvar); eps = randn_like(std); z = eps*std + mu

eps comes from Philox4x32-10 keyed by (seed, offset); it is returned so a checker can reproduce z.
"""
import torch

from . import _lib
from .ico_conv import _stream, _require_cuda_f32

_state = {'seed': None, 'offset': 0}


def manual_seed(seed):
    _state['seed'], _state['offset'] = int(seed) & (2 ** 64 - 1), 0


_counters = {}


def _graph_counter(device):
    key = (device.type, device.index)
    if key not in _counters:
        _counters[key] = torch.zeros(1, dtype=torch.int64, device=device)
    return _counters[key]


def _next_key():
    if _state['seed'] is None:
        _state['seed'] = torch.initial_seed() & (2 ** 64 - 1)
    _state['offset'] += 1
    return _state['seed'], _state['offset']


class _ReparamFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, logvar, seed, offset):
        _require_cuda_f32(mu, 'reparameterize')
        _require_cuda_f32(logvar, 'reparameterize')
        if mu.shape != logvar.shape:
            raise ValueError('reparameterize: mu %s and logvar %s differ in shape' % (tuple(mu.shape), tuple(logvar.shape)))
        if mu.stride() != logvar.stride() or not (mu.is_contiguous() or mu.is_contiguous(memory_format=torch.channels_last)):
            mu, logvar = mu.contiguous(), logvar.contiguous()
        eps, z = torch.empty_like(mu), torch.empty_like(mu)
        if torch.cuda.is_current_stream_capturing():
            # the (seed, offset) scalars would be frozen into the graph: add a device-side step counter that every replay advances
            step = _graph_counter(mu.device)
            _lib.check(_lib.lib.gin_reparam_fwd_step(mu.data_ptr(), logvar.data_ptr(), eps.data_ptr(), z.data_ptr(), mu.numel(),
                                                     seed, offset, step.data_ptr(), _stream()), 'gin_reparam_fwd_step')
        else:
            _graph_counter(mu.device)       # exists before any capture (allocating it inside one would re-zero it on every replay)
            _lib.check(_lib.lib.gin_reparam_fwd(mu.data_ptr(), logvar.data_ptr(), eps.data_ptr(), z.data_ptr(), mu.numel(),
                                                seed, offset, _stream()), 'gin_reparam_fwd')
        ctx.save_for_backward(logvar, eps)
        ctx.mark_non_differentiable(eps)
        return z, eps

    @staticmethod
    def backward(ctx, dz, _deps):
        logvar, eps = ctx.saved_tensors
        if dz.stride() != logvar.stride():
            dz = dz.contiguous(memory_format=torch.channels_last) if logvar.dim() == 4 and not logvar.is_contiguous() else dz.contiguous()
            if dz.stride() != logvar.stride():
                dz = torch.empty_like(logvar).copy_(dz)
        dmu, dlv = torch.empty_like(logvar), torch.empty_like(logvar)
        _lib.check(_lib.lib.gin_reparam_bwd(dz.data_ptr(), logvar.data_ptr(), eps.data_ptr(), dmu.data_ptr(), dlv.data_ptr(),
                                            logvar.numel(), _stream()), 'gin_reparam_bwd')
        return dmu, dlv, None, None


def reparameterize(mu, logvar, seed=None, offset=None, return_eps=False):
    if seed is None:
        seed, offset = _next_key()
    z, eps = _ReparamFn.apply(mu, logvar, int(seed), int(offset or 0))
    return (z, eps) if return_eps else z
