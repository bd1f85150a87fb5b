"""Test tool: build the geniconet_b200.models graphs over the ORACLE layers (CPU), so model-level
parity can be checked on the GPU box where /root/reference does not exist."""
import contextlib

import torch

from geniconet_b200 import models as gm
from oracle import icocnn_ref


@contextlib.contextmanager
def oracle_layers():
    saved = (gm.IcoConvS2S, gm.IcoUpsampleS2S, gm._reparameterize)
    gm.IcoConvS2S, gm.IcoUpsampleS2S = icocnn_ref.IcoConvS2S, icocnn_ref.IcoUpsampleS2S
    try:
        yield
    finally:
        gm.IcoConvS2S, gm.IcoUpsampleS2S, gm._reparameterize = saved


def build_oracle_model(name, params):
    with oracle_layers():
        return getattr(gm, name)(params)


def fill_params_deterministic(model, seed=0):
    """Name-keyed deterministic weights: independent of construction order and RNG consumption."""
    import zlib
    with torch.no_grad():
        for name, p in sorted(model.named_parameters()):
            g = torch.Generator().manual_seed(seed * 1000003 + zlib.crc32(name.encode()))
            if p.dim() == 1:
                if name.endswith('weight'):      # BN scale
                    p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
                else:
                    p.copy_(0.05 * torch.randn(p.shape, generator=g))
            else:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) * (1.5 / fan_in) ** 0.5)
    return model


def ref_p2p_loss(level, out, target, f_pos, f_nor, f_lap):
    """Point2Point_Loss.forward (losses.py:47-82) restated over the oracle mesh helpers."""
    from oracle import ico_geometry_ref as geo, mesh_ref
    faces = torch.from_numpy(geo.get_ico_faces(level))
    adj = mesh_ref.compute_adjacency_matrix_sparse(int(faces.max()) + 1, faces)
    rings = torch.from_numpy(geo.pole_rings(level))
    B, C = out.shape[:2]
    flat = out.reshape(B, C, -1)
    poles = flat[:, :, rings].mean(-1)
    v = torch.cat((flat, poles), dim=2).transpose(1, 2).contiguous()
    nrm = mesh_ref.compute_vertex_normals(v, faces)
    lap = mesh_ref.compute_laplacian_batch(v, adj)
    t = target.transpose(1, 2).contiguous()
    l_pos = torch.nn.functional.mse_loss(v, t[:, :, :3])
    l_nor = torch.mean(1 - torch.nn.functional.cosine_similarity(nrm, t[:, :, 3:6], dim=2))
    l_lap = torch.nn.functional.mse_loss(lap, t[:, :, 6:9])
    return f_pos * l_pos + f_nor * l_nor + f_lap * l_lap, (l_pos, l_nor, l_lap)


def ref_kld(mu, logvar):
    mu, logvar = torch.flatten(mu, 1), torch.flatten(logvar, 1)
    return torch.mean(-0.5 * torch.mean(1 + logvar - mu.pow(2) - logvar.exp(), dim=1), dim=0)
