"""ORACLE (test infrastructure, NOT product code) -- the ico2ico / ico2ico_vae training graphs on CPU.

Self-contained restatement of the two graphs the hot path trains (/root/reference/models.py):

    residual blocks           models.py:22-40 (down: conv00/conv10 stride 2), :42-62 (up: IcoUpsampleS2S first)
    encoder / decoder stacks  models.py:101-160 (ico2ico), :162-216 (ico2ico_vae)
    ico2ico                   models.py:219-232      out = enc2icoConv(decoder(enc(encoder(x))))
    ico2ico_vae               models.py:254-300      mu/logvar heads :268-286, reparameterize :89-92

built ONLY over oracle/icocnn_ref.py + torch -- nothing here imports the product package, so the CPU arm of bench.py
(`--impl reference`) and the parity tests can run it without loading libgeniconet_b200.so.  Module attribute names
are the reference's, hence the state-dict keys are too (tests/test_models_host.py checks them key for key against the
reference's own models.py in the build container).  The reference hard-codes level 5; `level` shifts every layer.

    parity unpinned (inherits oracle/icocnn_ref.py; see oracle/ico_geometry_ref.py).
"""
import torch
from torch import nn

from .icocnn_ref import IcoConvS2S, IcoUpsampleS2S

# channel plan of the two graphs: encoder widths after the stem, decoder widths from the latent down to the head input
ENCODER = {'ico2ico': (64, 128, 256, 256), 'ico2ico_vae': (64, 128, 256)}
DECODER = {'ico2ico': (256, 256, 128, 64), 'ico2ico_vae': (512, 256, 128, 64)}
LOSS_FACTORS = {'ico2ico': (1.0, 0.0, 0.0), 'ico2ico_vae': (0.6, 0.2, 0.2)}        # run.py:689-696


def _conv(cin, cout, stride, level, corner_mode):
    return IcoConvS2S(in_features=cin, out_features=cout, stride=stride, bias=True, subdivisions=level, corner_mode=corner_mode)


class ResBlock(nn.Module):
    """relu( bn01(conv01(relu(bn00(conv00(a))))) + bn10(conv10(b)) ), a = b = x (down) or the two upsamples of x (up)."""

    def __init__(self, cin, cout, level_in, corner_mode, up):
        super().__init__()
        level_out = level_in + 1 if up else level_in - 1
        first = (cin, cout, 1, level_out, corner_mode) if up else (cin, cout, 2, level_in, corner_mode)
        if up:
            self.upsample00 = IcoUpsampleS2S(cin, level_in, corner_mode)
        self.conv00 = _conv(*first)
        self.icobn00 = nn.BatchNorm2d(cout)
        self.conv01 = _conv(cout, cout, 1, level_out, corner_mode)
        self.icobn01 = nn.BatchNorm2d(cout)
        if up:
            self.upsample10 = IcoUpsampleS2S(cin, level_in, corner_mode)
        self.conv10 = _conv(*first)
        self.icobn10 = nn.BatchNorm2d(cout)
        self.up = up

    def forward(self, x):
        a = self.upsample00(x) if self.up else x
        b = self.upsample10(x) if self.up else x
        main = self.icobn01(self.conv01(torch.relu(self.icobn00(self.conv00(a)))))
        return torch.relu(main + self.icobn10(self.conv10(b)))


def make_encoder(name, level, corner_mode):
    w = ENCODER[name]
    mods = [_conv(3, w[0], 1, level, corner_mode), nn.BatchNorm2d(w[0]), nn.ReLU()]
    mods += [ResBlock(w[d], w[d + 1], level - d, corner_mode, up=False) for d in range(len(w) - 1)]
    return nn.Sequential(*mods)


def make_decoder(name, level, corner_mode):
    w = DECODER[name]
    body = nn.Sequential(*[ResBlock(w[d], w[d + 1], level - 3 + d, corner_mode, up=True) for d in range(3)])
    head = nn.Sequential(nn.Conv2d(w[3], 3, kernel_size=1), nn.Tanh())
    return body, head


class Ico2Ico(nn.Module):
    def __init__(self, level=5, corner_mode='average'):
        super().__init__()
        self.encoder = make_encoder('ico2ico', level, corner_mode)
        self.enc = nn.Identity()
        self.decoder, self.enc2icoConv = make_decoder('ico2ico', level, corner_mode)

    def forward(self, x):
        return self.enc2icoConv(self.decoder(self.enc(self.encoder(x))))


class Ico2IcoVAE(nn.Module):
    """forward(x, eps=None): `eps` (same shape as mu) replaces torch.randn_like so a test can feed the CUDA path's noise."""

    def __init__(self, level=5, corner_mode='average'):
        super().__init__()
        self.encoder = make_encoder('ico2ico_vae', level, corner_mode)
        self.mu = nn.Sequential(_conv(256, 512, 2, level - 2, corner_mode), nn.BatchNorm2d(512))
        self.logvar = nn.Sequential(_conv(256, 512, 2, level - 2, corner_mode), nn.BatchNorm2d(512))
        self.mu_hook, self.logvar_hook, self.reparameterize_hook = nn.Identity(), nn.Identity(), nn.Identity()
        self.decoder, self.final_layer = make_decoder('ico2ico_vae', level, corner_mode)

    def forward(self, x, eps=None):
        h = self.encoder(x)
        mu, logvar = self.mu_hook(self.mu(h)), self.logvar_hook(self.logvar(h))
        noise = torch.randn_like(mu) if eps is None else eps
        z = self.reparameterize_hook(noise * torch.exp(0.5 * logvar) + mu)
        return self.final_layer(self.decoder(z)), mu, logvar


def build(name, level=5, corner_mode='average'):
    return {'ico2ico': Ico2Ico, 'ico2ico_vae': Ico2IcoVAE}[name](level, corner_mode)


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


# ------------------------------------------------------------------ losses (losses.py:47-82, 92-108, 137-142)
_mesh_cache = {}


def _mesh(level):
    if level not in _mesh_cache:
        from . import ico_geometry_ref as geo, mesh_ref
        faces = torch.from_numpy(geo.get_ico_faces(level))
        adj = mesh_ref.compute_adjacency_matrix_sparse(int(faces.max()) + 1, faces)
        _mesh_cache[level] = (faces, adj, torch.from_numpy(geo.pole_rings(level)))
    return _mesh_cache[level]


def p2p_loss(level, out, target, f_pos, f_nor, f_lap):
    """Point2Point_Loss.forward: (loss, (l_pos, l_nor, l_lap)); out [B,3,5n,2n], target [B,9,P+2]."""
    from . import mesh_ref
    faces, adj, rings = _mesh(level)
    B, C = out.shape[:2]
    flat = out.reshape(B, C, -1)
    v = torch.cat((flat, flat[:, :, rings].mean(-1)), dim=2).transpose(1, 2).contiguous()      # losses.py:49-51
    t = target.transpose(1, 2).contiguous()
    l_pos = nn.functional.mse_loss(v, t[:, :, :3])
    l_nor = torch.mean(1 - nn.functional.cosine_similarity(mesh_ref.compute_vertex_normals(v, faces), t[:, :, 3:6], dim=2))
    l_lap = nn.functional.mse_loss(mesh_ref.compute_laplacian_batch(v, adj), t[:, :, 6:9])
    return f_pos * l_pos + f_nor * l_nor + f_lap * l_lap, (l_pos, l_nor, l_lap)


def kld_loss(mu, logvar):
    """KLD_Loss.forward: mean over the batch of -0.5 * mean_i(1 + logvar - mu^2 - exp(logvar))."""
    mu, logvar = torch.flatten(mu, 1), torch.flatten(logvar, 1)
    return torch.mean(-0.5 * torch.mean(1 + logvar - mu.pow(2) - logvar.exp(), dim=1), dim=0)


def training_loss(name, level, out, target, factors=None, factor_kl=1.0):
    """The scalar run.py:244-250 back-propagates for `name` (P2P_Loss, or P2PKLD_Loss = recons + factor_kl * kld)."""
    f = LOSS_FACTORS[name] if factors is None else factors
    if name == 'ico2ico_vae':
        rec, mu, logvar = out
        return p2p_loss(level, rec, target, *f)[0] + factor_kl * kld_loss(mu, logvar)
    return p2p_loss(level, out, target, *f)[0]
