"""GPU: ``luna_tokis.install`` + ``forward_b200`` end to end on CUDA.  The reference checkout is absent on the GPU box, so the
model is a stand-in with the reference's module tree names, flags and ``state_dict`` keys for everything the patched forward
touches (``Ours.py:414-510``): a smooth-flow ``flow_predictor``, small conv ``encoder`` / ``flow_process``, the three SIREN
MLPs.  What is checked is the glue on the device: fused front end -> sub-modules -> decoder, return triple, decoder cache."""
import math

import pytest
import torch
import torch.nn as nn

from oracle import decoder_ref

pytestmark = pytest.mark.gpu


class _SineLayer(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.linear = nn.Linear(i, o)


class _Siren(nn.Module):
    def __init__(self, i, hidden, o):
        super().__init__()
        w = [i] + hidden
        self.net = nn.Sequential(*[_SineLayer(w[k], w[k + 1]) for k in range(len(hidden))], nn.Linear(hidden[-1], o))


class _Flow(nn.Module):
    def forward(self, a, b, iters=4):
        n, _, hh, ww = a.shape
        g = torch.Generator().manual_seed(5)
        low = torch.randn(n, 2, max(hh // 16, 2), max(ww // 16, 2), generator=g).to(a.device) * 3.0
        return [torch.nn.functional.interpolate(low, size=(hh, ww), mode="bilinear", align_corners=False)]


class _Encoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.c = nn.Conv2d(3, 64, 3, padding=1)

    def forward(self, x, _):            # [B, 2, 3, H, W] -> [B, 3, 64, H, W] (two frames and the one between them)
        f0, f1 = self.c(x[:, 0]), self.c(x[:, 1])
        return torch.stack([f0, 0.5 * (f0 + f1), f1], 1)


class FakeLunaTokis(nn.Module):
    def __init__(self):
        super().__init__()
        self.flow_predictor, self.encoder = _Flow(), _Encoder()
        self.flow_process = nn.Conv2d(14, 64, 3, padding=1)
        self.flow_imnet, self.imnet, self.synth_net = _Siren(67, [64, 64, 256], 3), _Siren(66, [64, 64, 256], 64), _Siren(198, [64, 64, 64, 256], 3)
        self.alpha = nn.Parameter(torch.ones(1) * -20.0)
        self.g_filter = nn.Parameter(torch.tensor([[1., 2., 1.], [2., 4., 2.], [1., 2., 1.]]).view(1, 1, 1, 3, 3) / 16.0, requires_grad=False)
        self.trans, self.input_Z, self.res_liff, self.siren, self.warp_to_many, self.local_ensemble = False, True, False, True, False, False

    def forward(self, *a, **k):
        raise RuntimeError("the reference forward is not part of this stand-in")


def _model():
    torch.manual_seed(0)
    m = FakeLunaTokis()
    p = decoder_ref.random_params(seed=5, **decoder_ref.REALISTIC)
    sd = m.state_dict()
    for k, v in p.items():
        sd[k].copy_(v)
    with torch.no_grad():
        m.encoder.c.weight.mul_(0.5)
    return m.cuda().eval()


def test_installed_forward_runs_the_b200_path_on_cuda():
    from motif_b200 import luna_tokis
    from motif_b200.decoder import SpaceTimeDecoder

    model = _model()
    keys = list(model.state_dict().keys())
    luna_tokis.install(model, raft_lookup=False)
    assert list(model.state_dict().keys()) == keys
    torch.manual_seed(1)
    B, H, W, scale = 1, 24, 32, 4
    x = torch.rand(B, 2, 3, H, W).cuda()
    target_t = [torch.full((B, 1), 0.25).cuda(), torch.full((B, 1), 0.75).cuda()]
    rgb, flow, flow_gt = model(x, None, target_t, scale, use_GT=False, iter=2)
    assert rgb.shape == (2, B, 3, H * scale, W * scale) and flow.shape == (2 * B * 2, 2, H * scale, W * scale) and flow_gt == 0.0
    assert rgb.min().item() >= 0.0 and rgb.max().item() <= 1.0 and torch.isfinite(flow).all()
    # the same latents through the decoder directly, and through the CPU oracle
    with torch.no_grad():
        feat, ff, res, tt, hr = luna_tokis.surround(model, x, target_t, scale, iter=2)
    assert hr == (H * scale, W * scale)
    dec = SpaceTimeDecoder.from_state_dict(model.state_dict(), device="cuda")
    rgb2, flow2 = dec.decode(feat.float(), ff.float(), res.float(), tt, hr)
    assert torch.equal(flow, flow2) and (rgb - rgb2).abs().max().item() < 1e-5
    params = {k: v.detach().cpu() for k, v in model.state_dict().items() if k == "alpha" or k.split(".")[0] in ("imnet", "flow_imnet", "synth_net")}
    r_rgb, r_flow = decoder_ref.decode(feat.cpu(), ff.cpu(), res.cpu(), tt.cpu(), hr[0], hr[1], params)
    assert (flow.cpu() - r_flow).abs().max().item() < 2e-6
    stable = ~decoder_ref.count_unstable_mask(r_flow * 20.0 * scale, B, 2).expand_as(r_rgb)
    assert (rgb.cpu() - r_rgb).abs()[stable].max().item() < 1e-3
    # second call reuses the cached decoder; an in-place weight update invalidates it
    cache = model.__dict__["_motif_decoders"]
    d0 = next(iter(cache.values()))[0]
    model(x, None, target_t, scale, use_GT=False, iter=2)
    assert next(iter(cache.values()))[0] is d0
    with torch.no_grad():
        model.synth_net.net[4].bias.add_(0.05)
    rgb3, _, _ = model(x, None, target_t, scale, use_GT=False, iter=2)
    assert next(iter(cache.values()))[0] is not d0 and (rgb3 - rgb).abs().max().item() > 1e-3
    model.train()
    with pytest.raises(NotImplementedError):
        model(x, None, target_t, scale, use_GT=False)
