"""Generate the committed golden vectors under ``tests/golden/`` (TEST INFRASTRUCTURE ONLY).

Run in the build container, where ``/root/reference`` exists:

    python -m oracle.make_golden

* ``splat_*.npz``        inputs + outputs of the reference's own splat kernel strings
                         compiled for the host (``oracle/build_ref.py``).
* ``correlation_*.npz``  same for ``kernel_Correlation_rearrange`` + ``updateOutput``.
* ``decoder_*.npz``      hot-path inputs (encoder / ``flow_process`` outputs captured by
                         forward hooks), hot-path weights and the outputs of the
                         UNMODIFIED ``LunaTokis.forward`` run on the CPU under
                         ``oracle/ref_shims.py``.  ``decoder_raft`` uses the reference's
                         RAFT; the two smaller cases replace the (out-of-scope) RAFT
                         module by a seeded smooth-flow stand-in so that LR sizes below
                         RAFT's 128-px floor can be used -- everything from
                         ``flow_process`` to the clamp is reference code either way.

All arrays are fp32/int64 little-endian ``numpy`` in compressed ``.npz``.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import build_ref, decoder_ref, ref_shims  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

HOT_KEYS = ("imnet", "flow_imnet", "synth_net", "alpha")


def _save(name, **arrays):
    os.makedirs(GOLDEN, exist_ok=True)
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print("wrote", name, {k: tuple(v.shape) for k, v in out.items() if v.ndim})


def splat_cases():
    g = torch.Generator().manual_seed(1234)
    cases = {"s05": (2, 5, 17, 23, 0.5), "s4": (1, 6, 16, 20, 4.0), "s32": (2, 3, 9, 31, 32.0)}
    for tag, (n, c, h, w, sigma) in cases.items():
        inp = torch.randn(n, c, h, w, generator=g)
        flow = torch.randn(n, 2, h, w, generator=g) * sigma
        # integer-valued, exactly-on-border and far out-of-frame flows
        flow[0, 0, 0, :6] = torch.tensor([1.0, -2.0, 0.0, 3.0, -4.0, float(w)])
        flow[0, 1, 0, :6] = torch.tensor([0.0, 1.0, -1.0, 2.0, 0.0, 0.0])
        flow[0, :, h - 1, w - 1] = torch.tensor([0.0, 0.0])
        flow[0, :, h - 1, w - 2] = torch.tensor([1.0, 0.0])
        flow[0, :, 1, 0] = torch.tensor([-0.5, -0.5])
        metric = -torch.randn(n, 1, h, w, generator=g).abs() * 2.0
        soft_in = torch.cat([inp * metric.exp(), metric.exp()], 1)
        lin_in = torch.cat([inp * metric, metric], 1)
        avg_in = torch.cat([inp, torch.ones(n, 1, h, w)], 1)
        _save(
            f"splat_{tag}",
            input=inp, flow=flow, metric=metric,
            out_summation=build_ref.ref_splat_sum(inp, flow),
            out_softmax=build_ref.ref_splat_sum(soft_in, flow),
            out_linear=build_ref.ref_splat_sum(lin_in, flow),
            out_average=build_ref.ref_splat_sum(avg_in, flow),
            out_max=build_ref.ref_splat_max(inp, flow),
            out_max_exp=build_ref.ref_splat_max(metric.exp().expand(n, 1, h, w).contiguous(), flow),
            out_count=build_ref.ref_splat_count(torch.ones(n, 1, h, w), flow),
        )


def correlation_cases():
    g = torch.Generator().manual_seed(99)
    for tag, (b, c, h, w) in {"c8": (1, 8, 5, 6), "c32": (2, 32, 12, 20), "c196": (1, 196, 6, 10)}.items():
        f1 = torch.randn(b, c, h, w, generator=g)
        f2 = torch.randn(b, c, h, w, generator=g)
        _save(f"correlation_{tag}", first=f1, second=f2, out=build_ref.ref_correlation(f1, f2))


class _SmoothFlow(torch.nn.Module):
    """Seeded smooth-flow stand-in for RAFT (out-of-scope surround), same call signature."""

    def __init__(self, seed, magnitude):
        super().__init__()
        self.seed, self.magnitude = seed, magnitude
        self.args = type("A", (), {"alternate_corr": False})()

    def forward(self, img1, img2, iters=4):
        n, _, hh, ww = img1.shape
        g = torch.Generator().manual_seed(self.seed)
        low = torch.randn(n, 2, max(hh // 16, 2), max(ww // 16, 2), generator=g) * self.magnitude
        return [torch.nn.functional.interpolate(low, size=(hh, ww), mode="bilinear", align_corners=False)]


def _load_hot_params(model, params):
    sd = model.state_dict()
    for k, v in params.items():
        assert sd[k].shape == v.shape, k
        sd[k].copy_(v)


def decoder_case(name, lr_hw, scale, times, batch, use_raft, seed, gain, first_gain, alpha, ensemble=False, z_bias=0.03):
    model = ref_shims.build_reference_model(seed=seed, splat_backend="reference_kernels")
    model.local_ensemble = ensemble  # Ours.py:453 (False as shipped); True = the four-latent LIIF ensemble of Ours.py:660-663, 758-764
    if not use_raft:
        model.flow_predictor = _SmoothFlow(seed + 7, magnitude=3.0)
    if gain is not None:
        _load_hot_params(model, decoder_ref.random_params(seed=seed + 11, weight_gain=gain, first_gain=first_gain,
                                                          alpha=alpha, rgb_bias=0.5, rgb_gain=3.0, z_bias=z_bias))
    torch.manual_seed(seed + 1)
    h, w = lr_hw
    low = torch.rand(batch, 2, 3, max(h // 4, 2), max(w // 4, 2))
    x = torch.nn.functional.interpolate(low.view(batch * 2, 3, *low.shape[-2:]), size=(h, w), mode="bilinear").view(batch, 2, 3, h, w)
    x = (x + 0.05 * torch.rand(batch, 2, 3, h, w)).clamp(0, 1)
    target_t = [torch.full((batch, 1), float(t)) for t in times]
    r = ref_shims.run_reference_forward(model, x, target_t, scale)
    assert not any(torch.isnan(v).any() for v in r.values()), name
    sd = model.state_dict()
    weights = {k.replace(".", "__"): v for k, v in sd.items() if k.split(".")[0] in HOT_KEYS}
    HH, WW = r["out"].shape[-2:]
    _save(
        name,
        feat=r["feat"], flow_feat=r["flow_feat"], residual=r["residual"],
        target_t=torch.stack(target_t, 1).squeeze(-1), hr_size=np.array([HH, WW]), scale=np.array([scale], dtype=np.float64),
        out=r["out"], flow_out=r["flow_out"], **weights,
    )


def decoder_cases():
    # reference RAFT, reference default initialisation (alpha = -20: exp(z) ~ 1)
    decoder_case("decoder_raft", (32, 48), 4, [0.25, 0.5], 1, True, seed=0, gain=None, first_gain=None, alpha=None)
    # non-degenerate variant: first layers x4, RGB centred in the clamp range, z switching on/off (alpha=-20
    # as initialised by the reference, Ours.py:509); hidden layers keep the reference SIREN scale -- larger
    # hidden gains make the fp32 reference itself ill-conditioned (1-ulp weight changes move flows by >1e-3 px)
    decoder_case("decoder_x4", (16, 24), 4, [0.125, 0.5, 0.875], 1, False, seed=1, gain=1.0, first_gain=4.0, alpha=-20.0)
    # non-integer scale (round(H*3.5)), two clips: exercises index ties and batch ordering
    decoder_case("decoder_x3p5_b2", (16, 20), 3.5, [0.3, 0.75], 2, False, seed=2, gain=1.0, first_gain=4.0, alpha=-20.0)
    ensemble_case()


def front_cases():
    """Reliability-map front end (Ours.py:562-578, 613-637): LR frames, the HR flows RAFT (or its stand-in) returned and
    the tensor the reference handed to flow_process, captured while its unmodified forward ran."""
    for name, lr_hw, scale, batch, use_raft, seed in (("front_raft", (32, 48), 4, 1, True, 5), ("front_smooth_b2", (20, 28), 3, 2, False, 6)):
        model = ref_shims.build_reference_model(seed=seed, splat_backend="reference_kernels")
        if not use_raft:
            model.flow_predictor = _SmoothFlow(seed + 7, magnitude=6.0)
        torch.manual_seed(seed + 1)
        h, w = lr_hw
        low = torch.rand(batch, 2, 3, max(h // 4, 2), max(w // 4, 2))
        x = torch.nn.functional.interpolate(low.view(batch * 2, 3, *low.shape[-2:]), size=(h, w), mode="bilinear").view(batch, 2, 3, h, w)
        x = (x + 0.05 * torch.rand(batch, 2, 3, h, w)).clamp(0, 1)
        r = ref_shims.run_reference_forward(model, x, [torch.full((batch, 1), 0.5)], scale)
        _save(name, x=x, flow_hr=r["flow_hr"], g_filter=model.g_filter.detach().reshape(3, 3), flow_process_in=r["flow_process_in"])


def raft_corr_case():
    """RAFT correlation lookup: the reference's own CorrBlock class (models/core/corr.py:8-56) on seeded feature maps and
    coordinates that leave the image on two sides (zero padding) and hit integer positions."""
    ref_shims.import_reference()  # module stubs (alt_cuda_corr, cupy) + sys.path
    from models.core.corr import CorrBlock  # the reference's class, imported from where it lies

    g = torch.Generator().manual_seed(21)
    B, C, H, W, r = 2, 32, 16, 24, 3  # coarsest level 2 x 3: bilinear_sampler divides by (H - 1)
    fmap1, fmap2 = torch.randn(B, C, H, W, generator=g), torch.randn(B, C, H, W, generator=g)
    ys, xs = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    coords = torch.stack([xs, ys], 0)[None].repeat(B, 1, 1, 1) + torch.randn(B, 2, H, W, generator=g) * 3.0
    coords[0, :, 0, 0] = torch.tensor([2.0, 3.0])      # exactly on a sample
    coords[0, :, 0, 1] = torch.tensor([-5.0, 20.0])    # far outside
    coords[1, :, 5, 5] = torch.tensor([W - 1.0, H - 1.0])
    out = CorrBlock(fmap1, fmap2, num_levels=4, radius=r)(coords)
    _save("raft_corr", fmap1=fmap1, fmap2=fmap2, coords=coords, radius=np.array([r]), out=out)


def regime_cases():
    """Weight regimes SURVEY 8c asks for beyond the reference initialisation: alpha in {-1, +0.5}
    (exp(z) not degenerate; alpha > 0 makes max-splat candidates exceed the 1.0 the output starts at, Ours.py:794, 834,
    softsplat_max_cp.py:254) and O(1)-scaled SIREN hidden weights (gain 2 and 4: sine arguments up to ~100)."""
    decoder_case("decoder_alpha_m1", (12, 16), 4, [0.3, 0.8], 1, False, seed=4, gain=1.0, first_gain=4.0, alpha=-1.0)
    decoder_case("decoder_alpha_p05", (12, 16), 4, [0.3, 0.8], 1, False, seed=5, gain=1.0, first_gain=4.0, alpha=0.5, z_bias=0.6)
    decoder_case("decoder_gain2", (12, 16), 4, [0.5], 1, False, seed=6, gain=2.0, first_gain=4.0, alpha=-1.0)
    decoder_case("decoder_gain4", (12, 16), 4, [0.5], 1, False, seed=7, gain=4.0, first_gain=4.0, alpha=-1.0)


def ensemble_case():
    # the reference's local_ensemble=True branch (four shifted latents, diagonally swapped area weights)
    decoder_case("decoder_ens_x3", (12, 16), 3, [0.4], 2, False, seed=3, gain=1.0, first_gain=4.0, alpha=-20.0, ensemble=True)


if __name__ == "__main__":
    if not build_ref.reference_available():
        raise SystemExit("reference checkout absent: golden vectors can only be regenerated in the build container")
    if "--ensemble-only" in sys.argv:
        ensemble_case()
        raise SystemExit(0)
    if "--regimes-only" in sys.argv:
        regime_cases()
        raise SystemExit(0)
    if "--raft-corr-only" in sys.argv:
        raft_corr_case()
        raise SystemExit(0)
    if "--front-only" in sys.argv:
        front_cases()
        raise SystemExit(0)
    splat_cases()
    correlation_cases()
    decoder_cases()
    regime_cases()
    front_cases()
    raft_corr_case()
