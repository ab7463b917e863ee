"""GPU parity of the default f16x3 decoder at BASELINE's FULL sizes against the reference's GPU path
(``oracle/decoder_ref_gpu.py``: the oracle's eager torch operators on CUDA tensors + the reference's own splat kernels
compiled unmodified for sm_100a), with the max-abs gate of north_star -- not a statistical one -- and the error INSIDE the
excluded count-unstable set reported; plus the weight regimes SURVEY 8c asks for (alpha in {-1, +0.5}, SIREN gains 2 / 4)
against goldens captured from the reference's own forward (``oracle/make_golden.py``).

Every test appends its measured numbers to ``gpurun_out/parity_report.jsonl`` (DESIGN.md section 3 quotes them)."""
import json
import os

import pytest
import torch

from conftest import ROOT, hot_params, load_golden, psnr
from oracle import decoder_ref, decoder_ref_gpu

pytestmark = pytest.mark.gpu

TOL = 1e-3
PSNR_MIN = 60.0
FLOW_TOL = 2e-6


def _report(**kw):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.jsonl"), "a") as f:
        f.write(json.dumps(kw) + "\n")
    print(kw)


def _need_ref():
    if not decoder_ref_gpu.available():
        pytest.skip("oracle/_ref/ref_gpu_sm100a.so not built (reference checkout absent at build time)")


def _fullsize(workload, H, W, HH, WW, times, seed, chunk):
    from motif_b200 import synthetic
    from motif_b200.decoder import SpaceTimeDecoder

    _need_ref()
    feat, ff, res = [t.cuda() for t in synthetic.synthetic_latents(1, H, W, seed=seed)]
    params = decoder_ref.random_params(seed=2, **decoder_ref.REALISTIC)
    tt = torch.tensor([times])
    B, N = tt.shape
    rgb, flow = SpaceTimeDecoder(params, device="cuda", precision="f16x3").decode(feat, ff, res, tt, (HH, WW))
    r_rgb, r_flow, aux = decoder_ref_gpu.decode(feat, ff, res, tt, HH, WW, params, chunk=chunk)
    assert rgb.shape == r_rgb.shape and flow.shape == r_flow.shape
    d_flow = (flow - r_flow).abs().max().item()
    # the two places where the reference function is discontinuous: the count splat (floor of the landing position) and the
    # exact-equality tests on the blended normaliser (Ours.py:813, 829), which its own float atomics decide differently from
    # run to run -- measured right here as reference-vs-reference
    r2_rgb, _, _ = decoder_ref_gpu.decode(feat, ff, res, tt, HH, WW, params, chunk=chunk)
    self_d = (r_rgb - r2_rgb).abs()
    ref_self_max, ref_self_over = self_d.max().item(), int((self_d > TOL).sum().item())
    del r2_rgb, self_d
    m_count = decoder_ref_gpu.count_unstable_mask(aux["flow_hr"], B, N)
    # raw flow_imnet outputs agree to d_flow; z_raw is the third output of the same layer, amplified by |alpha| inside exp()
    m_eq = decoder_ref_gpu.equality_unstable_mask(aux["wz"], alpha=float(params["alpha"][0]), z_err=max(d_flow, 1e-7))
    unstable = (m_count | m_eq).expand_as(r_rgb)
    frac_count, frac_eq = m_count.float().mean().item(), m_eq.float().mean().item()
    frac = unstable.float().mean().item()
    d = (rgb - r_rgb).abs()
    d_out = d[~unstable].max().item()
    n_out_bad = int((d[~unstable] > TOL).sum().item())
    inside = d[unstable]
    n_in = int(inside.numel())
    n_in_bad = int((inside > TOL).sum().item())
    max_in = inside.max().item() if n_in else 0.0
    offenders = []
    if n_out_bad:  # diagnostics: where, and what the reference's normaliser looks like there
        dm = torch.where(unstable, torch.zeros_like(d), d)
        for flat in torch.topk(dm.flatten(), min(4, n_out_bad)).indices.tolist():
            n_, b_, c_, y_, x_ = [int(v) for v in torch.unravel_index(torch.tensor(flat), dm.shape)]
            offenders.append({"n": n_, "c": c_, "y": y_, "x": x_, "d": dm.flatten()[flat].item(), "ref_wz": aux["wz"][n_, b_, 0, y_, x_].item(),
                              "ref_wz_minus_1_ulps": (aux["wz"][n_, b_, 0, y_, x_].item() - 1.0) / 2.0 ** -23})
    a = torch.where(unstable, r_rgb, rgb)
    p_stable = psnr(a.cpu(), r_rgb.cpu())
    p_all = psnr(rgb.cpu(), r_rgb.cpu())
    _report(test="fullsize_vs_reference_gpu", workload=workload, hr=[HH, WW], timestamps=N, flow_max_abs=d_flow, rgb_max_abs_outside_mask=d_out,
            values_over_tol_outside_mask=n_out_bad, excluded_fraction=frac, excluded_fraction_count_splat=frac_count, excluded_fraction_wz_equality=frac_eq,
            excluded_values=n_in, excluded_values_over_tol=n_in_bad, rgb_max_abs_inside_mask=max_in,
            reference_vs_reference_rgb_max_abs=ref_self_max, reference_vs_reference_values_over_tol=ref_self_over,
            psnr_stable_db=p_stable, psnr_all_pixels_db=p_all, offenders=offenders)
    assert d_flow < FLOW_TOL, d_flow
    assert frac_count < 0.03, frac_count
    assert d_out < TOL, d_out
    assert p_stable > PSNR_MIN
    # every pixel included (no mask): the frames as a user sees them still agree to far better than 0.01 dB-vs-GT needs
    assert p_all > 50.0, p_all


def test_adobe_full_size_all_timestamps_vs_reference_gpu_path():
    """BASELINE config 1 (180x320 -> 720x1280, 7 timestamps): the headline workload itself."""
    _fullsize("adobe240_x4_t8", 180, 320, 720, 1280, [k / 8 for k in range(1, 8)], seed=7, chunk=2)


def test_x3p5_full_size_vs_reference_gpu_path():
    """BASELINE config 3 (x3.5 space -> 630x1120, x12 time: 11 timestamps, two timestamp groups)."""
    _fullsize("adobe240_x3p5_t12", 180, 320, 630, 1120, [k / 12 for k in range(1, 12)], seed=5, chunk=2)


def test_uhd_quarter_crop_vs_reference_gpu_path():
    """A quarter-area crop of BASELINE config 4 (LR 270x480 -> 1080x1920, x4), three timestamps."""
    _fullsize("uhd4k_x4_t8 (quarter crop)", 270, 480, 1080, 1920, [0.125, 0.5, 0.875], seed=9, chunk=1)


def _sensitivity(g, HH, WW):
    """How far the REFERENCE ITSELF moves when every weight moves by one ulp (CPU oracle): the conditioning of the fp32
    problem, which bounds what any other evaluation order can be asked to reproduce."""
    p = hot_params(g)
    rgb, flow, inter = decoder_ref.decode(g["feat"], g["flow_feat"], g["residual"], g["target_t"], HH, WW, p, return_intermediates=True)
    p2 = {k: (torch.nextafter(v, v * 2) if k != "alpha" else v) for k, v in p.items()}
    rgb2, flow2 = decoder_ref.decode(g["feat"], g["flow_feat"], g["residual"], g["target_t"], HH, WW, p2)
    B, N = g["target_t"].shape
    unstable = decoder_ref.count_unstable_mask(inter["flow_hr"], B, N).expand_as(rgb)
    return (flow2 - flow).abs().max().item(), (rgb2 - rgb).abs()[~unstable].max().item(), unstable


@pytest.mark.parametrize("precision", ["f16x3", "fp32"])
@pytest.mark.parametrize("case", ["decoder_alpha_m1", "decoder_alpha_p05", "decoder_gain2", "decoder_gain4"])
def test_weight_regimes_vs_reference_golden(case, precision):
    """alpha = -1 (exp(z) in (0.9, 1]), alpha = +0.5 (exp(z) > 1: the max splat leaves its initial 1.0 -- zmax up to 1.34 --
    Ours.py:794, 834, softsplat_max_cp.py:254) and SIREN hidden gains 2 and 4 (sine arguments of tens of radians).  The
    gates are north_star's, widened only where the reference's own one-ulp sensitivity exceeds them (gain 4)."""
    from motif_b200.decoder import SpaceTimeDecoder

    g = load_golden(case)
    HH, WW = [int(v) for v in g["hr_size"]]
    dec = SpaceTimeDecoder(hot_params(g), device="cuda", precision=precision)
    rgb, flow = dec.decode(g["feat"].cuda(), g["flow_feat"].cuda(), g["residual"].cuda(), g["target_t"], (HH, WW))
    s_flow, s_rgb, unstable = _sensitivity(g, HH, WW)
    d_flow = (flow.cpu() - g["flow_out"]).abs().max().item()
    d = (rgb.cpu() - g["out"]).abs()
    d_rgb = d[~unstable].max().item()
    _report(test="weight_regime", case=case, precision=precision, alpha=float(g["alpha"][0]), flow_max_abs=d_flow, rgb_max_abs_outside_mask=d_rgb,
            reference_one_ulp_flow=s_flow, reference_one_ulp_rgb=s_rgb, excluded_fraction=unstable.float().mean().item())
    # K one-ulp sensitivities of the reference: the exact-fp32 CUDA-core path stays within 1, f16x3 (22-bit operands, MUFU.SIN
    # after a single-step range reduction) within ~6 at gain 4 -- a regime where NO evaluation order of the fp32 function,
    # the reference's own included, reproduces another to 1e-3 (DESIGN.md section 3 quotes the measured numbers)
    k = 8.0 if precision == "f16x3" else 4.0
    assert d_flow < max(FLOW_TOL, k * s_flow), (d_flow, s_flow)
    assert d_rgb < max(TOL, k * s_rgb), (d_rgb, s_rgb)
    if case == "decoder_alpha_p05":  # the fixture does exercise zmax > 1: pretending the max splat is identically 1 must fail
        dbg = dec.decode(g["feat"].cuda(), g["flow_feat"].cuda(), g["residual"].cuda(), g["target_t"], (HH, WW), debug_synth_in=True)[2] if precision == "fp32" else None
        if dbg is not None:
            assert dbg[:, 130].max().item() > 1.2


def test_two_scales_to_one_hr_size_on_one_decoder():
    """Arbitrary-scale evaluation (LQ_size = GT_size // scale): one decoder object, the same HR size reached from two LR
    sizes back to back.  The workspace's armed accumulators must not be mistaken for armed when only (H, W) changed
    (the arming magic ignored the LR size while the armed region's offset depended on it)."""
    from motif_b200 import synthetic
    from motif_b200.decoder import SpaceTimeDecoder

    params = decoder_ref.random_params(seed=5, **decoder_ref.REALISTIC)
    HH, WW = 96, 128
    tt = torch.tensor([[0.25, 0.75]])
    dec = SpaceTimeDecoder(params, device="cuda", precision="f16x3")
    outs = {}
    for H, W in ((48, 64), (24, 32), (48, 64), (32, 32), (24, 32)):
        lat = [t.cuda() for t in synthetic.synthetic_latents(1, H, W, seed=H)]
        rgb, flow = dec.decode(*lat, tt, (HH, WW))
        fresh, fflow = SpaceTimeDecoder(params, device="cuda", precision="f16x3").decode(*lat, tt, (HH, WW))
        assert torch.equal(flow, fflow), (H, W)
        assert (rgb - fresh).abs().max().item() < 1e-5, (H, W)
        if (H, W) in outs:
            assert (rgb - outs[(H, W)]).abs().max().item() < 1e-5
        outs[(H, W)] = rgb
    r_rgb, r_flow = decoder_ref.decode(*[t.cpu() for t in synthetic.synthetic_latents(1, 24, 32, seed=24)], tt, HH, WW, params)
    un = decoder_ref.count_unstable_mask(r_flow * 20.0 * (HH / 24), 1, 2).expand_as(r_rgb)
    assert (outs[(24, 32)].cpu() - r_rgb).abs()[~un].max().item() < TOL
