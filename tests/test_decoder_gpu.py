"""GPU parity: the space-time local implicit decoder (Ours.py:659-858) through the C ABI."""
import pytest
import torch

from conftest import hot_params, load_golden, psnr
from oracle import decoder_ref

pytestmark = pytest.mark.gpu

TOL = 1e-3        # north_star: decoder outputs within 1e-3 max-abs ...
PSNR_MIN = 60.0   # ... and 0.01 dB PSNR: PSNR(new, reference) >= 60 dB keeps any PSNR-vs-GT within 0.01 dB
FLOW_TOL = 2e-6   # raw flow units (HR pixels / (20 * scale)): < 2e-4 HR px at x4, below the stability margin


def _compare_frames(rgb, ref_rgb, ref_flow_out, scale_hh_over_h, B, N):
    """max-abs over the pixels where the reference function is continuous (see
    oracle.decoder_ref.count_unstable_mask), PSNR over the same pixels, and the excluded fraction."""
    unstable = decoder_ref.count_unstable_mask(ref_flow_out * 20.0 * scale_hh_over_h, B, N).expand_as(ref_rgb)
    frac = unstable.float().mean().item()
    assert frac < 0.03, frac
    d = (rgb - ref_rgb).abs()
    a = torch.where(unstable, ref_rgb, rgb)
    return d[~unstable].max().item(), psnr(a, ref_rgb), frac

# (H, W, HH, WW): Vimeo x4, Adobe x4, x3.5 (round(H*3.5)), 4K x4, plus awkward ratios with index ties
GEOMS = [(64, 112, 256, 448), (180, 320, 720, 1280), (180, 320, 630, 1120), (540, 960, 2160, 3840),
         (16, 20, 56, 70), (7, 9, 20, 31), (5, 5, 5, 5), (12, 16, 18, 24), (3, 4, 96, 100)]


def _decoder(params, precision):
    from motif_b200.decoder import SpaceTimeDecoder

    return SpaceTimeDecoder(params, device="cuda", precision=precision)


@pytest.mark.parametrize("geom", GEOMS)
def test_query_geometry_bit_exact(geom):
    """Nearest-latent index, shifted coordinate and relative coordinate: integer / bit equality."""
    H, W, HH, WW = geom
    dec = _decoder(decoder_ref.random_params(0), "fp32")
    iy, ix, coord, rel = dec.query_geometry(H, W, HH, WW)
    ref = decoder_ref.query_geometry(H, W, HH, WW)
    assert torch.equal(iy.cpu().long(), ref["iy"])
    assert torch.equal(ix.cpu().long(), ref["ix"])
    assert torch.equal(coord.cpu(), ref["coord_"])
    assert torch.equal(rel.cpu(), ref["rel"])


def test_coord_sequence_host_matches_make_coord():
    from motif_b200.decoder import coord_sequence

    for n in (5, 56, 720, 1120, 3840):
        assert torch.equal(coord_sequence(n), decoder_ref.make_coord((n,)).view(-1))


@pytest.mark.parametrize("precision", ["fp32", "tf32x3", "f16x3"])
@pytest.mark.parametrize("case", ["decoder_raft", "decoder_x4", "decoder_x3p5_b2"])
def test_decoder_vs_reference_golden(case, precision):
    g = load_golden(case)
    HH, WW = [int(v) for v in g["hr_size"]]
    dec = _decoder(hot_params(g), precision)
    rgb, flow = dec.decode(g["feat"].cuda(), g["flow_feat"].cuda(), g["residual"].cuda(), g["target_t"], (HH, WW))
    assert rgb.shape == g["out"].shape and flow.shape == g["flow_out"].shape
    d_flow = (flow.cpu() - g["flow_out"]).abs().max().item()
    assert d_flow < FLOW_TOL, d_flow
    B, N = g["target_t"].shape
    d_rgb, p, _ = _compare_frames(rgb.cpu(), g["out"], g["flow_out"], HH / g["feat"].shape[-2], B, N)
    assert d_rgb < TOL, d_rgb
    assert p > PSNR_MIN


@pytest.mark.parametrize("precision", ["fp32", "f16x3"])
def test_local_ensemble_vs_reference_golden(precision):
    """LunaTokis.local_ensemble = True (Ours.py:660-663, 754-764; off as shipped): four shifted latents blended by the
    diagonally swapped area weights.  Golden = the reference's own forward with the flag set (oracle/make_golden.py)."""
    from motif_b200.decoder import SpaceTimeDecoder

    g = load_golden("decoder_ens_x3")
    HH, WW = [int(v) for v in g["hr_size"]]
    dec = SpaceTimeDecoder(hot_params(g), device="cuda", precision=precision, local_ensemble=True)
    rgb, flow = dec.decode(g["feat"].cuda(), g["flow_feat"].cuda(), g["residual"].cuda(), g["target_t"], (HH, WW))
    assert rgb.shape == g["out"].shape and flow.shape == g["flow_out"].shape
    d_flow = (flow.cpu() - g["flow_out"]).abs().max().item()
    assert d_flow < FLOW_TOL, d_flow
    B, N = g["target_t"].shape
    d_rgb, p, _ = _compare_frames(rgb.cpu(), g["out"], g["flow_out"], HH / g["feat"].shape[-2], B, N)
    assert d_rgb < TOL, d_rgb
    assert p > PSNR_MIN
    # the flag changes the function: the single-latent decode of the same inputs is far from this golden
    plain, _ = SpaceTimeDecoder(hot_params(g), device="cuda", precision=precision).decode(
        g["feat"].cuda(), g["flow_feat"].cuda(), g["residual"].cuda(), g["target_t"], (HH, WW))
    assert (plain.cpu() - g["out"]).abs().max().item() > 5e-3


def test_weight_images_are_reused_and_two_decoders_do_not_mix():
    """Second call on the same workspace skips the weight repacking (motif_decode_t.weights_ready); a second decoder
    with other weights in between re-uploads its own output-layer constants (the constant bank is per device)."""
    from motif_b200 import synthetic

    B, H, W, HH, WW = 1, 12, 16, 48, 64
    feat, ff, res = [t.cuda() for t in synthetic.synthetic_latents(B, H, W, seed=3)]
    tt = torch.tensor([[0.25, 0.75]])
    dec_a = _decoder(synthetic.synthetic_params(seed=3), "f16x3")
    dec_b = _decoder(synthetic.synthetic_params(seed=4), "f16x3")
    a1, fa1 = dec_a.decode(feat, ff, res, tt, (HH, WW))
    b1, _ = dec_b.decode(feat, ff, res, tt, (HH, WW))
    a2, fa2 = dec_a.decode(feat, ff, res, tt, (HH, WW))      # weights_ready = 1, constants of dec_b in the bank
    b2, _ = dec_b.decode(feat, ff, res, tt, (HH, WW))
    a3, _ = dec_a.decode(feat, ff, res, tt, (HH, WW), precision="fp32")   # another layout overwrites the images
    a4, _ = dec_a.decode(feat, ff, res, tt, (HH, WW))                     # ... so they are rebuilt
    assert torch.equal(fa1, fa2)
    assert (a1 - a2).abs().max().item() < 1e-6 and (b1 - b2).abs().max().item() < 1e-6 and (a1 - a4).abs().max().item() < 1e-6
    assert (a1 - b1).abs().max().item() > 1e-3
    assert (a1 - a3).abs().max().item() < 1e-2  # same function in exact fp32 (unstable-count pixels aside)


def test_local_ensemble_is_refused_by_the_first_generation_path():
    from motif_b200.decoder import SpaceTimeDecoder

    with pytest.raises(NotImplementedError):
        SpaceTimeDecoder(decoder_ref.random_params(0), device="cuda", precision="tf32x3", local_ensemble=True)


@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
def test_decoder_stages_vs_oracle(precision):
    """Seeded O(1)-scaled weights; compares the blended splat + reliability features (the synth_net input,
    Ours.py:839-844) and the final frames against the oracle."""
    gen = torch.Generator().manual_seed(3)
    B, N, H, W, HH, WW = 1, 2, 12, 20, 42, 70
    feat = torch.randn(2 * B, 64, H, W, generator=gen) * 0.3
    ff = torch.randn(2 * B, 64, H, W, generator=gen) * 0.3
    res = torch.randn(B, 64, H, W, generator=gen) * 0.3
    tt = torch.tensor([[0.25, 0.8]])
    params = decoder_ref.random_params(seed=5, **decoder_ref.REALISTIC)
    r_rgb, r_flow, inter = decoder_ref.decode(feat, ff, res, tt, HH, WW, params, return_intermediates=True)
    dec = _decoder(params, precision)
    rgb, flow, synth_in = dec.decode(feat.cuda(), ff.cuda(), res.cuda(), tt, (HH, WW), debug_synth_in=True)
    assert (flow.cpu() - r_flow).abs().max().item() < FLOW_TOL
    unstable = decoder_ref.count_unstable_mask(inter["flow_hr"], B, N)          # [N,B,1,HH,WW]
    st = ~unstable.permute(1, 0, 2, 3, 4).reshape(B * N, 1, HH, WW)
    d_in = (synth_in.cpu() - inter["synth_in"]).abs()
    assert d_in[:, 130:133][st.expand(-1, 3, -1, -1)].max().item() < 1e-4      # zmax, count/16, wz/count
    cnt_new, cnt_ref = synth_in.cpu()[:, 131:132], inter["synth_in"][:, 131:132]
    assert torch.equal(cnt_new[st], cnt_ref[st])                                # count is exact where it is stable
    assert d_in[st.expand_as(d_in)].max().item() < TOL
    d_rgb, p, _ = _compare_frames(rgb.cpu(), r_rgb, r_flow, HH / H, B, N)
    assert d_rgb < TOL and p > PSNR_MIN, (d_rgb, p)


def test_f16x3_stages_vs_oracle():
    """Default arithmetic (fp16 two-piece split, layer 0 folded through the splat): the synth_net layer-0
    pre-activation -- the quantity that replaces the 198-channel input -- against the oracle's
    synth_in @ W0^T + b0 (fp64), then the frames."""
    gen = torch.Generator().manual_seed(3)
    B, N, H, W, HH, WW = 1, 2, 12, 20, 42, 70
    feat = torch.randn(2 * B, 64, H, W, generator=gen) * 0.3
    ff = torch.randn(2 * B, 64, H, W, generator=gen) * 0.3
    res = torch.randn(B, 64, H, W, generator=gen) * 0.3
    tt = torch.tensor([[0.25, 0.8]])
    params = decoder_ref.random_params(seed=5, **decoder_ref.REALISTIC)
    r_rgb, r_flow, inter = decoder_ref.decode(feat, ff, res, tt, HH, WW, params, return_intermediates=True)
    dec = _decoder(params, "f16x3")
    rgb, flow, pre0 = dec.decode(feat.cuda(), ff.cuda(), res.cuda(), tt, (HH, WW), debug_pre0=True)
    assert (flow.cpu() - r_flow).abs().max().item() < FLOW_TOL
    w0 = params["synth_net.net.0.linear.weight"].double()
    b0 = params["synth_net.net.0.linear.bias"].double()
    ref_pre0 = torch.einsum("nchw,oc->nohw", inter["synth_in"].double(), w0) + b0.view(1, -1, 1, 1)
    unstable = decoder_ref.count_unstable_mask(inter["flow_hr"], B, N)          # [N,B,1,HH,WW]
    st = ~unstable.permute(1, 0, 2, 3, 4).reshape(B * N, 1, HH, WW)
    d = (pre0.cpu().double() - ref_pre0).abs()
    scale = ref_pre0.abs().max().item()
    assert d[st.expand_as(d)].max().item() < 2e-5 * max(scale, 1.0), (d[st.expand_as(d)].max().item(), scale)
    d_rgb, p, _ = _compare_frames(rgb.cpu(), r_rgb, r_flow, HH / H, B, N)
    assert d_rgb < TOL and p > PSNR_MIN, (d_rgb, p)


def test_f16x3_overfull_lists_spill():
    """A contracting flow field piles > 16 contributions on some destinations: the spill accumulator path must
    agree with the exact-fp32 CUDA-core path (which scatters with float atomics)."""
    gen = torch.Generator().manual_seed(21)
    B, H, W, HH, WW = 1, 8, 8, 64, 64
    feat = torch.randn(2 * B, 64, H, W, generator=gen) * 0.3
    ff = torch.randn(2 * B, 64, H, W, generator=gen) * 2.0      # strong, varied flows
    res = torch.randn(B, 64, H, W, generator=gen) * 0.3
    tt = torch.tensor([[0.5]])
    params = decoder_ref.random_params(seed=9, **decoder_ref.REALISTIC)
    a, fa = _decoder(params, "f16x3").decode(feat.cuda(), ff.cuda(), res.cuda(), tt, (HH, WW))
    r_rgb, r_flow, inter = decoder_ref.decode(feat, ff, res, tt, HH, WW, params, return_intermediates=True)
    cnt = inter["synth_in"][:, 131] * 16.0
    assert cnt.max().item() > 16, cnt.max().item()       # the case really overfills some lists
    assert (fa.cpu() - r_flow).abs().max().item() < FLOW_TOL
    unstable = decoder_ref.count_unstable_mask(inter["flow_hr"], B, 1).expand_as(r_rgb)
    d = (a.cpu() - r_rgb).abs()
    assert d[~unstable].max().item() < TOL, d[~unstable].max().item()
    over = (cnt > 16).view(1, B, 1, HH, WW).expand_as(r_rgb) & ~unstable
    assert over.any() and d[over].max().item() < TOL


@pytest.mark.parametrize("precision", ["tf32x3", "f16x3"])
def test_timestamp_range_equals_full_decode(precision):
    """Sharding contract: decoding timestamps [a,b) gives the same frames as the full decode."""
    g = load_golden("decoder_x4")
    HH, WW = [int(v) for v in g["hr_size"]]
    dec = _decoder(hot_params(g), precision)
    args = (g["feat"].cuda(), g["flow_feat"].cuda(), g["residual"].cuda(), g["target_t"], (HH, WW))
    full, _ = dec.decode(*args)
    part, _ = dec.decode(*args, n_range=(1, 3))
    assert (part[1:3] - full[1:3]).abs().max().item() < 1e-5


@pytest.mark.parametrize("precision", ["tf32x3", "f16x3"])
def test_vimeo_config_vs_oracle_and_psnr(precision):
    """BASELINE config 0 (64x112 -> 256x448, t=0.5): whole hot path against the oracle."""
    gen = torch.Generator().manual_seed(11)
    H, W, HH, WW = 64, 112, 256, 448
    low = torch.randn(5, 64, H // 4, W // 4, generator=gen)
    lat = torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear") * 0.4
    feat, ff, res = lat[0:2].contiguous(), lat[2:4].contiguous(), lat[4:5].contiguous()
    tt = torch.tensor([[0.5]])
    params = decoder_ref.random_params(seed=2, **decoder_ref.REALISTIC)
    r_rgb, r_flow = decoder_ref.decode(feat, ff, res, tt, HH, WW, params)
    rgb, flow = _decoder(params, precision).decode(feat.cuda(), ff.cuda(), res.cuda(), tt, (HH, WW))
    assert (flow.cpu() - r_flow).abs().max().item() < FLOW_TOL
    d_rgb, p, _ = _compare_frames(rgb.cpu(), r_rgb, r_flow, HH / H, 1, 1)
    assert d_rgb < TOL and p > PSNR_MIN, (d_rgb, p)


@pytest.mark.parametrize("precision", ["tf32x3", "f16x3"])
def test_adobe_full_size_properties(precision):
    """BASELINE config 1 size (180x320 -> 720x1280, 7 timestamps): tensor path vs the exact-fp32
    CUDA-core path on the device, finite and clamped output, determinism of shape/ordering."""
    gen = torch.Generator().manual_seed(7)
    H, W, HH, WW = 180, 320, 720, 1280
    low = torch.randn(5, 64, H // 4, W // 4, generator=gen)
    lat = torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear") * 0.4
    feat, ff, res = lat[0:2].contiguous().cuda(), lat[2:4].contiguous().cuda(), lat[4:5].contiguous().cuda()
    tt = torch.tensor([[k / 8 for k in range(1, 8)]])
    params = decoder_ref.random_params(seed=2, **decoder_ref.REALISTIC)
    a, fa = _decoder(params, precision).decode(feat, ff, res, tt, (HH, WW))
    b, fb = _decoder(params, "fp32").decode(feat, ff, res, tt, (HH, WW), n_range=(0, 2))
    assert a.shape == (7, 1, 3, HH, WW) and torch.isfinite(a).all()
    assert a.min().item() >= 0.0 and a.max().item() <= 1.0
    assert (fa[:2] - fb[:2]).abs().max().item() < FLOW_TOL
    # on-device comparison of the two arithmetic paths: a handful of count-unstable pixels may differ by more
    d = (a[:2] - b[:2]).abs()
    assert (d > TOL).float().mean().item() < 2e-3
    assert torch.quantile(d.flatten()[::7].float(), 0.999).item() < TOL


@pytest.mark.parametrize("workload,n_take", [("adobe240_x3p5_t12", 2), ("uhd4k_x4_t8", 1)])
def test_other_baseline_configs_full_size(workload, n_take):
    """BASELINE configs 3 (x3.5 space, round(H * 3.5); x12 time) and 4 (4K output) at FULL size: the f16x3 path against
    the exact-fp32 CUDA-core path on the device for the first timestamps, finite / clamped output for all decoded ones."""
    from motif_b200 import synthetic

    H, W, HH, WW, times = synthetic.WORKLOADS[workload]
    feat, ff, res = [t.cuda() for t in synthetic.synthetic_latents(1, H, W, seed=5)]
    params = decoder_ref.random_params(seed=2, **decoder_ref.REALISTIC)
    n_f16 = min(len(times), 3 if HH > 2000 else len(times))
    tt = torch.tensor([times[:n_f16]])
    a, fa = _decoder(params, "f16x3").decode(feat, ff, res, tt, (HH, WW))
    assert a.shape == (n_f16, 1, 3, HH, WW) and torch.isfinite(a).all()
    assert a.min().item() >= 0.0 and a.max().item() <= 1.0
    b, fb = _decoder(params, "fp32").decode(feat, ff, res, tt, (HH, WW), n_range=(0, n_take))
    # flow_out is [2 * B * N, 2, HH, WW] with index r * N + n: compare the timestamps the fp32 path decoded
    N = n_f16
    for r in range(2):
        assert (fa[r * N:r * N + n_take] - fb[r * N:r * N + n_take]).abs().max().item() < FLOW_TOL
    d = (a[:n_take] - b[:n_take]).abs()
    assert (d > TOL).float().mean().item() < 2e-3
    assert torch.quantile(d.flatten()[::97].float(), 0.999).item() < TOL


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process (the reference's DataParallel case)")
def test_two_devices_in_one_process():
    """DataParallel drives several GPUs from one process (VideoSR_base_model.py:36): kernel attributes and the constant
    bank are per device, so the second device must work after the first has initialised the library."""
    from motif_b200 import synthetic
    from motif_b200.correlation import FunctionCorrelation
    from motif_b200.decoder import SpaceTimeDecoder
    from motif_b200.softsplat_cp import FunctionSoftsplat

    B, H, W, HH, WW = 1, 12, 16, 48, 64
    lat = synthetic.synthetic_latents(B, H, W, seed=3)
    params = synthetic.synthetic_params(seed=3)
    tt = torch.tensor([[0.25, 0.75]])
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        with torch.cuda.device(dev):
            dec = SpaceTimeDecoder(params, device=dev)
            rgb, flow = dec.decode(*[t.to(dev) for t in lat], tt, (HH, WW))
            x = torch.arange(2 * 130 * 40 * 64, dtype=torch.float32, device=dev).reshape(2, 130, 40, 64).sin()
            fl = torch.full((2, 2, 40, 64), 1.25, device=dev)
            o, n = FunctionSoftsplat(x, fl, -x[:, :1].abs(), "softmax")
            c = FunctionCorrelation(x[:, :32].contiguous(), x[:, 32:64].contiguous())
            outs.append([t.cpu() for t in (rgb, flow, o, n, c)])
    names = ("rgb", "flow", "splat", "norm", "corr")
    for name, a, b in zip(names, *outs):
        if name == "rgb":  # the decoder's destination lists are filled in atomic order: fp32 sums differ in the last bits run to run
            assert (a - b).abs().max().item() < 1e-5
        else:
            assert torch.equal(a, b), name


def test_state_dict_loader_accepts_full_checkpoint_layout():
    from motif_b200.decoder import SpaceTimeDecoder

    params = decoder_ref.random_params(0)
    sd = {"module." + k: v for k, v in params.items()}
    sd["module.encoder.dummy"] = torch.zeros(3)
    dec = SpaceTimeDecoder.from_state_dict(sd, device="cuda")
    assert dec.alpha == -20.0
    with pytest.raises(KeyError):
        SpaceTimeDecoder({k: v for k, v in params.items() if "synth_net.net.4" not in k})


def test_f16x3_timestamp_groups_and_rearmed_workspace():
    """x12-time shape (11 timestamps > one group of 8, BASELINE config 3): the grouped three-phase decode against the
    oracle; then the SAME decoder object again (accumulators re-armed by their consumer, no clear), after a decode of
    another precision and of another geometry scribbled over the shared workspace."""
    gen = torch.Generator().manual_seed(13)
    B, H, W, HH, WW = 1, 10, 12, 35, 42          # x3.5 space
    feat = torch.randn(2 * B, 64, H, W, generator=gen) * 0.3
    ff = torch.randn(2 * B, 64, H, W, generator=gen) * 0.3
    res = torch.randn(B, 64, H, W, generator=gen) * 0.3
    tt = torch.tensor([[k / 12 for k in range(1, 12)]])
    N = tt.shape[1]
    params = decoder_ref.random_params(seed=5, **decoder_ref.REALISTIC)
    r_rgb, r_flow = decoder_ref.decode(feat, ff, res, tt, HH, WW, params)
    dec = _decoder(params, "f16x3")
    args = (feat.cuda(), ff.cuda(), res.cuda(), tt, (HH, WW))
    rgb1, flow1 = dec.decode(*args)
    assert (flow1.cpu() - r_flow).abs().max().item() < FLOW_TOL
    d_rgb, p, _ = _compare_frames(rgb1.cpu(), r_rgb, r_flow, HH / H, B, N)
    assert d_rgb < TOL and p > PSNR_MIN, (d_rgb, p)
    rgb2, _ = dec.decode(*args)                                   # armed workspace: must be bit-identical
    assert (rgb1 - rgb2).abs().max().item() < 1e-5            # list order (atomic slot order) may differ: not bit-identical
    dec.decode(*args, precision="tf32x3")                         # another layout scribbles over the workspace
    dec.decode(feat.cuda()[..., :8, :8].contiguous(), ff.cuda()[..., :8, :8].contiguous(), res.cuda()[..., :8, :8].contiguous(),
               tt[:, :3], (16, 24))                               # another geometry
    rgb3, _ = dec.decode(*args)
    assert (rgb1 - rgb3).abs().max().item() < 1e-5
    part, _ = dec.decode(*args, n_range=(6, 10))                  # a range that straddles the group boundary
    assert (part[6:10] - rgb1[6:10]).abs().max().item() < 1e-5


def test_clip_stream_band_copy_in_pulls_only_the_bands_rows():
    """``ClipStream(band_copy_in=True)``: a band decode from host buffers copies only the LR rows its band and halo read (one strided
    copy per tensor, ``motif_memcpy2d_async``) -- the rest of the device latents stays whatever it was (poisoned here) -- and equals the
    resident band decode."""
    from motif_b200 import synthetic
    from motif_b200.clip_stream import ClipStream
    from motif_b200.decoder import SpaceTimeDecoder

    params = decoder_ref.random_params(seed=5, **decoder_ref.REALISTIC)
    dec = SpaceTimeDecoder(params, device="cuda", precision="f16x3")
    H, W, HH, WW = 24, 40, 96, 160
    hosts = [t.pin_memory() for t in synthetic.synthetic_latents(1, H, W, seed=4)]
    tt = torch.tensor([[0.25, 0.75]])
    band, halo = (32, 64), 16
    want, _ = dec.decode(*[t.cuda() for t in hosts], tt, (HH, WW), row_range=band, halo=halo)
    want = want[:, :, :, band[0]:band[1]].clone()
    cs = ClipStream(dec, depth=2, band_copy_in=True)
    out_h = torch.empty(2, 1, 3, band[1] - band[0], WW).pin_memory()
    for k in range(3):  # the slots are reused: poison them between clips
        for sl in cs._slots:
            if sl is not None:
                sl["flat"].fill_(float("nan"))
        cs.submit(*hosts, tt, (HH, WW), out_h, row_range=band, halo=halo)
        cs.synchronize()
        assert torch.isfinite(out_h).all()
        assert (out_h.cuda() - want).abs().max().item() < 1e-5
    lr0, lr1 = dec.lr_rows_of_band(H, HH, band, halo)
    lat = cs._slots[0]["lat"][0]
    assert torch.isnan(lat[:, :, :lr0]).all() and torch.isnan(lat[:, :, lr1:]).all() and torch.isfinite(lat[:, :, lr0:lr1]).all()


def test_nchw_latents_equal_the_packed_form():
    """f16x3 reads the reference's NCHW latents directly (motif_decode_t.latents_nchw); the pixel-major form the C ABI also accepts
    (motif_pack_latents) must give the same flows bit for bit and the same frames -- the per-LR-pixel tables see the same numbers in the
    same order -- for the whole frame and for a destination row band (only the band's LR rows are tabulated)."""
    from motif_b200 import synthetic
    from motif_b200.decoder import SpaceTimeDecoder

    params = decoder_ref.random_params(seed=5, **decoder_ref.REALISTIC)
    lat = [t.cuda() for t in synthetic.synthetic_latents(1, 24, 40, seed=3)]
    tt = torch.tensor([[0.25, 0.5, 0.75]])
    a = SpaceTimeDecoder(params, device="cuda", precision="f16x3")
    b = SpaceTimeDecoder(params, device="cuda", precision="f16x3")
    b.force_packed_latents = True
    for kw in ({}, {"row_range": (32, 64), "halo": 16}):
        ra, fa = a.decode(*lat, tt, (96, 160), **kw)
        rb, fb = b.decode(*lat, tt, (96, 160), **kw)
        rows = slice(*kw["row_range"]) if kw else slice(None)
        # the flows have no atomics on their way: bit-equal; the frames sum list entries in the order the binning atomics handed the
        # slots out, which differs from run to run in the last bits
        assert (ra[..., rows, :] - rb[..., rows, :]).abs().max().item() < 1e-5
        if not kw:
            assert torch.equal(fa, fb)


def test_clip_stream_matches_resident_decode():
    """Host-buffer front end (double-buffered copy-in / decode / copy-out): five different clips through two slots
    give the frames of the plain resident decode of each clip, in order."""
    from motif_b200.clip_stream import ClipStream

    gen = torch.Generator().manual_seed(17)
    B, H, W, HH, WW = 1, 12, 16, 48, 64
    params = decoder_ref.random_params(seed=5, **decoder_ref.REALISTIC)
    dec = _decoder(params, "f16x3")
    tt = torch.tensor([[0.25, 0.5, 0.75]])
    clips, outs = [], []
    for _ in range(5):
        clips.append([(torch.randn(s, generator=gen) * 0.3).pin_memory() for s in ((2 * B, 64, H, W), (2 * B, 64, H, W), (B, 64, H, W))])
        outs.append(torch.zeros(2, B, 3, HH, WW).pin_memory())
    cs = ClipStream(dec, depth=2)
    for c, o in zip(clips, outs):
        cs.submit(c[0], c[1], c[2], tt, (HH, WW), o, n_range=(1, 3))
    cs.synchronize()
    for c, o in zip(clips, outs):
        ref, _ = dec.decode(c[0].cuda(), c[1].cuda(), c[2].cuda(), tt, (HH, WW), return_flow=False)
        assert (o - ref[1:3].cpu()).abs().max().item() < 1e-5


def test_row_bands_assemble_to_the_full_decode():
    """SURVEY 8e, second sharding axis: destination row bands with a source halo.  Every band decoded on its own (as another
    rank would) reproduces the rows of the full decode; the reported max |flow_y| is the one of the full flow field; a halo
    that is too small is detected by the same number."""
    from motif_b200 import sharding, synthetic
    from motif_b200.decoder import SpaceTimeDecoder

    B, H, W, HH, WW = 1, 22, 20, 88, 80       # 5.5 aligned bands of 16 rows
    feat, ff, res = [t.cuda() for t in synthetic.synthetic_latents(B, H, W, seed=4)]
    params = decoder_ref.random_params(seed=5, **decoder_ref.REALISTIC)
    tt = torch.tensor([[0.2, 0.5, 0.9]])
    dec = SpaceTimeDecoder(params, device="cuda", precision="f16x3")
    full, flow = dec.decode(feat, ff, res, tt, (HH, WW))
    fy_max = (flow[:, 1].abs() * 20.0 * (HH / H)).max().item()
    halo = int(fy_max) + 3
    for world in (2, 3, 5):
        bands = sharding.partition_rows(HH, world)
        assert bands[0][0] == 0 and bands[-1][1] == HH and all(b % 16 == 0 or b == e for b, e in bands)
        out = torch.full_like(full, -1.0)
        seen = 0.0
        for r0, r1 in bands:
            if r1 == r0:
                continue
            stat = torch.zeros(64, device="cuda")
            part, pflow = SpaceTimeDecoder(params, device="cuda", precision="f16x3").decode(feat, ff, res, tt, (HH, WW), row_range=(r0, r1), halo=halo, flow_y_max=stat)
            out[..., r0:r1, :] = part[..., r0:r1, :]
            s0, s1 = max(r0 - halo, 0), min(r1 + halo, HH)
            assert torch.equal(pflow[..., s0:s1, :], flow[..., s0:s1, :])      # flows of the evaluated source rows are the same numbers
            seen = max(seen, sharding.check_halo(stat, halo))
        assert (out - full).abs().max().item() < 1e-5, world
        assert abs(seen - fy_max) < 1e-5 and seen < halo - 1
    # halo too small: the decode of a middle band misses contributions, and the check says so
    r0, r1 = sharding.partition_rows(HH, 3)[1]
    stat = torch.zeros(64, device="cuda")
    tiny = 1
    dec.decode(feat, ff, res, tt, (HH, WW), row_range=(r0, r1), halo=tiny, flow_y_max=stat)
    assert sharding.check_halo(stat, tiny) >= tiny - 1
    again, _ = dec.decode(feat, ff, res, tt, (HH, WW))                          # the workspace is intact after band decodes
    assert (again - full).abs().max().item() < 1e-5
    with pytest.raises(Exception):
        dec.decode(feat, ff, res, tt, (HH, WW), row_range=(8, 64), halo=halo)    # not a multiple of 16
    with pytest.raises(Exception):
        dec.decode(feat, ff, res, tt, (HH, WW), row_range=(0, 64), halo=halo, precision="fp32")
