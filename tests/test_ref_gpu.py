"""GPU parity against the reference's OWN CUDA kernels running on the same B200.

``oracle/build_ref_gpu.py`` compiles the reference's kernel strings (unmodified, expanded by the reference's own
pre-processor) with nvcc for sm_100a into ``oracle/_ref/ref_gpu_sm100a.so``.  These tests run them beside the product
kernels on identical inputs at BASELINE sizes, where the CPU oracle would take minutes: count and max splat bit-exact
(integer / order-independent), sum splat within the float-atomic ordering noise of the reference itself (measured:
the reference against a second run of itself), correlation within summation-order noise."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref():
    from oracle import build_ref_gpu

    if build_ref_gpu.load() is None:
        pytest.skip("oracle/_ref/ref_gpu_sm100a.so not built (reference checkout absent at build time)")
    return build_ref_gpu


def _flow(n, h, w, cell, mag, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    low = torch.randn(n, 2, max(h // cell, 2), max(w // cell, 2), device="cuda", generator=g) * mag
    return torch.nn.functional.interpolate(low, size=(h, w), mode="bilinear", align_corners=False).contiguous()


@pytest.mark.parametrize("tag,n,c,h,w,cell", [("small", 2, 5, 37, 52, 4), ("vimeo", 1, 130, 256, 448, 32), ("adobe", 1, 130, 720, 1280, 64), ("adobe", 1, 130, 720, 1280, 16)])
def test_softmax_splat_vs_reference_kernel_on_device(tag, n, c, h, w, cell):
    from motif_b200.softsplat_cp import FunctionSoftsplat

    ref = _ref()
    g = torch.Generator(device="cuda").manual_seed(1)
    inp = torch.randn(n, c, h, w, device="cuda", generator=g)
    metric = -torch.rand(n, 1, h, w, device="cuda", generator=g)
    flow = _flow(n, h, w, cell, 6.0, 2)
    out, norm = FunctionSoftsplat(inp, flow, metric, "softmax")
    # the reference wrapper forms [in * exp(metric) | exp(metric)] with torch, then launches its kernel (softsplat_cp.py:332-346)
    e = metric.exp()
    ref_in = torch.cat([inp * e, e], 1).contiguous()
    r1 = ref.splat("sum", tag, ref_in, flow)
    r2 = ref.splat("sum", tag, ref_in, flow)
    noise = (r1 - r2).abs().max().item()            # the reference's own run-to-run atomic-order noise
    full = torch.cat([out, norm], 1)
    err = (full - r1).abs().max().item()
    assert err <= max(4.0 * noise, 2e-5), (err, noise)
    assert err < 1e-3                                # north_star tolerance


@pytest.mark.parametrize("tag,n,h,w", [("small", 2, 37, 52), ("vimeo", 1, 256, 448), ("adobe", 1, 720, 1280)])
def test_max_and_count_splat_bit_exact_vs_reference_kernel_on_device(tag, n, h, w):
    from motif_b200 import softsplat_count_cp, softsplat_max_cp

    ref = _ref()
    g = torch.Generator(device="cuda").manual_seed(3)
    z = (torch.randn(n, 1, h, w, device="cuda", generator=g) * 0.5).exp().contiguous()   # candidates on both sides of 1.0
    flow = _flow(n, h, w, 16, 5.0, 4)
    assert torch.equal(softsplat_max_cp.FunctionSoftsplat(z, flow), ref.splat("max", tag, z, flow))
    ones = torch.ones_like(z)                        # the count wrapper feeds ones (softsplat_count_cp.py:163-165)
    assert torch.equal(softsplat_count_cp.FunctionSoftsplat(z, flow), ref.splat("count", tag, ones, flow))


@pytest.mark.parametrize("tag,b,c,h,w", [("small", 2, 16, 12, 20), ("l6", 1, 196, 12, 20), ("l3", 1, 64, 96, 160), ("l2", 1, 32, 192, 320)])
def test_correlation_vs_reference_kernel_on_device(tag, b, c, h, w):
    from motif_b200.correlation import FunctionCorrelation

    ref = _ref()
    g = torch.Generator(device="cuda").manual_seed(5)
    first = torch.randn(b, c, h, w, device="cuda", generator=g)
    second = torch.randn(b, c, h, w, device="cuda", generator=g)
    out = FunctionCorrelation(first, second)
    r = ref.correlation(tag, first, second)
    assert out.shape == r.shape
    assert (out - r).abs().max().item() < 2e-6 * (c ** 0.5) + 1e-6   # fp32 summation order over c terms of O(1), divided by c
