"""GPU: the tcgen05 building blocks (TMEM staging, swizzled smem weight image, 3xTF32 MMA) on one tile."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(x, w, terms):
    from motif_b200 import _lib

    lib = _lib.load()
    d = torch.full((128, 64), float("nan"), device="cuda")
    scratch = torch.empty(32768 // 4, device="cuda")
    rc = lib.motif_tc_selftest(x.data_ptr(), w.data_ptr(), d.data_ptr(), scratch.data_ptr(), terms, _lib.current_stream_ptr())
    _lib.check(rc, "motif_tc_selftest")
    torch.cuda.synchronize()
    return d


def test_tile_layout_exact_on_small_integers():
    """Small integers are exact in TF32: any descriptor / swizzle / lane mapping error shows up as a wrong entry."""
    torch.manual_seed(0)
    x = torch.randint(-8, 9, (128, 64), device="cuda").float()
    w = torch.randint(-8, 9, (64, 64), device="cuda").float()
    for terms in (1, 3):
        d = _run(x, w, terms)
        assert torch.equal(d, x @ w.t()), terms


def test_3xtf32_is_fp32_accurate():
    torch.manual_seed(1)
    x = torch.randn(128, 64, device="cuda")
    w = torch.randn(64, 64, device="cuda") * 0.1
    ref = (x.double() @ w.double().t())
    e1 = (_run(x, w, 1).double() - ref).abs().max().item()
    e3 = (_run(x, w, 3).double() - ref).abs().max().item()
    efp32 = ((x @ w.t()).double() - ref).abs().max().item()
    print(f"max|err| vs fp64: 1xTF32 {e1:.3e}  3xTF32 {e3:.3e}  torch fp32 {efp32:.3e}")
    assert e1 < 5e-3 and e3 < 5e-6
    assert e3 < 4 * max(efp32, 2e-7)
