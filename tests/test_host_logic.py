"""CPU: host-side logic -- coordinate tables, HR size rule, timestamp sharding, and the world_size-2
gloo run of the one exchange step of the path (latent broadcast) plus the frame gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from motif_b200 import sharding, synthetic
from motif_b200.decoder import coord_sequence, hr_size_from_scale
from oracle import decoder_ref


def test_coord_sequence_is_make_coord():
    for n in (1, 2, 5, 56, 70, 180, 630, 720, 1120, 1280, 2160, 3840):
        assert torch.equal(coord_sequence(n), decoder_ref.make_coord((n,)).view(-1))


def test_hr_size_rule():
    assert hr_size_from_scale(180, 320, 4) == (720, 1280)
    assert hr_size_from_scale(180, 320, 3.5) == (630, 1120)
    assert hr_size_from_scale(16, 20, 3.5) == (56, 70)
    assert hr_size_from_scale(64, 112, [[256], [448]]) == (256, 448)


@pytest.mark.parametrize("n,world", [(7, 1), (7, 2), (7, 4), (7, 8), (11, 4), (1, 8), (3, 3)])
def test_partition_covers_every_timestamp_once(n, world):
    parts = sharding.partition_timestamps(n, world)
    assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == n
    covered = [i for b, e in parts for i in range(b, e)]
    assert covered == list(range(n))
    sizes = [e - b for b, e in parts]
    assert max(sizes) - min(sizes) <= 1


def test_synthetic_params_have_checkpoint_layout():
    p = synthetic.synthetic_params(0)
    ref = decoder_ref.random_params(0)
    assert set(p) == set(ref)
    assert all(p[k].shape == ref[k].shape for k in p)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_ts):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, H, W = 1, 6, 8
        if rank == 0:
            feat, ff, res = synthetic.synthetic_latents(B, H, W, seed=3)
        else:
            feat, ff, res = torch.zeros(2 * B, 64, H, W), torch.zeros(2 * B, 64, H, W), torch.zeros(B, 64, H, W)
        feat, ff, res = sharding.broadcast_latents(feat, ff, res, src=0)
        want = synthetic.synthetic_latents(B, H, W, seed=3)
        assert torch.equal(feat, want[0]) and torch.equal(ff, want[1]) and torch.equal(res, want[2])
        # every rank "decodes" its timestamps: frame n is filled with n, then frames are gathered
        ranges = sharding.partition_timestamps(n_ts, world)
        b, e = ranges[rank]
        local = torch.stack([torch.full((B, 3, 4, 5), float(n)) for n in range(b, e)]) if e > b else torch.zeros(0, B, 3, 4, 5)
        full = sharding.gather_frames(local, ranges)
        assert full.shape == (n_ts, B, 3, 4, 5)
        assert torch.equal(full[:, 0, 0, 0, 0], torch.arange(n_ts, dtype=torch.float32))
    finally:
        dist.destroy_process_group()


def test_partition_rows_covers_the_frame_in_aligned_bands():
    from motif_b200 import sharding

    for n_rows, world in ((720, 8), (720, 7), (630, 4), (2160, 8), (20, 4), (8, 3), (5, 2)):
        bands = sharding.partition_rows(n_rows, world)
        assert len(bands) == world and bands[0][0] == 0 and bands[-1][1] == n_rows
        assert all(b <= e and (b % 16 == 0 or b == e) for b, e in bands)   # empty bands (more ranks than blocks) sit at the end
        assert all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))
        sizes = [e - b for b, e in bands if e > b]
        assert max(sizes) - min(sizes) <= 16 + 15        # balanced up to one aligned band (and the ragged last one)


def _halo_worker(rank, world, port):
    import torch.distributed as dist

    from motif_b200 import sharding

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        stat = torch.zeros(64)
        stat[3] = 5.5 if rank == 0 else 9.25
        assert sharding.check_halo(stat, 16) == 9.25     # the maximum over the ranks' bands, on every rank
    finally:
        dist.destroy_process_group()


def _slice_worker(rank, world, port):
    import torch.distributed as dist

    from motif_b200 import sharding, synthetic

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        hosts = synthetic.synthetic_latents(1, 5, 7, seed=8)  # every rank can read the clip; 11 200 floats do not divide by 3
        sizes = [t.numel() for t in hosts]
        per, parts = sharding.slice_plan(sizes, world, rank)
        flat = torch.full((per * world,), float("nan"))
        for i, a, b, dst in parts:
            flat[dst:dst + (b - a)].copy_(hosts[i].reshape(-1)[a:b])
        assert sum(b - a for _, a, b, _ in parts) <= per
        sharding.all_gather_slices(flat, per)
        assert torch.equal(flat[:sum(sizes)], torch.cat([t.reshape(-1) for t in hosts]))
    finally:
        dist.destroy_process_group()


def _exchange_worker(rank, world, port):
    import torch.distributed as dist

    from motif_b200 import sharding, synthetic

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        shapes = [(2, 64, 4, 6), (2, 64, 4, 6), (1, 64, 4, 6)]
        ex = sharding.LatentExchange(shapes, "cpu", src=0)
        clips = [synthetic.synthetic_latents(1, 4, 6, seed=s) for s in range(3)]
        ex.start(clips[0] if rank == 0 else None)
        for k in range(3):
            got = ex.take()
            assert all(torch.equal(a, b) for a, b in zip(got, clips[k]))
            ex.release()
            if k + 1 < 3:
                ex.start(clips[k + 1] if rank == 0 else None)
        with pytest.raises(RuntimeError):
            ex.take()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_latent_exchange_pipeline():
    mp.spawn(_exchange_worker, args=(2, _free_port()), nprocs=2, join=True)


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_sliced_copy_in_and_all_gather(world):
    mp.spawn(_slice_worker, args=(world, _free_port()), nprocs=world, join=True)


def test_two_rank_gloo_halo_check():
    mp.spawn(_halo_worker, args=(2, _free_port()), nprocs=2, join=True)


@pytest.mark.parametrize("n_ts", [7, 1])
def test_two_rank_gloo_broadcast_and_gather(n_ts):
    mp.spawn(_worker, args=(2, _free_port(), n_ts), nprocs=2, join=True)


def test_luna_tokis_surround_matches_reference_forward():
    """CPU, build container only: the re-stated surround glue (Ours.py:512-638) reproduces, bit for bit, the
    hot-path inputs of the unmodified reference forward (captured with hooks)."""
    from oracle import ref_shims

    if not ref_shims.reference_available():
        pytest.skip("reference checkout absent (GPU box)")
    from motif_b200 import luna_tokis

    model = ref_shims.build_reference_model(seed=0)
    torch.manual_seed(3)
    x = torch.rand(1, 2, 3, 32, 48)
    target_t = [torch.tensor([[0.25]]), torch.tensor([[0.75]])]
    ref = ref_shims.run_reference_forward(model, x, target_t, 4)
    with torch.no_grad(), ref_shims.cpu_cuda_aliases():
        from oracle import flow_front_ref

        # the product's front operator is the fused CUDA kernel; on the CPU the glue is checked with the oracle's restatement
        feat, flow_feat, residual, tt, hr = luna_tokis.surround(model, x, target_t, 4, iter=4, front=flow_front_ref.flow_front)
    assert hr == (128, 192) and tuple(tt.shape) == (1, 2)
    assert torch.equal(feat, ref["feat"])
    assert torch.equal(residual, ref["residual"])
    # The surround estimates only the pairs 01 / 10 (raft_schedule): the reference's own RAFT forward gives flows that differ by
    # ~1e-5 px between a batch of four and a batch of two pairs (batch-dependent convolution kernels, see the test below), which
    # reaches flow_feat as ~1e-7.
    assert (flow_feat - ref["flow_feat"]).abs().max().item() < 2e-6


def test_raft_schedule_equals_the_reference_forward_on_the_two_live_pairs():
    """CPU, build container only: ``raft_schedule.flow_two_pairs`` (feature encoder once per distinct frame, only the last flow
    upsampled) is BIT-EQUAL to the unmodified ``RAFT.forward`` called on the pairs 01 and 10, and within the reference's own
    batch-size noise of the four-pair call of ``Ours.py:544-545``; the pairs the reference zeroes come back as zeros."""
    from oracle import ref_shims

    if not ref_shims.reference_available():
        pytest.skip("reference checkout absent (GPU box)")
    from motif_b200 import raft_schedule

    model = ref_shims.build_reference_model(seed=0)
    raft = model.flow_predictor
    assert raft_schedule.is_raft(raft)
    torch.manual_seed(4)
    low = torch.rand(2, 3, 16, 24)  # smooth frames; 128x192 keeps the coarsest correlation level larger than one pixel
    fr0, fr1 = torch.nn.functional.interpolate(low, size=(128, 192), mode="bilinear", align_corners=False).split(1)
    with torch.no_grad():
        four = raft(torch.cat([fr0, fr0, fr1, fr1]) * 255.0, torch.cat([fr0, fr1, fr0, fr1]) * 255.0, iters=3)[-1]
        two = raft(torch.cat([fr0, fr1]) * 255.0, torch.cat([fr1, fr0]) * 255.0, iters=3)[-1]
        sched = raft_schedule.four_pair_flows(raft, fr0, fr1, 3)
    assert sched.shape == four.shape and torch.isfinite(four).all()
    assert torch.equal(sched[1:3], two)
    assert (sched[1:3] - four[1:3]).abs().max().item() < 1e-4 * (1.0 + four.abs().max().item())
    assert not sched[0].any() and not sched[3].any()
    assert four[0].abs().max().item() > 0  # the reference does estimate (and then discards) the pair 00


def test_install_keeps_state_dict_layout_and_refuses_training():
    from oracle import ref_shims

    if not ref_shims.reference_available():
        pytest.skip("reference checkout absent (GPU box)")
    from motif_b200 import luna_tokis

    model = ref_shims.build_reference_model(seed=0)
    keys = list(model.state_dict().keys())
    luna_tokis.install(model)
    assert list(model.state_dict().keys()) == keys  # best.pth still loads with strict=True
    model.train()
    with pytest.raises(NotImplementedError):
        model(torch.rand(1, 2, 3, 32, 48), None, [torch.tensor([[0.5]])], 4, use_GT=False)


def test_single_pass_test_equals_the_reference_chunk_loop():
    """VideoSR_base_model.py:188-195: chunks of three timestamps concatenated along dim 0, `flow` of the last chunk --
    reproduced by ONE forward over all timestamps (motif_b200.luna_tokis.test_b200)."""
    from motif_b200 import luna_tokis

    B, HH, WW = 2, 4, 6
    calls = []

    class FakeNet(torch.nn.Module):
        _motif_precision = "f16x3"

        def forward(self, x, real_h, times, scale, use_GT=True, iter=12):
            calls.append(len(times))
            n = len(times)
            tt = torch.stack(times, 1).squeeze(-1)  # [B, n]
            rgb = tt.t().reshape(n, B, 1, 1, 1).expand(n, B, 3, HH, WW).clone()
            # flow index (r*B + b) * n + k, value encodes (r*B + b, t)
            rb = torch.arange(2 * B).view(2 * B, 1).float()
            flow = (rb * 10 + tt.repeat(2, 1)).reshape(2 * B * n, 1, 1, 1).expand(2 * B * n, 2, HH, WW).clone()
            return rgb, flow, 0.0

    class Wrapper:
        net_base = "Ours"

        def __init__(self, times):
            self.netG, self.var_L, self.real_H, self.scale = FakeNet(), torch.zeros(B, 2, 3, 2, 3), None, 2
            self.times = times

        def test(self, output=False):  # the reference loop, VideoSR_base_model.py:169-197 ("Ours" branch)
            self.netG.eval()
            with torch.no_grad():
                self.fake_H, flow, flow_GT = self.netG(self.var_L, self.real_H, self.times[:3], self.scale, use_GT=False, iter=4)
                if len(self.times) != 3:
                    for l in range(3, len(self.times), 3):
                        tmp, flow, flow_GT = self.netG(self.var_L, None, self.times[l:l + 3], self.scale, use_GT=False, iter=4)
                        self.fake_H = torch.cat((self.fake_H, tmp), 0)
                self.flow, self.flow_GT = flow, flow_GT
            self.netG.train()
            if output:
                return self.fake_H

    for n_ts in (1, 3, 7, 11):
        times = [torch.full((B, 1), (k + 1) / (n_ts + 1)) + torch.arange(B).view(B, 1) * 0.001 for k in range(n_ts)]
        ref = Wrapper(times)
        ref_out = ref.test(output=True)
        calls.clear()
        new = Wrapper(times)
        new._motif_reference_test = None
        type(new)._motif_reference_test = Wrapper.test
        new.test = __import__("types").MethodType(luna_tokis.test_b200, new)
        out = new.test(output=True)
        assert calls == [n_ts]  # one forward for the whole clip
        assert torch.equal(out, ref_out) and torch.equal(new.flow, ref.flow) and new.flow_GT == ref.flow_GT
