"""One meeting across several ranks (notsofar_b200.sharded): partition plan, the three exchanges under
torch.distributed (gloo, world_size 2, CPU tensors), and -- on a GPU -- the sharded result against the
single-device path."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import css_oracle as O

from conftest import rel_l2


@pytest.fixture(scope="module")
def N():
    import notsofar_b200
    return notsofar_b200


# ----------------------------------------------------------------------------------------------- partition plan
@pytest.mark.parametrize("hop_sec", [1.5, 1.0, 0.7])
@pytest.mark.parametrize("seconds,world", [(30.0, 1), (30.0, 2), (30.0, 3), (61.3, 4), (1800.0, 8), (4.0, 8), (2.0, 2)])
def test_shards_cover_segments_frames_and_samples(N, seconds, world, hop_sec):
    from notsofar_b200.sharded import make_shard, shard_bounds
    cfg = N.CssCfg(hop_size_sec=hop_sec)
    n = int(seconds * 16000)
    plan = N.plan_segments(n, 16000, cfg)
    shards = [make_shard(plan, r, world) for r in range(world)]
    assert shard_bounds(plan.num_segments, world)[-1] == plan.num_segments
    seg_next, frame_next = 0, 0
    for sh in shards:
        assert sh.seg_lo == seg_next
        seg_next = sh.seg_hi
        if sh.n_own_seg == 0:
            assert sh.n_own_frames == 0 and sh.n_loc_seg == 0
            continue
        assert sh.own_lo == frame_next
        frame_next = sh.own_hi
        # every segment that overlaps an owned output frame is local (ceil(T / hop) - 1 segments to the left of seg_lo)
        T, hop = plan.segment_frames, plan.hop_frames
        touching = [s for s in range(plan.num_segments) if s * hop < sh.own_hi and s * hop + T > sh.own_lo]
        assert min(touching) >= sh.seg_lo - sh.halo and max(touching) <= sh.seg_hi - 1
        assert sh.halo == min(sh.seg_lo, max(1, -(-T // hop) - 1))
        assert sh.frame0 == (sh.seg_lo - sh.halo) * plan.hop_frames
        assert sh.frame0 + sh.n_frames >= min(plan.mix_frames, (sh.seg_hi - 1) * plan.hop_frames + plan.segment_frames)
        assert sh.own_lo >= sh.frame0 and sh.own_hi <= sh.frame0 + sh.n_frames
        # the sample range holds exactly the valid local frames
        assert sh.sample_lo == sh.frame0 * 256
        if sh.valid_frames:
            assert sh.sample_hi <= n and (sh.sample_hi - sh.sample_lo - 512) // 256 + 1 == sh.valid_frames
    assert seg_next == plan.num_segments and frame_next == plan.mix_frames
    sizes = [s.n_own_seg for s in shards]
    assert max(sizes) - min(sizes) <= 1


def test_assemble_waveforms_is_the_overlap_add_of_the_pieces(N):
    """Pieces cut from per-frame contributions re-assemble to the global overlap-add (seams included)."""
    from notsofar_b200.sharded import make_shard, assemble_waveforms
    cfg = N.CssCfg()
    plan = N.plan_segments(16000 * 20, 16000, cfg)
    rng = np.random.default_rng(0)
    frames = rng.standard_normal((3, plan.mix_frames, 512)).astype(np.float32)       # windowed frame outputs
    ref = np.zeros((3, (plan.mix_frames - 1) * 256 + 512), np.float32)
    for t in range(plan.mix_frames):
        ref[:, t * 256:t * 256 + 512] += frames[:, t]
    world = 3
    shards = [make_shard(plan, r, world) for r in range(world)]
    pieces = []
    for sh in shards:
        p = np.zeros((3, sh.n_own_frames * 256 + 256), np.float32)
        for t in range(sh.own_lo, sh.own_hi):
            o = (t - sh.own_lo) * 256
            p[:, o:o + 512] += frames[:, t]
        pieces.append(torch.from_numpy(p))
    got = assemble_waveforms(pieces, shards, plan.mix_frames).numpy()
    assert np.array_equal(got, ref)


# ----------------------------------------------------------------------------------------------- gloo, world_size 2
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, n_samples, ret):
    import torch.distributed as dist
    import notsofar_b200 as N
    from notsofar_b200.sharded import make_shard, allgather_varlen, gather_varlen, assemble_waveforms, gather_waveforms
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = N.CssCfg()
        plan = N.plan_segments(n_samples, 16000, cfg)
        shards = [make_shard(plan, r, world) for r in range(world)]
        sh = shards[rank]
        # per-segment masks from a seeded generator every rank can replay (stands in for the mask network)
        rng = np.random.default_rng(5)
        masks = rng.random((plan.num_segments, 3, 257, plan.segment_frames)).astype(np.float32)
        ov = plan.overlap_frames
        costs = np.zeros((plan.num_segments, 3, 3), np.float32)
        for i in range(1, plan.num_segments):
            l, r_ = masks[i - 1][:, :, -ov:], masks[i][:, :, :ov]
            costs[i] = np.abs(l[:, None] - r_[None]).mean(axis=(2, 3))
        # exchange 1: owned costs -> all ranks
        own = torch.from_numpy(costs[sh.seg_lo:sh.seg_hi].copy())
        allc = allgather_varlen(own, [s.n_own_seg for s in shards]).numpy()
        assert np.array_equal(allc, costs)
        perms = N.permutation_chain(allc)
        # exchange 2: owned activity rows
        act = rng.random((plan.mix_frames, 3)).astype(np.float32)
        alla = allgather_varlen(torch.from_numpy(act[sh.own_lo:sh.own_hi].copy()), [s.n_own_frames for s in shards]).numpy()
        assert np.array_equal(alla, act)
        # exchange 3: waveform pieces -> rank 0, seams overlap-added
        frames = rng.standard_normal((3, plan.mix_frames, 512)).astype(np.float32)
        p = np.zeros((3, sh.n_own_frames * 256 + 256), np.float32)
        for t in range(sh.own_lo, sh.own_hi):
            o = (t - sh.own_lo) * 256
            p[:, o:o + 512] += frames[:, t]
        counts = [s.n_own_frames * 256 + 256 if s.n_own_frames else 0 for s in shards]
        pieces = gather_varlen(torch.from_numpy(p), counts, dst=0, dim=1)
        direct = gather_waveforms(torch.from_numpy(p), shards, plan.mix_frames, dst=0)
        if rank == 0:
            got = assemble_waveforms(pieces, shards, plan.mix_frames).numpy()
            ref = np.zeros_like(got)
            for t in range(plan.mix_frames):
                ref[:, t * 256:t * 256 + 512] += frames[:, t]
            assert np.array_equal(got, ref)
            # the point-to-point hand-off (bodies in place + seam adds): a + b against (0 + a) + b -- the same float32
            assert np.array_equal(direct.numpy(), ref)
            ret["perms"] = perms
        else:
            assert pieces is None and direct is None
        ret[f"ok{rank}"] = True
    finally:
        dist.destroy_process_group()


def test_exchanges_under_gloo_world2(N):
    import torch.multiprocessing as mp
    port = _free_port()
    n = 16000 * 25 + 777
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_gloo_worker, args=(2, port, n, ret), nprocs=2, join=True)
        assert ret.get("ok0") and ret.get("ok1")
        plan = N.plan_segments(n, 16000, N.CssCfg())
        assert ret["perms"].shape == (plan.num_segments, 3)


# ----------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_sharded_no_beamformer_power_norm_equals_single_device(N):
    """Single-channel model, no MVDR, normalize_segment_power: the per-segment power ratios use the global frame count."""
    from notsofar_b200.css import css_device
    from notsofar_b200.sharded import css_sharded_on_one_device
    from notsofar_b200 import synth
    dev = torch.device("cuda", 0)
    w = O.random_weights(seed=2, d_model=128, n_heads=2, d_ff=256, n_blocks=2, in_features=257)
    sep = N.ConformerCssB200(w, device=dev)
    x = torch.from_numpy(np.ascontiguousarray(synth.synthetic_meeting(10.3, seed=4)[:, :1])).to(dev)
    cfg = N.CssCfg(activity_th=0.5, show_progressbar=False, normalize_segment_power=True)
    one = css_device(x, sep, 16000, cfg)
    sh = css_sharded_on_one_device(x, sep, 16000, cfg, 3)
    torch.cuda.synchronize()
    assert np.array_equal(sh["perms"], one["perms"]) and torch.equal(sh["mask_stitched"], one["mask_stitched"])
    for w_Y, s in zip(sh["Y"], sh["shards"]):
        assert torch.equal(w_Y.view(torch.float32), one["Y"][s.seg_lo - s.halo:s.seg_hi].view(torch.float32))      # no beamformer: bit-exact
    assert rel_l2(sh["wav"].cpu().numpy(), one["wav"].cpu().numpy()) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("world,seconds,hop_sec", [(2, 14.0, 1.5), (3, 21.7, 1.5), (5, 6.1, 1.5), (3, 21.7, 1.0), (2, 9.0, 0.7)])
def test_sharded_equals_single_device(N, small_weights, world, seconds, hop_sec):
    """All ranks of the sharded algorithm played on one GPU vs css_device: integers and everything up to the stitched
    masks bit-exact; waveforms equal away from the shard seams and within 1e-6 at them (the iSTFT packs frame
    pairs into one complex FFT, so a frame's rounding depends on its partner)."""
    from notsofar_b200.css import css_device
    from notsofar_b200.sharded import css_sharded_on_one_device
    from notsofar_b200 import synth
    dev = torch.device("cuda", 0)
    sep = N.ConformerCssB200(small_weights, device=dev)
    x = torch.from_numpy(synth.synthetic_meeting(seconds, seed=3)).to(dev)
    # random-init masks hover around 0.5: put the threshold at their 99th percentile (sparse islands) so that the gate (and its
    # dilate / erode across the shard seams) is exercised
    probe = css_device(x, sep, 16000, N.CssCfg(show_progressbar=False, hop_size_sec=hop_sec))
    cfg = N.CssCfg(activity_th=float(torch.quantile(probe["activity"].flatten(), 0.99)), show_progressbar=False, hop_size_sec=hop_sec)
    one = css_device(x, sep, 16000, cfg)
    sh = css_sharded_on_one_device(x, sep, 16000, cfg, world)
    torch.cuda.synchronize()
    assert np.array_equal(sh["perms"], one["perms"])
    assert torch.equal(sh["activity_b"], one["activity_b"]) and torch.equal(sh["activity_final"], one["activity_final"])
    assert torch.equal(sh["mask_stitched"], one["mask_stitched"])
    assert torch.equal(sh["activity"], one["activity"])
    for w_masks, w_Y, s in zip(sh["masks"], sh["Y"], sh["shards"]):
        if s.n_loc_seg:
            assert torch.equal(w_masks, one["masks"][s.seg_lo - s.halo:s.seg_hi])
            # MVDR: the streaming kernel sums a block's frames in the order of the (older, newer) winner classes, so the last
            # segment of a rank (no newer neighbour in its batch) rounds its fp64 covariances differently: equal to fp32 round-off
            a_, b_ = w_Y.view(torch.float32), one["Y"][s.seg_lo - s.halo:s.seg_hi].view(torch.float32)
            assert float((a_ - b_).norm() / b_.norm()) < 1e-5
    a, b = sh["wav"].cpu().numpy(), one["wav"].cpu().numpy()
    assert a.shape == b.shape
    err = rel_l2(a, b)
    print(f"sharded (world {world}) vs single device: waveform rel_l2 = {err:.2e}, max abs = {np.abs(a - b).max():.2e}")
    assert err < 1e-6
    assert one["activity_final"].any() and not one["activity_final"].all(), "test input must exercise the gate"


# ----------------------------------------------------------------------------------------------- sessions over ranks
@pytest.mark.gpu
@pytest.mark.parametrize("world,seconds,hop_sec,per_batch", [(3, 30.0, 1.5, 2), (2, 20.0, 1.0, 3), (4, 24.0, 1.5, 1)])
def test_sharded_progressive_host_pieces(N, small_weights, world, seconds, hop_sec, per_batch):
    """Every rank reads the interior of its waveform piece back while its segments are still in the network, under
    the labels of a local permutation chain, and relabels at the end (ShardWorker.phase1(host_piece) / finish_host):
    the host rows must be bit for bit the piece the global tail produces.  All ranks are played by one device."""
    from notsofar_b200.sharded import ShardWorker
    from notsofar_b200.css import css_device
    from notsofar_b200 import synth
    dev = torch.device("cuda:0")
    x = synth.synthetic_meeting(seconds, seed=4)
    n = x.shape[0]
    probe = css_device(torch.from_numpy(x).to(dev), N.ConformerCssB200(small_weights, device=dev), 16000,
                       N.CssCfg(show_progressbar=False, hop_size_sec=hop_sec))
    cfg = N.CssCfg(activity_th=float(torch.quantile(probe["activity"].flatten(), 0.9)), show_progressbar=False, hop_size_sec=hop_sec)
    import itertools
    orders = list(itertools.permutations(range(3)))

    class Shuffled(N.ConformerCssB200):
        """The network's speaker channels in a different order for every (global) segment: the chain has to undo it, so
        the ranks behind the first one start in a permuted order and the relabelling is exercised."""
        seg_offset = 0

        def masks(self, X, T_valid, s0, nb, T, hop, out=None):
            out = super().masks(X, T_valid, s0, nb, T, hop, out=out)
            for i in range(nb):
                order = list(orders[((self.seg_offset + s0 + i) * 5 + 2) % 6])
                out[i, :3] = out[i, order].clone()
            return out

    sep = Shuffled(small_weights, device=dev, segments_per_batch=per_batch)
    xd = torch.from_numpy(x).to(dev)
    wks = [ShardWorker(sep, 16000, cfg, n, r, world) for r in range(world)]
    hosts = [torch.empty((3, w.sh.n_own_frames * 256 + 256), dtype=torch.float32).pin_memory() for w in wks]
    for h in hosts:
        h.fill_(float("nan"))
    costs = []
    for w, h in zip(wks, hosts):
        sep.seg_offset = w.sh.seg_lo - w.sh.halo
        costs.append(w.phase1(xd[w.sh.sample_lo:w.sh.sample_hi].contiguous(), h))
    assert all(w.prog is not None for w in wks if w.sh.n_loc_seg > per_batch)
    costs_all = torch.cat(costs, 0).cpu().numpy()
    # what the one-shot path computes for the same block (costs of all local segments in one launch)
    ref = [ShardWorker(sep, 16000, cfg, n, r, world) for r in range(world)]
    costs_ref = []
    for w in ref:
        sep.seg_offset = w.sh.seg_lo - w.sh.halo
        costs_ref.append(w.phase1(xd[w.sh.sample_lo:w.sh.sample_hi].contiguous()))
    assert np.array_equal(costs_all, torch.cat(costs_ref, 0).cpu().numpy())
    activity_all = torch.cat([w.phase2(costs_all) for w in wks], 0)
    early, relabelled = 0, 0
    for w, h in zip(wks, hosts):
        out = w.phase3(activity_all)
        rows = w.finish_host(out["wav_piece"], h)
        torch.cuda.synchronize()
        assert w.progressive_ok
        piece = out["wav_piece"].cpu()
        for k in range(3):
            assert torch.equal(rows[k], piece[k]), (w.sh.rank, k)
        early += sum(hi - lo for lo, hi in w.prog["copied"])
        print(f"rank {w.sh.rank}: relabel {w.relabel.tolist()}, early samples {sum(hi - lo for lo, hi in w.prog['copied'])} of {piece.shape[1]}")
        relabelled += int(w.relabel.tolist() != [0, 1, 2])
    assert early > 0                                   # something did leave early
    assert relabelled > 0                              # ... and some rank's local labels were not the global ones


def test_assign_sessions_balances_and_covers():
    from notsofar_b200.scheduler import assign_sessions
    rng = np.random.default_rng(0)
    d = rng.uniform(300, 900, size=37).tolist()
    for world in (1, 2, 8):
        shares = assign_sessions(d, world)
        flat = sorted(i for s in shares for i in s)
        assert flat == list(range(37))
        loads = [sum(d[i] for i in s) for s in shares]
        assert max(loads) - min(loads) <= max(d)                      # LPT bound
    assert assign_sessions([], 4) == [[], [], [], []]
    assert assign_sessions([5.0, 5.0], 4) == [[0], [1], [], []]


def _sched_worker(rank, world, port, ret):
    import pandas as pd
    import torch.distributed as dist
    from notsofar_b200.scheduler import css_inference_distributed
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sessions = [pd.Series(dict(session_id=f"multichannel/MTG_{i}", wav_file_names=[f"/x/{i}.wav"])) for i in range(5)]
        calls = []

        def fake_css(out_dir, models_dir, session, cfg, fetch):
            calls.append(session.session_id)
            s = session.copy()
            s["sep_wav_file_names"] = [f"{out_dir}/{session.session_id}/sep_stream{k}.wav" for k in range(3)]
            return s
        out = css_inference_distributed("/out", "/models", sessions, None, False, css_fn=fake_css, durations=[10, 50, 20, 40, 30])
        assert [o.session_id for o in out] == [s.session_id for s in sessions]
        assert all(len(o.sep_wav_file_names) == 3 for o in out)
        ret[f"calls{rank}"] = calls
    finally:
        dist.destroy_process_group()


def test_sessions_distributed_under_gloo_world2():
    import torch.multiprocessing as mp
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_sched_worker, args=(2, port, ret), nprocs=2, join=True)
        c0, c1 = ret["calls0"], ret["calls1"]
        assert sorted(c0 + c1) == [f"multichannel/MTG_{i}" for i in range(5)] and c0 and c1


# ----------------------------------------------------------------------------------------------- diarization: words over ranks
def _words_df():
    import pandas as pd
    words = [[f"w{i}", 0.5 + 0.7 * i, 0.5 + 0.7 * i + 0.15 + 0.05 * (i % 7)] for i in range(23)]
    df = pd.DataFrame({"start_time": [0.5, 8.0], "end_time": [8.0, 17.0], "text": ["a", "b"], "word_timing": [words[:11], words[11:]],
                       "meeting_id": ["m", "m"], "session_id": ["s", "s"], "wav_file_name": ["s0.wav", "s1.wav"],
                       "wav_file_name_ind": [0, 1]})
    df["wav_file_name"] = df["wav_file_name"].astype("category")
    return df


class _FakePcm:
    """Host stand-in for the device-resident PCM16 streams (the gloo test has no GPU): shape / device like a tensor."""
    def __init__(self, a):
        self.a, self.shape, self.device = a, a.shape, torch.device("cpu")


def _host_gather(pcm, plan, first=0, count=None):
    count = len(plan.start) - first if count is None else count
    L = int(plan.length[first:first + count].max()) if count else 0
    out = torch.zeros((count, L))
    for i in range(count):
        r = first + i
        out[i, :plan.length[r]] = torch.from_numpy(pcm.a[plan.stream_id[r], plan.start[r]:plan.start[r] + plan.length[r]].astype(np.float32) / 32767.0)
    return out, torch.from_numpy(plan.length[first:first + count].copy())


def _fake_embed(crops, lens, cfg):
    n = torch.arange(crops.shape[1])[None, :] < lens[:, None]
    s = (crops * n).sum(1)
    return torch.stack([s, (crops.abs() * n).sum(1), lens.float() / 16000.0, torch.ones(len(lens))], 1)


def _words_worker(rank, world, port, ret):
    import torch.distributed as dist
    import notsofar_b200.diarization as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(3)
        pcm = _FakePcm((rng.standard_normal((2, 16000 * 18)) * 4000).astype(np.int16))
        seen = {}
        D.gather_word_crops = _host_gather                       # test stand-in for the CUDA gather
        D.set_embedding_backend(_fake_embed)
        D.set_clustering_backend(lambda emb, cfg: (seen.setdefault("emb", emb.clone()), (emb[:, 0, 2] > emb[:, 0, 2].median()).int().numpy())[1])
        cfg = D.DiarizationCfg(method="word_nmesc", min_embedding_windows=[1.0, 0.5])
        out = D.word_based_clustering(pcm, 16000, _words_df(), cfg, batch_words=4, shard_words=True)
        ret[f"emb{rank}"] = seen["emb"].numpy()
        ret[f"spk{rank}"] = out.speaker_id.tolist()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_word_sharded_embeddings_under_gloo(world):
    """SURVEY 8e, diarization: every rank embeds a block of the words, one all-gather, identical clustering input everywhere
    and equal to the unsharded run."""
    import torch.multiprocessing as mp
    import notsofar_b200.diarization as D
    assert D.word_shard_bounds(23, 3) == [0, 7, 15, 23] and D.word_shard_bounds(2, 4) == [0, 0, 1, 1, 2]
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_words_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
        rng = np.random.default_rng(3)
        pcm = _FakePcm((rng.standard_normal((2, 16000 * 18)) * 4000).astype(np.int16))
        seen = {}
        real_gather = D.gather_word_crops
        D.gather_word_crops = _host_gather
        D.set_embedding_backend(_fake_embed)
        D.set_clustering_backend(lambda emb, cfg: (seen.setdefault("emb", emb.clone()), (emb[:, 0, 2] > emb[:, 0, 2].median()).int().numpy())[1])
        try:
            cfg = D.DiarizationCfg(method="word_nmesc", min_embedding_windows=[1.0, 0.5])
            one = D.word_based_clustering(pcm, 16000, _words_df(), cfg, batch_words=4)
        finally:
            D.gather_word_crops = real_gather
            D.set_embedding_backend(None)
            D.set_clustering_backend(None)
        for r in range(world):
            assert np.array_equal(ret[f"emb{r}"], seen["emb"].numpy())
            assert ret[f"spk{r}"] == one.speaker_id.tolist()
