"""TitaNet speaker-embedding forward + multi-scale cosine affinity.

NeMo is absent offline and the reference holds no vectors for this path: **parity unpinned**.  The CUDA path is compared,
through the C ABI, with the numpy restatement of the published architecture (oracle/titanet_oracle.py); the CPU tests pin the
restatement's own invariants (parameter count of titanet-large, front-end closed forms, masking == cropping)."""
import numpy as np
import pytest

from oracle import titanet_oracle as O
from conftest import rel_l2

SMALL = ((64, 1, 3, False), (64, 2, 7, True), (128, 2, 5, True), (192, 1, 1, False))


def _crops(rng, lens):
    out = []
    for n in lens:
        t = np.arange(n) / 16000.0
        f0 = rng.uniform(90, 250)
        x = sum(np.sin(2 * np.pi * f0 * h * t + rng.uniform(0, 6.28)) / h for h in range(1, 12))
        x = x * (0.5 + 0.5 * np.sin(2 * np.pi * 3.1 * t)) * 0.05 + 0.01 * rng.standard_normal(n)
        out.append(x.astype(np.float32))
    return out


# ------------------------------------------------------------------------------------------------ CPU: the restatement itself
def test_titanet_large_parameter_count():
    """encoder + pooling + embedding layer of titanet-large: 25.3 M published minus the 192 x 16681 classifier."""
    w = O.random_weights(0)
    n = sum(v.size for k, v in w.items() if "running" not in k)
    assert 21.9e6 < n < 22.3e6, n
    assert O.block_plan()[-1][:2] == (1024, 3072)


def test_frontend_closed_forms():
    rng = np.random.default_rng(0)
    x = _crops(rng, [16000])[0]
    f = O.features(x)
    assert f.shape == (O.seq_len(16000), 80) == (101, 80)
    np.testing.assert_allclose(f.mean(0), 0, atol=1e-9)
    np.testing.assert_allclose(f.std(0, ddof=1), 1, atol=1e-4)          # unbiased std + 1e-5
    fb = O.mel_filterbank()
    assert fb.shape == (80, 257) and (fb >= 0).all() and (fb.sum(1) > 0).all()
    # slaney normalisation: every triangle has unit area in Hz (up to the discretisation of the 31.25-Hz grid)
    np.testing.assert_allclose(fb.sum(1) * 31.25, 1.0, atol=0.2)


def test_pack_matches_plan():
    import notsofar_b200.titanet as T
    w = O.random_weights(3, blocks=SMALL, att_ch=32, emb=16)
    assert T.infer_blocks(w) == SMALL
    dims, blob, offsets = T.pack_titanet(w, SMALL)
    lib = T._cabi.load()
    import ctypes as C
    assert lib.nsf_titanet_num_offsets(C.byref(dims)) == len(offsets)
    assert blob.dtype == np.float32 and (offsets % 64 == 0).all()
    assert lib.nsf_titanet_workspace_bytes(C.byref(dims), 4, 32) > 0


def test_checkpoint_loader_nemo_archive_and_plain_file(tmp_path):
    """load_titanet_state_dict: a .nemo archive (tar with model_weights.ckpt) and a plain torch state-dict file; classifier /
    preprocessor entries are dropped, the block structure is recovered from the names."""
    import tarfile
    import torch
    import notsofar_b200.titanet as T
    w = O.random_weights(3, blocks=SMALL, att_ch=32, emb=16)
    sd = {k: torch.from_numpy(v) for k, v in w.items()}
    sd["decoder.final.weight"] = torch.zeros(10, 16)                       # classifier: kept under decoder.*, unused by the packer
    sd["preprocessor.featurizer.window"] = torch.zeros(400)
    ck = tmp_path / "model_weights.ckpt"
    torch.save(sd, ck)
    arch = tmp_path / "m.nemo"
    with tarfile.open(arch, "w:gz") as tar:
        tar.add(ck, arcname="./model_weights.ckpt")
    for path in (arch, ck):
        got = T.load_titanet_state_dict(str(path))
        assert "preprocessor.featurizer.window" not in got and set(w) <= set(got)
        assert T.infer_blocks(got) == SMALL
        dims, blob, offs = T.pack_titanet(got, T.infer_blocks(got))
        assert (dims.feat_in, dims.n_blocks, dims.att_ch, dims.emb) == (80, len(SMALL), 32, 16)
    bad = tmp_path / "empty.nemo"
    with tarfile.open(bad, "w") as tar:
        tar.add(__file__, arcname="readme.txt")
    with pytest.raises(T._cabi.NsfError):
        T.load_titanet_state_dict(str(bad))


def test_cos_affinity_oracle():
    rng = np.random.default_rng(1)
    e = rng.standard_normal((7, 16))
    a = O.cos_affinity(e)
    assert a.shape == (7, 7) and np.allclose(a, a.T) and a.max() == 1.0 and a.min() == 0.0
    assert O.cos_affinity(e[:1]).tolist() == [[1.0]]


def _blob_affinity(rng, sizes, d=24, noise=0.25):
    centers = 3.0 * np.eye(len(sizes), d)                                  # orthogonal 'voices'
    emb = np.concatenate([c + noise * rng.standard_normal((n, d)) for c, n in zip(centers, sizes)])
    truth = np.concatenate([np.full(n, i) for i, n in enumerate(sizes)])
    perm = rng.permutation(len(truth))
    return O.cos_affinity(emb[perm]), truth[perm]


@pytest.mark.parametrize("sizes", [(160, 100, 140), (150, 150), (200, 80, 80, 120), (300,)])
def test_nmesc_spectral_clustering_recovers_blobs(sizes):
    """clustering.py on the CPU (it runs where the affinity lives): speaker count and partition of well-separated speakers (a few
    hundred words, like a session: with a handful of words the candidate list of NMESC ends at p = 2..3 neighbours and upstream's
    getMinimumConnection walk then returns the estimate of a shattered graph)."""
    import torch
    import notsofar_b200.clustering as K
    rng = np.random.default_rng(len(sizes))
    aff, truth = _blob_affinity(rng, sizes)
    a = torch.from_numpy(aff).float()
    k, p_hat = K.nmesc(a)
    assert k == len(sizes) and p_hat >= 1
    labels = K.run_clustering(a)
    assert len(np.unique(labels)) == len(sizes)
    for c in np.unique(labels):                                   # every cluster is pure: labels equal truth up to a permutation
        assert len(np.unique(truth[labels == c])) == 1
    g = K.affinity_graph(a, p_hat)
    assert torch.equal(g, g.T)          # (well-separated blobs may leave the p-neighbour graph disconnected: the search list ends at max_N)
    L = K.laplacian(g)
    assert torch.allclose(L.sum(1), torch.zeros(len(truth)), atol=1e-5)


@pytest.mark.parametrize("n", [2, 3, 5, 40, 1100])
def test_clustering_degenerate_sizes(n):
    """Tiny sessions (a handful of words) and more words than NeMo's 512-point search matrix: labels for every word, no crash."""
    import torch
    import notsofar_b200.clustering as K
    rng = np.random.default_rng(n)
    a = torch.from_numpy(O.cos_affinity(rng.standard_normal((n, 16)))).float()
    labels = K.run_clustering(a)
    assert labels.shape == (n,) and labels.min() >= 0 and labels.max() < 8


def test_kneighbors_graph_small_known_answer():
    import torch
    import notsofar_b200.clustering as K
    a = torch.tensor([[1.0, 0.9, 0.1], [0.9, 1.0, 0.2], [0.1, 0.2, 1.0]])
    x = K.kneighbors_connections(a, 2)                            # row i's two best are {i, its nearest other}; stored transposed
    assert x.tolist() == [[1.0, 1.0, 0.0], [1.0, 1.0, 1.0], [0.0, 0.0, 1.0]]
    assert K.affinity_graph(a, 2).tolist() == [[1.0, 1.0, 0.0], [1.0, 1.0, 0.5], [0.0, 0.5, 1.0]]
    assert K.run_clustering(torch.ones(1, 1)).tolist() == [0]


# ------------------------------------------------------------------------------------------------ GPU parity
@pytest.fixture(scope="module")
def dev():
    import torch
    return torch.device("cuda", 0)


def _run(model, crops, dev):
    import torch
    n, L = len(crops), max(len(c) for c in crops)
    buf = np.zeros((n, L), np.float32)
    for i, c in enumerate(crops):
        buf[i, :len(c)] = c
    lens = torch.tensor([len(c) for c in crops], dtype=torch.int32, device=dev)
    return model, torch.from_numpy(buf).to(dev), lens


@pytest.mark.gpu
def test_titanet_features_vs_oracle(dev):
    import torch
    import notsofar_b200.titanet as T
    w = O.random_weights(3, blocks=SMALL, att_ch=32, emb=16)
    model = T.TitaNetB200(w, dev)
    rng = np.random.default_rng(2)
    crops = _crops(rng, [8000, 48000, 4321, 24000, 16001])
    _, x, lens = _run(model, crops, dev)
    hi, lo, nf, t_pad = model.features(x, lens)
    assert t_pad % 16 == 0 and nf.cpu().tolist() == [O.seq_len(len(c)) for c in crops]          # integer frame counts: exact
    feat = (hi.float() + lo.float()).cpu().numpy()
    for i, c in enumerate(crops):
        ref = O.features(c)
        T_i = ref.shape[0]
        assert rel_l2(feat[i, :T_i], ref) < 2e-4, (i, rel_l2(feat[i, :T_i], ref))
        assert not feat[i, T_i:].any()                                                           # masked beyond the length


@pytest.mark.gpu
@pytest.mark.parametrize("blocks,att,emb,lens,precision,tol", [
    (SMALL, 32, 16, [8000, 48000, 4321, 24000, 16001, 12000], "fp32", 2e-4),
    (O.BLOCKS, 128, 192, [8000, 24000, 40000], "fp32", 5e-4),             # titanet-large dims
    # the default engine: fp16 operands, fp32 accumulation (the reference's autocast arithmetic): 2^-11 per operand
    (SMALL, 32, 16, [8000, 48000, 4321, 24000, 16001, 12000], "fp16", 5e-3),
    (O.BLOCKS, 128, 192, [8000, 24000, 40000], "fp16", 5e-3),
])
def test_titanet_embedding_vs_oracle(dev, blocks, att, emb, lens, precision, tol):
    import notsofar_b200.titanet as T
    w = O.random_weights(5, blocks=blocks, att_ch=att, emb=emb)
    model = T.TitaNetB200(w, dev, precision=precision)
    assert model.precision == precision and model.dims.precision == (1 if precision == "fp16" else 0)
    rng = np.random.default_rng(4)
    crops = _crops(rng, lens)
    _, x, l = _run(model, crops, dev)
    e = model.embed(x, l).cpu().numpy()
    ref = O.embed(w, crops, blocks)
    assert e.shape == ref.shape == (len(lens), emb)
    errs = [rel_l2(e[i], ref[i]) for i in range(len(lens))]
    print(f"titanet embedding rel err vs fp64 oracle [{precision}]", errs)
    assert max(errs) < tol, errs
    if precision == "fp16":
        cos = [float(np.dot(e[i], ref[i]) / (np.linalg.norm(e[i]) * np.linalg.norm(ref[i]))) for i in range(len(lens))]
        assert min(cos) > 0.9999, cos                                      # what the affinity matrix sees
    # batching must not change a crop's embedding (padding is masked everywhere): the same crop alone
    _, x1, l1 = _run(model, crops[:1], dev)
    e1 = model.embed(x1, l1).cpu().numpy()
    assert rel_l2(e1[0], e[0]) < 1e-5


@pytest.mark.gpu
def test_multiscale_affinity_vs_oracle(dev):
    import torch
    import notsofar_b200.titanet as T
    rng = np.random.default_rng(6)
    emb = rng.standard_normal((37, 6, 192)).astype(np.float32)
    a = T.multiscale_affinity(torch.from_numpy(emb).to(dev)).cpu().numpy()
    ref = np.mean([O.cos_affinity(emb[:, s]) for s in range(6)], axis=0)
    assert np.abs(a - ref).max() < 2e-6
    assert T.multiscale_affinity(torch.from_numpy(emb[:1]).to(dev)).cpu().tolist() == [[1.0]]


@pytest.mark.gpu
def test_word_based_clustering_with_titanet_backend(dev):
    """The a16 pipeline end to end on device-resident streams: crop plan -> gather -> TitaNet embeddings -> affinity ->
    a clustering backend (NMESC itself is NeMo: plug-in point)."""
    import pandas as pd
    import torch
    import notsofar_b200.diarization as D
    import notsofar_b200.titanet as T
    w = O.random_weights(3, blocks=SMALL, att_ch=32, emb=16)
    model = T.TitaNetB200(w, dev)
    rng = np.random.default_rng(8)
    sr, n = 16000, 16000 * 20
    pcm = torch.from_numpy((rng.standard_normal((3, n)) * 3000).astype(np.int16)).to(dev)
    words = [[f"w{i}", 1.0 + 0.9 * i, 1.0 + 0.9 * i + 0.2 + 0.1 * (i % 5)] for i in range(18)]
    df = pd.DataFrame({"start_time": [1.0, 9.0], "end_time": [9.0, 19.0], "text": ["a", "b"], "word_timing": [words[:9], words[9:]],
                       "meeting_id": ["m", "m"], "session_id": ["s", "s"], "wav_file_name": ["s0.wav", "s1.wav"],
                       "wav_file_name_ind": [0, 1]})
    df["wav_file_name"] = df["wav_file_name"].astype("category")          # as diarization_inference prepares it (diarization.py:94-97)
    cfg = D.DiarizationCfg(method="word_nmesc", min_embedding_windows=[1.5, 1.0, 0.5])
    seen = {}

    def cluster(emb, cfg_):
        seen["emb"] = emb
        aff = T.multiscale_affinity(emb.float())
        seen["aff"] = aff
        return (aff[0] < aff[0].median()).int().cpu().numpy()

    D.set_embedding_backend(model.as_embedding_backend())
    D.set_clustering_backend(cluster)
    try:
        out = D.word_based_clustering(pcm, sr, df, cfg)
    finally:
        D.set_embedding_backend(None)
        D.set_clustering_backend(None)
    assert seen["emb"].shape == (18, 3, 16) and torch.isfinite(seen["emb"]).all()
    assert seen["aff"].shape == (18, 18)
    assert set(out["speaker_id"]) <= {"spk0", "spk1"} and len(out) >= 2
    # embeddings equal the oracle's on the same crops (first word, all scales)
    plan = D.word_crop_plan(df, n, sr, cfg.min_embedding_windows, cfg.max_allowed_word_duration)
    x = pcm.cpu().numpy().astype(np.float32) / 32767.0
    ref = O.embed(w, [x[plan.stream_id[i], plan.start[i]:plan.start[i] + plan.length[i]] for i in range(3)], SMALL)
    got = seen["emb"][0].cpu().numpy()
    assert rel_l2(got, ref) < 2e-4


@pytest.mark.gpu
def test_word_nmesc_default_backends_from_checkpoint(dev, tmp_path, monkeypatch):
    """diarization_inference-style use with nothing registered: weights from a .nemo-like archive (NSF_TITANET_CKPT), TitaNet
    kernels, CUDA affinity, NMESC.  Two synthetic 'speakers' (different harmonic stacks) on two streams are told apart."""
    import io
    import tarfile
    import pandas as pd
    import torch
    import notsofar_b200.diarization as D
    w = O.random_weights(3, blocks=SMALL, att_ch=32, emb=16)
    ck = tmp_path / "model_weights.ckpt"
    torch.save({k: torch.from_numpy(v) for k, v in w.items()}, ck)
    arch = tmp_path / "titanet_small.nemo"
    with tarfile.open(arch, "w") as tar:
        tar.add(ck, arcname="./model_weights.ckpt")
    monkeypatch.setenv("NSF_TITANET_CKPT", str(arch))
    sr, n = 16000, 16000 * 60
    t = np.arange(n) / sr
    rng = np.random.default_rng(12)
    voices = []
    for f0, tilt in ((110.0, 1.0), (235.0, 2.2)):
        v = sum(np.sin(2 * np.pi * f0 * h * t + rng.uniform(0, 6.28)) / h ** tilt for h in range(1, 20))
        voices.append(v * (0.6 + 0.4 * np.sin(2 * np.pi * 2.7 * t)))
    pcm_np = np.stack([voices[0], voices[1], 0.01 * rng.standard_normal(n)])
    pcm_np = (pcm_np / np.abs(pcm_np).max(1, keepdims=True) * 20000).astype(np.int16)
    pcm = torch.from_numpy(pcm_np).to(dev)
    # a session-like number of words: with a handful, NMESC's candidate list ends at p = 2 neighbours and the speaker count read off
    # that shattered graph is meaningless (upstream's getMinimumConnection walk behaves the same)
    words = [[f"w{i}", 1.0 + 0.47 * (i % 60), 1.0 + 0.47 * (i % 60) + 0.3] for i in range(120)]
    df = pd.DataFrame({"start_time": [1.0, 1.0], "end_time": [30.0, 30.0], "text": ["a", "b"], "word_timing": [words[:60], words[60:]],
                       "meeting_id": ["m", "m"], "session_id": ["s", "s"], "wav_file_name": ["s0.wav", "s1.wav"],
                       "wav_file_name_ind": [0, 1]})
    df["wav_file_name"] = df["wav_file_name"].astype("category")
    cfg = D.DiarizationCfg(method="word_nmesc", min_embedding_windows=[1.5, 1.0, 0.5], apply_deduplication=False)
    out = D.word_based_clustering(pcm, sr, df, cfg)
    by_stream = out.groupby("wav_file_name", observed=True)["speaker_id"].agg(lambda x: sorted(set(x)))
    # the two voices never share a speaker label (the number of labels per voice is NMESC's business: on a disconnected
    # neighbour graph upstream's getMinimumConnection walk reads the count off the last candidate, and a random-weight
    # embedding network leaves sub-structure inside a voice)
    assert len(by_stream) == 2 and not (set(by_stream.iloc[0]) & set(by_stream.iloc[1])), by_stream
    assert 2 <= len(set(out.speaker_id)) <= 8
    monkeypatch.delenv("NSF_TITANET_CKPT")
    D._TITANET = None
    with pytest.raises(D._cabi.NsfError if hasattr(D, "_cabi") else Exception):
        D.word_based_clustering(pcm, sr, df, cfg)


@pytest.mark.gpu
def test_titanet_ragged_and_degenerate_crops(dev):
    """Ragged batch with crops at the stream edges (clipped to almost nothing), a single-crop batch, bucketed == one padded batch."""
    import torch
    import notsofar_b200.titanet as T
    w = O.random_weights(3, blocks=SMALL, att_ch=32, emb=16)
    model = T.TitaNetB200(w, dev, precision="fp32")
    rng = np.random.default_rng(9)
    lens = [0, 1, 100, 159, 160, 257, 1000, 8000, 47999]
    crops = _crops(rng, [max(n, 1) for n in lens])
    _, x, _ = _run(model, crops, dev)
    l = torch.tensor(lens, dtype=torch.int32, device=dev)
    hi, lo, nf, t_pad = model.features(x, l)
    assert nf.cpu().tolist() == [0, 1, 1, 1, 2, 2, 7, 51, 300]                 # len // 160 + 1 (0 for an empty crop)
    e_b = model.embed(x, l, bucket=True).cpu().numpy()
    e_p = model.embed(x, l, bucket=False).cpu().numpy()
    assert np.isfinite(e_b).all() and np.isfinite(e_p).all()
    assert rel_l2(e_b, e_p) < 1e-5
    ref = O.embed(w, [c[:n] for c, n in zip(crops, lens) if n >= 257], SMALL)   # reflect padding needs more than n_fft / 2 samples
    got = e_b[[i for i, n in enumerate(lens) if n >= 257]]
    assert max(rel_l2(g, r) for g, r in zip(got, ref)) < 2e-4
    one = model.embed(x[7:8, :8000].contiguous(), l[7:8]).cpu().numpy()
    assert rel_l2(one[0], e_b[7]) < 1e-5
    with pytest.raises(T._cabi.NsfError):
        model.features(x.cpu(), l)


# ------------------------------------------------------------------------------------------------ CPU: oracle vs torch building blocks
def test_oracle_frontend_against_torch_stft():
    """The numpy front end against the torch calls NeMo's FilterbankFeatures is made of [upstream]: torch.stft(n_fft 512, hop 160,
    win_length 400, hann_window(400, periodic=False), center=True -> reflect), |.|^2, mel matmul, log(. + 2^-24), per-feature
    normalisation with the unbiased std."""
    import torch
    rng = np.random.default_rng(21)
    x = _crops(rng, [20000])[0]
    xt = torch.from_numpy(x).double()
    xt = torch.cat([xt[:1], xt[1:] - 0.97 * xt[:-1]])
    spec = torch.stft(xt, n_fft=512, hop_length=160, win_length=400, window=torch.hann_window(400, periodic=False, dtype=torch.float64),
                      center=True, return_complex=True)
    power = spec.abs() ** 2                                                       # [257, T]
    mel = torch.from_numpy(O.mel_filterbank()) @ power
    lm = torch.log(mel + 2.0 ** -24)
    T = O.seq_len(len(x))
    assert lm.shape[1] == T
    lm = lm[:, :T]
    ref = ((lm - lm.mean(1, keepdim=True)) / (lm.std(1, keepdim=True) + 1e-5)).T.numpy()     # torch.std: unbiased
    assert rel_l2(O.features(x), ref) < 1e-9


def test_oracle_network_against_torch_modules():
    """The numpy encoder / decoder against the same network assembled from torch.nn modules the way the NeMo recipe assembles it
    (Conv1d groups=C depthwise + 1x1 pointwise, BatchNorm1d eps 1e-3 in eval mode, squeeze-excite with bias-free Linear layers,
    residual 1x1 conv + BatchNorm, attentive pooling with TDNN(conv, ReLU, BatchNorm) -> Tanh -> Conv1d, BatchNorm + Conv1d embedding),
    loading the oracle's weights by name."""
    import torch
    from torch import nn
    w = O.random_weights(11, blocks=SMALL, att_ch=32, emb=16)
    sd = {k: torch.from_numpy(v).double() for k, v in w.items()}

    def bn(name, c, eps):
        m = nn.BatchNorm1d(c, eps=eps).double()
        m.load_state_dict({k[len(name) + 1:]: v for k, v in sd.items() if k.startswith(name + ".")}, strict=False)
        return m.eval()

    def conv(name, ci, co, k, groups=1, bias=False):
        m = nn.Conv1d(ci, co, k, padding=k // 2, groups=groups, bias=bias).double()
        m.weight.data = sd[name + ".weight"]
        if bias:
            m.bias.data = sd[name + ".bias"]
        return m

    rng = np.random.default_rng(13)
    feats = [rng.standard_normal((T, 80)) for T in (37, 64)]
    got = O.decoder(w, O.encoder(w, feats, SMALL))
    outs = []
    with torch.no_grad():
        for f in feats:
            x = torch.from_numpy(f).T[None]                                         # [1, C, T]
            c_in = 80
            for b, (co, rep, k, res) in enumerate(SMALL):
                p, x_in, c, i = f"encoder.encoder.{b}.", x, c_in, 0
                for r in range(rep):
                    x = conv(p + f"mconv.{i}.conv", c, c, k, groups=c)(x)
                    x = bn(p + f"mconv.{i + 2}", co, 1e-3)(conv(p + f"mconv.{i + 1}.conv", c, co, 1)(x))
                    if r < rep - 1:
                        x = torch.relu(x)
                    i += 3 if r == rep - 1 else 5
                    c = co
                y = x.mean(-1)                                                      # squeeze-excite, context -1
                y = torch.sigmoid(torch.relu(y @ sd[p + f"mconv.{i}.fc.0.weight"].T) @ sd[p + f"mconv.{i}.fc.2.weight"].T)
                x = x * y[..., None]
                if res:
                    x = x + bn(p + "res.0.1", co, 1e-3)(conv(p + "res.0.0.conv", c_in, co, 1)(x_in))
                x = torch.relu(x)
                c_in = co
            q = "decoder._pooling.attention_layer."
            mean = x.mean(-1, keepdim=True)
            std = ((x - mean) ** 2).mean(-1, keepdim=True).clamp(1e-10).sqrt()
            ctx = torch.cat([x, mean.expand_as(x), std.expand_as(x)], 1)
            h = torch.tanh(bn(q + "0.bn", 32, 1e-5)(torch.relu(conv(q + "0.conv_layer", 3 * c_in, 32, 1, bias=True)(ctx))))
            alpha = torch.softmax(conv(q + "2", 32, c_in, 1, bias=True)(h), dim=2)
            mu = (alpha * x).sum(2)
            sg = (alpha * (x - mu[..., None]) ** 2).sum(2).clamp(1e-10).sqrt()
            pool = torch.cat([mu, sg], 1)[..., None]
            e = conv("decoder.emb_layers.0.1", 2 * c_in, 16, 1, bias=True)(bn("decoder.emb_layers.0.0", 2 * c_in, 1e-5)(pool))
            outs.append(e[0, :, 0].numpy())
    assert rel_l2(got, np.stack(outs)) < 1e-10


@pytest.mark.gpu
def test_titanet_vs_nemo_when_available():
    """The pin that cannot run offline (VERDICT r1 missing #5): with NeMo installed and NSF_TITANET_CKPT pointing at titanet_large.nemo,
    the embeddings of word_based_diarization.py:102-105 (``spk_model.forward`` under autocast) must agree with TitaNetB200 on the
    same crops.  Skips -- and says so -- where NeMo or the archive is missing: TitaNet / NMESC parity unpinned."""
    import os
    try:
        from nemo.collections.asr.models import EncDecSpeakerLabelModel
    except Exception:
        pytest.skip("NeMo is not installed: TitaNet / NMESC parity stays unpinned against upstream (SURVEY 8c)")
    path = os.environ.get("NSF_TITANET_CKPT")
    if not path or not os.path.exists(path):
        pytest.skip("NSF_TITANET_CKPT not set: TitaNet parity stays unpinned against upstream")
    import torch
    from notsofar_b200.titanet import load_titanet
    dev = torch.device("cuda", 0)
    ref = EncDecSpeakerLabelModel.restore_from(path, map_location=dev).eval()
    mine = load_titanet(path, dev)
    rng = np.random.default_rng(0)
    lens = torch.tensor([48000, 32000, 17000, 8000], device=dev)
    crops = torch.from_numpy((rng.standard_normal((4, 48000)) * 0.05).astype(np.float32)).to(dev)
    for i, n in enumerate(lens.tolist()):
        crops[i, n:] = 0
    with torch.no_grad(), torch.autocast("cuda"):
        _, emb_ref = ref.forward(input_signal=crops, input_signal_length=lens)
    emb = mine.as_embedding_backend()(crops, lens, None)
    cos = torch.nn.functional.cosine_similarity(emb.float(), emb_ref.float(), dim=-1)
    print("TitaNetB200 vs NeMo (autocast fp16): cosine similarity per crop", cos.tolist())
    assert float(cos.min()) > 0.999
