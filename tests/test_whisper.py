"""Whisper log-mel front end + audio encoder (row a15): the numpy oracle against the transformers implementation of the
published model (the offline pin; openai-whisper itself is absent -> parity unpinned upstream), and -- on a GPU -- the
tcgen05 encoder (csrc/whisper.cu) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import whisper_oracle as WO

from conftest import rel_l2


@pytest.fixture(autouse=True)
def _drop_reference_import_stubs():
    """oracle/reference_shim.py leaves spec-less stand-ins for librosa / soundfile in sys.modules (earlier tests of the same
    process); transformers probes optional packages with importlib.util.find_spec, which rejects those."""
    import sys
    for name in ("librosa", "soundfile", "sounddevice"):
        mod = sys.modules.get(name)
        if mod is not None and getattr(mod, "__spec__", None) is None:
            del sys.modules[name]
    yield


def _hf_encoder(d_model, layers, heads, ffn, n_mels, seed=0, gain=1.0):
    from transformers import WhisperConfig
    from transformers.models.whisper.modeling_whisper import WhisperEncoder
    cfg = WhisperConfig(d_model=d_model, encoder_layers=layers, encoder_attention_heads=heads, encoder_ffn_dim=ffn, num_mel_bins=n_mels)
    torch.manual_seed(seed)
    m = WhisperEncoder(cfg).eval()
    with torch.no_grad():
        for n, p in m.named_parameters():                       # away from the bland default init: exercise every term
            if p.dim() >= 2 and "embed_positions" not in n:
                p.mul_(gain)
            elif "bias" in n:
                p.normal_(0, 0.1)
            elif "layer_norm" in n and "weight" in n:
                p.add_(0.1 * torch.randn_like(p))
    return m, {"model.encoder." + k: v for k, v in m.state_dict().items()}


def _audio(seed, n=480000):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 16000.0
    a = 0.1 * np.sin(2 * np.pi * (200 + 50 * seed) * t) * (np.sin(2 * np.pi * 0.7 * t) > 0) + 0.02 * rng.standard_normal(n)
    a[n // 2:] *= 0.01                                           # a quiet half: the max - 8 clamp is active
    return a.astype(np.float32)


def test_mel_filterbank_and_logmel_oracle_vs_transformers():
    from transformers import WhisperFeatureExtractor
    from notsofar_b200.whisper import mel_filterbank
    for n_mels in (80, 128):
        fe = WhisperFeatureExtractor(feature_size=n_mels)
        f = mel_filterbank(n_mels)
        assert np.abs(f - fe.mel_filters.T).max() < 1e-7
        a = _audio(n_mels)
        ref = fe(a, sampling_rate=16000, return_tensors="np")["input_features"][0]
        assert ref.shape == (n_mels, 3000)
        assert np.abs(WO.log_mel(a, f) - ref).max() < 1e-4


def test_encoder_oracle_vs_transformers():
    from notsofar_b200.whisper import _canon, pack_whisper_encoder
    m, sd = _hf_encoder(128, 2, 2, 512, 80, gain=3.0)
    mel = np.random.default_rng(1).standard_normal((80, 3000)).astype(np.float32) * 0.5
    with torch.no_grad():
        ref = m(torch.from_numpy(mel[None])).last_hidden_state[0].numpy()
    got = WO.encoder(_canon(sd), mel)
    assert rel_l2(got, ref) < 1e-5
    dims, blob, offs, filt = pack_whisper_encoder(sd)
    assert (dims.n_mels, dims.n_ctx, dims.d_model, dims.n_heads, dims.n_layers, dims.d_ff) == (80, 1500, 128, 2, 2, 512)
    assert len(offs) == 8 + 12 * 2 and filt.shape == (80, 201) and (offs % 4 == 0).all()
    # openai-whisper naming is accepted too
    oa = {"encoder." + k: v for k, v in _canon(sd).items()}
    _, blob2, offs2, _ = pack_whisper_encoder(oa)
    assert np.array_equal(blob.view(np.uint32), blob2.view(np.uint32)) and np.array_equal(offs, offs2)


@pytest.mark.gpu
def test_logmel_kernel_vs_oracle():
    from notsofar_b200.whisper import WhisperEncoderB200, mel_filterbank
    dev = torch.device("cuda", 0)
    _, sd = _hf_encoder(128, 1, 2, 256, 80)
    enc = WhisperEncoderB200(sd, device=dev)
    a = np.stack([_audio(0), _audio(1)])
    hi, lo, _ = enc.log_mel(torch.from_numpy(a).to(dev))
    got = (hi.view(torch.bfloat16).float() + lo.view(torch.bfloat16).float()).cpu().numpy()       # [B, 3002, n_mels]
    assert not got[:, 0].any() and not got[:, 3001].any()
    f = mel_filterbank(80)
    for b in range(2):
        ref = WO.log_mel(a[b], f)
        err = np.abs(got[b, 1:3001].T - ref).max()
        print("log-mel max abs err", err)
        assert err < 2e-4


@pytest.mark.gpu
@pytest.mark.parametrize("d_model,layers,heads,ffn,n_mels,batch", [(128, 2, 2, 512, 80, 1), (384, 4, 6, 1536, 80, 2), (1280, 2, 20, 5120, 128, 1)])
def test_encoder_kernel_vs_oracle(d_model, layers, heads, ffn, n_mels, batch):
    """bf16 tensor-core encoder (fp32 residual stream) vs the float64 oracle on the same weights and features."""
    from notsofar_b200.whisper import WhisperEncoderB200, _canon
    dev = torch.device("cuda", 0)
    _, sd = _hf_encoder(d_model, layers, heads, ffn, n_mels, gain=2.0)
    enc = WhisperEncoderB200(sd, device=dev)
    rng = np.random.default_rng(d_model)
    mel = (rng.standard_normal((batch, n_mels, 3000)) * 0.5).astype(np.float32)
    out = enc.encode_mel_f32(torch.from_numpy(mel).to(dev)).cpu().numpy()
    assert np.isfinite(out).all()
    w = _canon(sd)
    for b in range(batch):
        ref = WO.encoder(w, mel[b], dtype=np.float32 if d_model > 512 else np.float64)
        err = rel_l2(out[b], ref)
        print(f"whisper encoder d={d_model} L={layers}: rel_l2 vs oracle = {err:.3e}")
        assert err < 2e-2                                        # bf16 operands (2^-9 per element), fp32 accumulate / residual
