"""First-party diarization / ASR plug-in logic against fixtures produced by the reference itself
(tests/golden/make_golden_diar.py): post-processing of diarization_common.py, the integer crop plan of
word_based_diarization.py:78-101 (bit-exact), the module surface of asr.py / diarization.py, and -- on a GPU -- the
device-side crop gather."""
import dataclasses
import json
import os

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def G():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "diar_golden.json")))


@pytest.fixture(scope="module")
def D():
    from notsofar_b200 import diarization
    return diarization


def _segments(seed, wav_seconds=40.0):
    import importlib.util
    spec = importlib.util.spec_from_file_location("mkdiar", os.path.join(ROOT, "tests", "golden", "make_golden_diar.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.synth_segments(seed, wav_seconds=wav_seconds)


def test_overlap_ratio_known_answers(D, G):
    for a, b, c, d, ref in G["overlap"]:
        assert D.compute_overlap_ratio(a, b, c, d) == ref


def test_post_processing_matches_reference(D, G):
    for case in G["post"]:
        df = _segments(case["seed"])
        res = D.prepare_diarized_data_frame([list(w) for w in case["words"]], df, case["dedup"])
        ref = case["segments"]
        assert res.start_time.tolist() == ref["start_time"] and res.end_time.tolist() == ref["end_time"]
        assert res.text.tolist() == ref["text"] and res.speaker_id.tolist() == ref["speaker_id"]
        assert res.wav_file_name.tolist() == ref["wav_file_name"]
        assert json.loads(json.dumps(res.word_timing.tolist())) == ref["word_timing"]
        assert set(res.columns) == {"start_time", "end_time", "text", "word_timing", "meeting_id", "session_id", "wav_file_name", "speaker_id"}


def test_deduplicate_drops_first_word_like_the_reference(D):
    words = [["a", 0.0, 1.0, 0, "spk0"], ["b", 1.0, 2.0, 0, "spk0"], ["b", 1.05, 2.0, 1, "spk0"], ["c", 2.0, 3.0, 0, "spk1"]]
    assert D.deduplicate(words) == [words[1], words[3]]


def test_crop_plan_bit_exact(D, G):
    for case in G["crops"]:
        df = _segments(case["seed"], wav_seconds=min(case["n_samples"] / case["sr"], 40.0))
        plan = D.word_crop_plan(df, case["n_samples"], case["sr"], case["windows"], 3)
        ref = np.asarray(case["crops"], np.int64)
        assert len(plan.start) == len(ref)
        nz = ref[:, 2] > 0                                          # empty crops carry no position in the recording
        assert np.array_equal(plan.length, ref[:, 2])
        assert np.array_equal(plan.start[nz], ref[nz, 1]) and np.array_equal(plan.stream_id[nz], ref[nz, 0])
        kept = [[w[0], w[1], w[2], int(w[3])] for w, t in zip(plan.words, plan.too_long) if not t]
        assert kept == case["kept_words"]
        assert plan.too_long.sum() >= 1


def test_cfg_and_signatures_match_the_reference(D):
    import inspect
    from notsofar_b200 import asr
    assert [f.name for f in dataclasses.fields(D.DiarizationCfg)] == [
        "method", "min_embedding_windows", "max_allowed_word_duration", "apply_deduplication", "embedding_model_name",
        "msdd_model_name", "vad_model_name"]
    assert [f.name for f in dataclasses.fields(asr.WhisperAsrCfg)] == [
        "model_name", "language", "word_level_time_stamps", "beam_size", "hallucination_silence_threshold"]
    assert list(inspect.signature(D.diarization_inference).parameters)[:5] == ["out_dir", "segments_df", "cfg", "fetch_from_cache", "device"]
    assert list(inspect.signature(asr.asr_inference).parameters)[:4] == ["out_dir", "session", "cfg", "fetch_from_cache"]
    asr.WhisperAsrCfg().assert_valid()
    with pytest.raises(AssertionError):
        asr.WhisperAsrCfg(model_name="huge").assert_valid()


def test_diarization_inference_modes_and_asr_frames(D, tmp_path):
    from notsofar_b200 import asr, NsfError
    df = _segments(0).drop(columns=["wav_file_name_ind"])
    df["wav_file_name"] = df["wav_file_name"].astype(str)
    out = D.diarization_inference(str(tmp_path), df, D.DiarizationCfg(method="skip"), False)
    assert (out.speaker_id == "spk0").all() and "speaker_id" not in df
    out = D.diarization_inference(str(tmp_path), df, D.DiarizationCfg(method="by_wav_file_name"), False)
    assert sorted(out.speaker_id.unique()) == [f"wav_{i}" for i in range(df.wav_file_name.nunique())]
    with pytest.raises(NsfError):
        D.diarization_inference(str(tmp_path), df, D.DiarizationCfg(method="word_nmesc", min_embedding_windows=[1.0]), False, pcm=object())
    # asr: a transcriber plug-in returning whisper-shaped results -> the reference's segments_df layout + cache file
    session = pd.Series(dict(meeting_id="MTG_1", session_id="multichannel/MTG_1_dev", sep_wav_file_names=["/x/s0.wav", "/x/s1.wav"]))
    with pytest.raises(NsfError):
        asr.asr_inference(str(tmp_path), session, asr.WhisperAsrCfg(model_name="tiny"), False)

    def fake(stream, cfg, options):
        assert options["beam_size"] == 5 and options["task"] == "transcribe"
        if stream.endswith("s1.wav"):
            return {"segments": []}
        return {"segments": [{"start": 0.0, "end": 1.0, "text": " hi there", "words": [{"word": " hi", "start": 0.0, "end": 0.4},
                                                                                     {"word": " there", "start": 0.5, "end": 1.0}]}]}
    asr.set_transcriber(fake)
    try:
        seg = asr.asr_inference(str(tmp_path), session, asr.WhisperAsrCfg(model_name="tiny"), False)
    finally:
        asr.set_transcriber(None)
    assert list(seg.columns) == ["start_time", "end_time", "text", "word_timing", "meeting_id", "session_id", "wav_file_name"]
    assert seg.word_timing[0] == [[" hi", 0.0, 0.4], [" there", 0.5, 1.0]] and len(seg) == 1
    assert (tmp_path / "asr" / "multichannel/MTG_1_dev" / "tiny" / "all_segments_df.pkl").exists()
    again = asr.asr_inference(str(tmp_path), session, asr.WhisperAsrCfg(model_name="tiny"), True)      # cache hit, no transcriber
    assert again.equals(seg)


@pytest.mark.gpu
def test_gather_word_crops_on_device(D, G):
    """nsf_gather_crops against the reference's crops (ramp streams: content identifies stream, start, padding)."""
    import torch
    case = G["crops"][0]
    n, sr = case["n_samples"], case["sr"]
    dev = torch.device("cuda", 0)
    ramp = (np.arange(n) % 30000).astype(np.int16)
    pcm = torch.from_numpy(np.stack([ramp, -ramp, ramp // 2])).to(dev)
    df = _segments(case["seed"])
    plan = D.word_crop_plan(df, n, sr, case["windows"], 3)
    got, lens = D.gather_word_crops(pcm, plan)
    got, lens = got.cpu().numpy(), lens.cpu().numpy()
    host = pcm.cpu().numpy()
    assert np.array_equal(lens, plan.length)
    for i in range(len(plan.start)):
        ref = host[plan.stream_id[i], plan.start[i]:plan.start[i] + plan.length[i]].astype(np.float32) / np.float32(32767)
        assert np.array_equal(got[i, :plan.length[i]], ref) and not got[i, plan.length[i]:].any()
    # batched like the reference (32 words x scales per call)
    step = 8 * len(case["windows"])
    assert len(plan.start) > step
    part, _ = D.gather_word_crops(pcm, plan, step, min(step, len(plan.start) - step))
    assert np.array_equal(part.cpu().numpy()[:, :1], got[step:step + part.shape[0], :1])


# ----------------------------------------------------------------------------------------------- boundary breaks of round 1
def test_text_normalizer_is_the_reference_one_when_the_host_repo_is_importable(monkeypatch):
    """asr.py:23-25 / inference_pipeline/inference.py:75,87: the unchanged caller invokes cfg.asr.text_normalizer() after
    diarization; it must hand back the host repository's chime8 normaliser, not raise."""
    import sys
    from notsofar_b200.asr import WhisperAsrCfg
    from oracle import reference_shim as R
    cfg = WhisperAsrCfg()
    if R.available():
        monkeypatch.syspath_prepend(R.REFERENCE_ROOT)
        try:
            import more_itertools  # noqa: F401  (a dependency of the reference's normaliser, requirements.txt; absent in this image)
        except ImportError:
            import types
            stub = types.ModuleType("more_itertools")
            stub.windowed = lambda seq, n: (tuple(seq[i:i + n]) for i in range(max(len(seq) - n + 1, 1)))
            monkeypatch.setitem(sys.modules, "more_itertools", stub)
        for m in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
            monkeypatch.delitem(sys.modules, m, raising=False)
        norm = cfg.text_normalizer()
        assert type(norm).__name__ == "EnglishTextNormalizer"
        assert norm("Mr. Smith's  twenty-two dogs, um, okay") == norm("mister smith's 22 dogs okay")
    else:
        monkeypatch.setattr(sys, "path", [p for p in sys.path if "reference" not in p])
        for m in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
            monkeypatch.delitem(sys.modules, m, raising=False)
        with pytest.raises(ImportError, match="text_norm_whisper_like"):
            cfg.text_normalizer()


def test_train_cfg_node_has_attribute_access():
    from notsofar_b200.css import CfgNode
    c = CfgNode({"single_channel": False, "conformer_css_cfg": {"nnet_conf": {"num_spks": 3, "conformer_conf": {"attention_dim": 512}}}})
    assert c.single_channel is False and c.conformer_css_cfg.nnet_conf.conformer_conf.attention_dim == 512
    assert c["conformer_css_cfg"]["nnet_conf"]["num_spks"] == 3
    with pytest.raises(AttributeError):
        c.missing


def test_diarization_skips_cache_under_several_ranks(D, tmp_path, monkeypatch):
    """diarization.py:82-89,104-106: with world_size > 1 neither the cache read nor the pickle write happens."""
    import notsofar_b200.diarization as dm
    df = _segments(3).drop(columns=["wav_file_name_ind"], errors="ignore")
    df["wav_file_name"] = df["wav_file_name"].astype(str)
    calls = []
    monkeypatch.setattr(dm, "_load_streams_as_pcm", lambda files, device: ("pcm", 16000))
    monkeypatch.setattr(dm, "word_based_clustering", lambda pcm, sr, seg, cfg: (calls.append((pcm, sr)), seg.assign(speaker_id="spk0"))[1])
    cfg = dm.DiarizationCfg(method="word_nmesc")
    monkeypatch.setattr(dm, "_world_size", lambda: 2)
    out = dm.diarization_inference(str(tmp_path), df, cfg, fetch_from_cache=True)
    assert len(calls) == 1 and calls[0] == ("pcm", 16000) and "speaker_id" in out
    assert not (tmp_path / "diarization").exists()
    monkeypatch.setattr(dm, "_world_size", lambda: 1)
    dm.diarization_inference(str(tmp_path), df, cfg, fetch_from_cache=True)
    assert len(list((tmp_path / "diarization").rglob("all_segments_df.pkl"))) == 1
    dm.diarization_inference(str(tmp_path), df, cfg, fetch_from_cache=True)          # served from the cache now
    assert len(calls) == 2
