"""Drop-in counterpart of the reference's ASR plug-in (asr/asr.py): ``WhisperAsrCfg`` and ``asr_inference`` with the
reference's signature, cache file, per-stream loop and segments_df layout (asr.py:31-101).

The transcription itself is openai-whisper in the reference (``whisper.load_model`` / ``model.transcribe``, asr.py:69-74):
a third-party package that is absent offline together with its weights and vocabulary (parity unpinned).  Here it
is ``whisper_asr.WhisperB200Transcriber`` -- log-mel, encoder, beam-search / fallback decoding, word timestamps on the in-tree
kernels -- built from ``NSF_WHISPER_CKPT`` + ``NSF_WHISPER_VOCAB`` on first use, or registered explicitly with
``set_transcriber``; a transcriber gets a stream (a WAV path, or the device-resident PCM16 stream the CSS stage left in HBM)
and returns whisper's result dict ``{'segments': [{'start', 'end', 'text', 'words': [{'word', 'start', 'end'}, ...]}, ...]}``.
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Callable, Optional

import pandas as pd


@dataclass
class WhisperAsrCfg:
    """asr.py:15-28, field for field."""
    model_name: str = 'large-v2'  # use 'large-v2' for experiments, use 'tiny' for fast debugging
    language: Optional[str] = 'en'  # language that the speech is in (if None, whisper runs language ID)
    word_level_time_stamps: bool = True
    beam_size: Optional[int] = 5
    hallucination_silence_threshold: Optional[float] = 2.

    def text_normalizer(self):
        """asr.py:23-25: the chime8 normaliser of the host repository (utils/text_norm_whisper_like, a scoring component that
        is not rebuilt here).  The unchanged caller invokes this after diarization (inference_pipeline/inference.py:75,87), in
        a process whose sys.path holds the reference checkout, so the reference's own module is what gets returned."""
        try:
            from utils.text_norm_whisper_like import get_txt_norm
        except ImportError as e:
            raise ImportError("WhisperAsrCfg.text_normalizer() returns the reference's chime8 normaliser "
                              "(utils.text_norm_whisper_like.get_txt_norm); run from the reference checkout or put it on sys.path") from e
        return get_txt_norm("chime8")

    def assert_valid(self):
        assert self.model_name in ['tiny.en', 'tiny', 'base.en', 'base', 'small.en', 'small', 'medium.en',
                                   'medium', 'large-v1', 'large-v2', 'large-v3', 'large']


_TRANSCRIBER: Optional[Callable] = None     # (stream, cfg: WhisperAsrCfg, options: dict) -> whisper result dict


def set_transcriber(fn: Optional[Callable]):
    global _TRANSCRIBER
    _TRANSCRIBER = fn


def segments_frame(results: dict, session, wav_file) -> Optional[pd.DataFrame]:
    """asr.py:75-96: whisper's per-stream result -> the segments_df rows of that stream (None when empty)."""
    if len(results['segments']) == 0:
        return None
    raw = pd.DataFrame(results['segments'])
    word_start_end = raw['words'].apply(lambda x: [[w['word'], w['start'], w['end']] for w in x])
    df = pd.DataFrame({'start_time': raw['start'], 'end_time': raw['end'], 'text': raw['text'], 'word_timing': word_start_end})
    df['meeting_id'] = session.meeting_id
    df['session_id'] = session.session_id
    df['wav_file_name'] = wav_file
    return df


def asr_inference(out_dir: str, session: pd.Series, cfg: WhisperAsrCfg, fetch_from_cache: bool, streams=None):
    """Same contract as the reference's asr_inference (asr.py:31-101).  ``streams`` (optional, extension): one
    device-resident stream per entry of session.sep_wav_file_names, handed to the transcriber instead of the path."""
    cfg.assert_valid()
    options = dict(task="transcribe", language=cfg.language, word_timestamps=cfg.word_level_time_stamps,
                   beam_size=cfg.beam_size, hallucination_silence_threshold=cfg.hallucination_silence_threshold)
    wav_files = session.sep_wav_file_names
    assert isinstance(wav_files, list)
    out_file = Path(out_dir) / 'asr' / session.session_id / cfg.model_name / "all_segments_df.pkl"
    if fetch_from_cache and out_file.exists():
        return pd.read_pickle(out_file)
    transcriber = _TRANSCRIBER
    if transcriber is None:
        from .whisper_asr import transcriber_from_env
        transcriber = transcriber_from_env()              # NSF_WHISPER_CKPT + NSF_WHISPER_VOCAB
        if transcriber is None:
            from ._cabi import NsfError
            raise NsfError("asr_inference needs Whisper weights and vocabulary (the reference downloads them through openai-whisper, asr.py:69; "
                           "there is no network here): set NSF_WHISPER_CKPT and NSF_WHISPER_VOCAB, or register a transcriber with "
                           "notsofar_b200.asr.set_transcriber")
        set_transcriber(transcriber)
    if streams is None:
        # the separated streams the CSS stage of this process left in HBM (the very samples of the WAV files), else the files
        from .css import device_streams_for, flush_wav_writes
        hit = device_streams_for(wav_files)
        if hit is not None and hit[1] == 16000:
            streams = [hit[0][i] for i in range(len(wav_files))]
        else:
            flush_wav_writes(wav_files)
    dfs = []
    for i, wav_file in enumerate(wav_files):
        results = transcriber(streams[i] if streams is not None else str(wav_file), cfg, options)
        df = segments_frame(results, session, wav_file)
        if df is not None:
            dfs.append(df)
    # (the reference's pd.concat raises on an all-silent session; an empty frame with the same columns is returned instead)
    all_segments_df = pd.concat(dfs, ignore_index=True) if dfs else pd.DataFrame(
        columns=['start_time', 'end_time', 'text', 'word_timing', 'meeting_id', 'session_id', 'wav_file_name'])
    out_file.parent.mkdir(parents=True, exist_ok=True)
    all_segments_df.to_pickle(out_file)
    return all_segments_df
