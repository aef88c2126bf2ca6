"""Drop-in counterpart of the reference's diarization plug-in (diarization/diarization.py,
diarization/diarization_common.py, diarization/word_based_diarization.py) -- first-party logic only.

What the reference computes itself, and is mirrored here with the same names, arguments and results:
  * DiarizationCfg                                  diarization_common.py:8-17
  * compute_overlap_ratio / deduplicate / merge_words_to_segments_by_spk_change / prepare_diarized_data_frame
                                                    diarization_common.py:20-102 (incl. the quirk that deduplicate
                                                    never emits word 0, :57-60)
  * the integer crop plan of extract_speaker_embedding_for_words (word_based_diarization.py:78-101): per word and
    per scale, ``int(t * sr)`` sample indices of the word itself or of a window centred on it -- bit-exact
    (``word_crop_plan``), plus the CSS -> diarization hand-off on the GPU: the PCM16 streams the CSS stage produced
    stay in HBM and ``gather_word_crops`` cuts / pads the batches there (libnsf_b200.so: nsf_gather_crops)
  * diarization_inference                           diarization.py:15-109 (modes, cache file, category codes)

What the reference takes from NeMo (third-party, unpinned, absent offline -- SURVEY 8c) is built from the published
algorithms and checked against this repository's own restatements only (**parity unpinned**):
  * TitaNet ``EncDecSpeakerLabelModel.forward``     titanet.py / csrc/titanet.cu (weights: NSF_TITANET_CKPT -> .nemo archive)
  * getCosAffinityMatrix + mean over scales         titanet.multiscale_affinity (CUDA)
  * NMESC + SpectralClustering (``run_clustering``)  clustering.py (torch; eigendecompositions are library calls)
``set_embedding_backend`` / ``set_clustering_backend`` replace either stage.  There is no CPU fallback for the GPU pieces.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from pathlib import Path
from typing import Callable, List, Optional, Sequence

import numpy as np
import pandas as pd


# diarization inference configuration -- field for field diarization_common.py:8-17
@dataclass
class DiarizationCfg:
    method: str = "nmesc"       # choose from "nmesc", "nmesc_msdd", "word_nmesc", or "skip"
    min_embedding_windows: list = field(default_factory=list)
    max_allowed_word_duration: float = 3    # maximum allowed word duration. If word is longer than this value, ignore it.
    apply_deduplication: bool = True
    embedding_model_name: str = "titanet_large"
    msdd_model_name: str = "diar_msdd_telephonic"
    vad_model_name: str = "vad_multilingual_marblenet"   # 16kHz


# ------------------------------------------------------------------------------------------- diarization_common.py
def merge_words_to_segments_by_spk_change(all_words: list):
    """diarization_common.py:20-41: a new segment starts whenever the speaker id (last field) or the stream id
    (second to last) changes.  Words are [text, start, end, stream, speaker]."""
    if len(all_words) == 0:
        return []
    if len(all_words) == 1:
        return all_words
    segments = {"word_timing": [], "speaker_id": []}
    seg_start = 0
    for i, word in enumerate(all_words):
        if i > 0 and (word[-1] != all_words[seg_start][-1] or word[-2] != all_words[seg_start][-2]):
            seg_words = all_words[seg_start:i]
            segments["word_timing"].append([w[:-1] for w in seg_words])
            segments["speaker_id"].append(seg_words[0][-1])
            seg_start = i
    segments["word_timing"].append([w[:-1] for w in all_words[seg_start:]])
    segments["speaker_id"].append(all_words[seg_start][-1])
    return segments


def compute_overlap_ratio(start1, end1, start2, end2):
    """diarization_common.py:44-56: overlap / longer duration, 0 when disjoint."""
    overlap = min(end1, end2) - max(start1, start2)
    if overlap < 0:
        return 0
    return overlap / max(end1 - start1, end2 - start2)


def deduplicate(all_words_sorted, overlap_threshold=0.5):
    """diarization_common.py:59-77: drops a word that repeats its predecessor (same text, same speaker, > 50 %
    overlap).  As in the reference, the loop starts at the second word, so word 0 is never emitted."""
    out = []
    for i in range(1, len(all_words_sorted)):
        cur, prev = all_words_sorted[i], all_words_sorted[i - 1]
        skip = False
        if cur[0] == prev[0] and cur[4] == prev[4]:
            if compute_overlap_ratio(cur[1], cur[2], prev[1], prev[2]) > overlap_threshold:
                skip = True
        if not skip:
            out.append(cur)
    return out


def prepare_diarized_data_frame(all_words, segments_df, apply_deduplication):
    """diarization_common.py:80-102: sort by end time, de-duplicate, cut into speaker / stream runs."""
    all_words_sorted = sorted(all_words, key=lambda x: x[2])
    final_words = deduplicate(all_words_sorted) if apply_deduplication else all_words_sorted
    segments = merge_words_to_segments_by_spk_change(final_words)
    diarized = pd.DataFrame(
        {'start_time': [seg[0][1] for seg in segments["word_timing"]],
         'end_time': [seg[-1][2] for seg in segments["word_timing"]],
         'text': ["".join([w[0] for w in seg]) for seg in segments["word_timing"]],
         'word_timing': segments["word_timing"]})
    diarized['meeting_id'] = segments_df['meeting_id'][0]
    diarized['session_id'] = segments_df['session_id'][0]
    stream_id = [seg[0][-1] for seg in diarized.word_timing.to_list()]
    diarized['wav_file_name'] = segments_df['wav_file_name'].cat.categories[stream_id]
    diarized['speaker_id'] = segments["speaker_id"]
    return diarized


# ------------------------------------------------------------------------------------------- word crops
@dataclass
class CropPlan:
    """One row per (word, scale), word-major like the reference's batches (word_based_diarization.py:78-101)."""
    stream_id: np.ndarray     # int32 [n]   unmixed channel of the word's segment (wav_file_name_ind)
    start: np.ndarray         # int64 [n]   first sample
    length: np.ndarray        # int32 [n]   samples (after clipping at the end of the stream, as slicing does)
    word_index: np.ndarray    # int32 [n]   index into ``words``
    words: list               # [[text, start, end, stream_id], ...] in segment order
    too_long: np.ndarray      # bool [n_words]  word_duration > max_allowed_word_duration (dropped later, :118-124)


def word_crop_plan(segments_df: pd.DataFrame, n_samples: int, sr: int, min_embedding_windows: Sequence[float],
                   max_allowed_word_duration: float = 3) -> CropPlan:
    """Integer sample ranges of every (word, scale) crop, exactly as word_based_diarization.py:78-101 computes them
    (float64 arithmetic on Python floats, ``int()`` truncation, slice clipping at the stream end)."""
    wav_duration = n_samples / sr
    sid, st, ln, wi, words, too_long = [], [], [], [], [], []
    for _, seg in segments_df.iterrows():
        channel_id = seg.wav_file_name_ind
        for word in seg["word_timing"]:
            start_time, end_time = word[1], word[2]
            center_time = (start_time + end_time) / 2
            word_duration = end_time - start_time
            for min_window_size in min_embedding_windows:
                if word_duration < min_window_size:
                    start_time2 = np.maximum(0, center_time - min_window_size / 2)
                    end_time2 = np.minimum(wav_duration, center_time + min_window_size / 2)
                    a, b = int(start_time2 * sr), int(end_time2 * sr)
                else:
                    a, b = int(start_time * sr), int(end_time * sr)
                # wavs[channel_id][a:b] -- Python slice semantics (negative indices count from the end)
                a_, b_, _ = slice(a, b).indices(n_samples)
                sid.append(int(channel_id)); st.append(a_); ln.append(max(0, b_ - a_)); wi.append(len(words))
            words.append(list(word) + [channel_id])
            too_long.append(word_duration > max_allowed_word_duration)
    return CropPlan(np.asarray(sid, np.int32), np.asarray(st, np.int64), np.asarray(ln, np.int32), np.asarray(wi, np.int32),
                    words, np.asarray(too_long, bool))


def gather_word_crops(pcm, plan: CropPlan, first: int = 0, count: Optional[int] = None):
    """pcm: int16 CUDA tensor [n_streams, n] (the CSS output after nsf_peaknorm_pcm16).  Returns (crops [count, max_len]
    float32 zero padded, lengths int32 [count]) for plan rows [first, first + count) -- the device-side twin of the
    reference's read_wav(normalize=True) + slicing + pad_sequence."""
    import torch
    from . import _cabi
    if not (isinstance(pcm, torch.Tensor) and pcm.is_cuda and pcm.dtype == torch.int16 and pcm.dim() == 2 and pcm.is_contiguous()):
        raise _cabi.NsfError("gather_word_crops needs a contiguous int16 CUDA tensor [n_streams, n]; there is no CPU path")
    lib = _cabi.load()
    count = len(plan.start) - first if count is None else count
    sl = slice(first, first + count)
    dev = pcm.device
    sid = torch.from_numpy(np.ascontiguousarray(plan.stream_id[sl])).to(dev)
    st = torch.from_numpy(np.ascontiguousarray(plan.start[sl])).to(dev)
    ln = torch.from_numpy(np.ascontiguousarray(plan.length[sl])).to(dev)
    max_len = int(plan.length[sl].max()) if count else 0
    out = torch.empty((count, max_len), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(lib.nsf_gather_crops(_cabi.ptr(pcm), pcm.shape[0], pcm.shape[1], _cabi.ptr(sid), _cabi.ptr(st), _cabi.ptr(ln),
                                         count, max_len, _cabi.ptr(out), _cabi.stream_ptr()), "nsf_gather_crops")
    return out, ln


# ------------------------------------------------------------------------------------------- plug-in points
_EMBEDDING_BACKEND: Optional[Callable] = None     # (crops [n, L] f32 cuda, lengths [n] i32 cuda, cfg) -> [n, D] embeddings
_CLUSTERING_BACKEND: Optional[Callable] = None    # (embeddings [n_words, n_scales, D], cfg) -> int labels [n_words]


def set_embedding_backend(fn: Optional[Callable]):
    global _EMBEDDING_BACKEND
    _EMBEDDING_BACKEND = fn


def set_clustering_backend(fn: Optional[Callable]):
    global _CLUSTERING_BACKEND
    _CLUSTERING_BACKEND = fn


def _embedding_backend(cfg: DiarizationCfg, device):
    """The registered backend, else TitaNetB200 on the checkpoint named by NSF_TITANET_CKPT (the reference downloads
    ``cfg.embedding_model_name`` from NGC, word_based_diarization.py:21-29; there is no network here)."""
    if _EMBEDDING_BACKEND is not None:
        return _EMBEDDING_BACKEND
    from . import _cabi
    path = os.environ.get("NSF_TITANET_CKPT")
    if not path:
        raise _cabi.NsfError(f"word_nmesc needs the weights of '{cfg.embedding_model_name}': set NSF_TITANET_CKPT to its .nemo archive "
                             "(or a torch state_dict file), or register a backend with set_embedding_backend")
    global _TITANET
    if _TITANET is None or _TITANET[0] != (path, str(device)):
        from .titanet import load_titanet
        _TITANET = ((path, str(device)), load_titanet(path, device))
    return _TITANET[1].as_embedding_backend()


_TITANET = None


def word_shard_bounds(n_words: int, world: int) -> List[int]:
    """Contiguous, balanced blocks of words per rank: rank r embeds words [bounds[r], bounds[r + 1])."""
    return [n_words * r // world for r in range(world + 1)]


def word_based_clustering(pcm, sr: int, segments_df: pd.DataFrame, cfg: DiarizationCfg, batch_words: int = 256, group=None,
                          shard_words: bool = False):
    """word_based_diarization.py:135-189 on device-resident streams: crops -> TitaNet embeddings (csrc/titanet.cu) -> multi-scale
    cosine affinity -> NMESC + spectral clustering (clustering.py) -> prepare_diarized_data_frame.  ``pcm`` int16 CUDA tensor
    [n_streams, n].  Both stages can be replaced through set_embedding_backend / set_clustering_backend.  Embeddings do not
    depend on the batch composition (padding is masked), so batches are larger than the reference's 32 words.

    ``shard_words``: with torch.distributed initialised, every rank (holding the same ``pcm`` and ``segments_df``)
    embeds a contiguous block of the words and one all-gather of the [n_words, n_scales, D] embeddings precedes the affinity and
    clustering, which every rank repeats identically (O(n_words^2), small next to the embedding forward)."""
    embed_fn = _embedding_backend(cfg, getattr(pcm, "device", None))
    if _CLUSTERING_BACKEND is not None:
        cluster_fn = _CLUSTERING_BACKEND
    else:
        from .clustering import nmesc_backend as cluster_fn
    import torch
    import torch.distributed as dist
    n_scales = len(cfg.min_embedding_windows)
    plan = word_crop_plan(segments_df, pcm.shape[1], sr, cfg.min_embedding_windows, cfg.max_allowed_word_duration)
    n_words = len(plan.words)
    world = dist.get_world_size(group) if (shard_words and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    bounds = word_shard_bounds(n_words, world)
    lo, hi = bounds[rank] * n_scales, bounds[rank + 1] * n_scales
    embs = []
    step = batch_words * n_scales
    for first in range(lo, hi, step):
        crops, lens = gather_word_crops(pcm, plan, first, min(step, hi - first))
        embs.append(embed_fn(crops, lens, cfg))
    if world > 1:
        from .sharded import allgather_varlen
        own = torch.cat(embs, 0) if embs else None
        d = torch.tensor([own.shape[1] if own is not None else 0], device=pcm.device)
        dist.all_reduce(d, op=dist.ReduceOp.MAX, group=group)
        if own is None:
            own = torch.zeros((0, int(d.item())), dtype=torch.float32, device=pcm.device)
        emb = allgather_varlen(own.float().contiguous(), [(bounds[r + 1] - bounds[r]) * n_scales for r in range(world)], group)
    else:
        emb = torch.cat(embs, 0)
    emb = emb.view(n_words, n_scales, -1)
    keep = ~plan.too_long
    labels = cluster_fn(emb[torch.from_numpy(keep).to(emb.device)], cfg)
    kept_words = [w for w, k in zip(plan.words, keep) if k]
    all_words = [w + [f"spk{int(l)}"] for w, l in zip(kept_words, labels)]
    return prepare_diarized_data_frame(all_words, segments_df, cfg.apply_deduplication)


def diarization_inference(out_dir: str, segments_df: pd.DataFrame, cfg: DiarizationCfg, fetch_from_cache: bool,
                          device: Optional[str] = None, pcm=None, sr: int = 16000) -> pd.DataFrame:
    """Same signature, modes, cache file and return value as the reference's diarization_inference
    (diarization.py:15-109).  ``pcm`` (optional, extension): the CSS streams as an int16 CUDA tensor [n_streams, n] in the
    order of the sorted wav file names, so that word_nmesc does not re-read the WAVs from disk."""
    assert segments_df.session_id.nunique() <= 1, 'no cross-session information is permitted'
    if cfg.method == "skip":
        out = segments_df.copy()
        out['speaker_id'] = 'spk0'
        return out
    elif cfg.method == "by_wav_file_name":
        out = segments_df.copy()
        ind, uniques = pd.factorize(out['wav_file_name'], sort=True)
        out['speaker_id'] = ind
        out['speaker_id'] = 'wav_' + out['speaker_id'].astype(str)
        return out

    session_name = segments_df.session_id[0]
    is_ct = session_name.startswith('close_talk')
    assert segments_df.wav_file_name.nunique() <= 3 or is_ct, 'expecting at most three separated channels'
    output_dir = Path(out_dir) / "diarization" / session_name / cfg.method
    out_file = output_dir / "all_segments_df.pkl"
    # diarization.py:82-89: no cache reads / writes when several ranks evaluate concurrently (they would race on the file)
    skip_cache_and_write = _world_size() > 1
    if not skip_cache_and_write:
        if fetch_from_cache and out_file.exists():
            return pd.read_pickle(out_file)
        os.makedirs(output_dir, exist_ok=True)

    segments_df = segments_df.copy()
    segments_df['wav_file_name'] = segments_df['wav_file_name'].astype('category')
    assert 'wav_file_name_ind' not in segments_df
    segments_df['wav_file_name_ind'] = segments_df['wav_file_name'].cat.codes
    wav_files = segments_df['wav_file_name'].cat.categories.to_list()

    if cfg.method == "word_nmesc":
        if pcm is None:
            # streams the CSS stage of this process left in HBM (same sample values as the WAV files it wrote), else the files
            from .css import device_streams_for
            hit = device_streams_for(wav_files)
            if hit is not None:
                pcm, sr = hit
        if pcm is None:
            from .css import flush_wav_writes
            flush_wav_writes(wav_files)
            pcm, sr = _load_streams_as_pcm(wav_files, device)
        out = word_based_clustering(pcm, sr, segments_df, cfg)
    else:
        from . import _cabi
        raise _cabi.NsfError(f"diarization method {cfg.method!r} is NeMo's time-based recipe (time_based_diarization.py): not built")
    if not skip_cache_and_write:
        out.to_pickle(out_file)
    return out


def _world_size() -> int:
    """utils/torch_utils.py get_world_size: 1 unless torch.distributed is initialised."""
    try:
        import torch.distributed as dist
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    except Exception:
        return 1


def _load_streams_as_pcm(wav_files: List[str], device):
    """File-boundary fallback of the hand-off: the 16-bit WAVs of the CSS stage -> int16 CUDA tensor, zero padded to the
    longest stream (word_based_diarization.py:156-164)."""
    import scipy.io.wavfile as wf
    import torch
    srs, data = zip(*[wf.read(str(f)) for f in wav_files])
    assert len(set(srs)) == 1, f'the separated streams disagree on the sample rate: {srs}'
    n = max(d.size for d in data)
    pcm = np.zeros((len(data), n), np.int16)
    for i, d in enumerate(data):
        pcm[i, :d.size] = d
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    return torch.from_numpy(pcm).to(dev), int(srs[0])                # word_based_diarization.py:156-157: sr comes from the files
