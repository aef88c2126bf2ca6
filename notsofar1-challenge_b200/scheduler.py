"""Sessions over the GPUs of a box: the reference processes sessions one after the other in a single process
(inference_pipeline/inference.py:59, "Process each session independently"); sessions are independent by rule, so a box
with N GPUs (one process per GPU, models resident) takes N at a time.

``assign_sessions`` balances by audio duration (longest-processing-time-first greedy: CSS cost is linear in segments);
``css_inference_distributed`` runs this rank's share through ``css_inference`` and returns the completed session rows of
ALL ranks, in the original order, on every rank (a gather of small Python objects; the WAVs are on the shared disk like in
the reference).  Very long sessions can instead be sharded by segments over all ranks: notsofar_b200.sharded.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence


def assign_sessions(durations: Sequence[float], world: int) -> List[List[int]]:
    """Indices of the sessions every rank processes: longest first onto the least loaded rank; ties keep input order."""
    loads = [0.0] * world
    shares: List[List[int]] = [[] for _ in range(world)]
    for i in sorted(range(len(durations)), key=lambda j: (-float(durations[j]), j)):
        r = min(range(world), key=lambda k: (loads[k], k))
        shares[r].append(i)
        loads[r] += float(durations[i])
    for s in shares:
        s.sort()
    return shares


def _wav_seconds(path: str) -> float:
    import wave
    with wave.open(str(path), "rb") as w:
        return w.getnframes() / float(w.getframerate())


def css_inference_distributed(out_dir: str, models_dir: str, sessions, cfg, fetch_from_cache: bool, group=None,
                              css_fn: Optional[Callable] = None, durations: Optional[Sequence[float]] = None) -> list:
    """sessions: list of session rows (pd.Series as load_meeting_data.py builds them).  Returns the list of rows with
    'sep_wav_file_names' added, identical on every rank."""
    import torch.distributed as dist
    if css_fn is None:
        from .css import css_inference as css_fn
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if durations is None:
        durations = [_wav_seconds(s.wav_file_names[0]) for s in sessions]
    mine = assign_sessions(durations, world)[rank]
    done = {i: css_fn(out_dir, models_dir, sessions[i], cfg, fetch_from_cache) for i in mine}
    if world == 1:
        return [done[i] for i in range(len(sessions))]
    gathered = [None] * world
    dist.all_gather_object(gathered, done, group=group)
    merged = {}
    for part in gathered:
        merged.update(part)
    return [merged[i] for i in range(len(sessions))]
