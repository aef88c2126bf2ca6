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


def _parse_cpulist(text: str) -> List[int]:
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11] (the kernel's cpulist format)."""
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_gpu(device_index: int, sysfs_root: str = "/sys/bus/pci/devices") -> dict:
    """Pin this process to the CPU cores of the NUMA node the GPU hangs off (its PCI device's ``local_cpulist``).

    One process per GPU moves ~1.2 GB per 30-min meeting through page-locked host buffers; page-locked memory is placed on
    the node of the allocating thread, so a rank that runs on the far socket pushes every byte over the inter-socket link,
    and 8 ranks then share that link instead of using 8 PCIe root ports.  Call before the first pinned allocation.
    Returns {'numa_node', 'cpus', 'bound'}; never raises (no sysfs entry, a cpuset that excludes the cores: left unbound)."""
    import os
    info = {"numa_node": None, "cpus": 0, "bound": False}
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index)
        bdf = f"{getattr(bus, 'pci_domain_id', 0):04x}:{bus.pci_bus_id:02x}:{bus.pci_device_id:02x}.0"
        base = os.path.join(sysfs_root, bdf)
        with open(os.path.join(base, "numa_node")) as f:
            info["numa_node"] = int(f.read().strip())
        with open(os.path.join(base, "local_cpulist")) as f:
            local = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus = sorted(set(local) & allowed)
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
        info["cpus"] = len(cpus)
    except Exception:
        pass
    return info


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
    if world > 1:
        import torch
        if torch.cuda.is_available():
            bind_host_to_gpu(torch.cuda.current_device())          # pinned staging buffers on the GPU's own NUMA node
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
