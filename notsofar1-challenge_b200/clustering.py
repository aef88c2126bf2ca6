"""NMESC speaker counting + spectral clustering of the word affinity matrix (the reference's
``run_clustering``, diarization/word_based_diarization.py:32-50).

The reference calls NeMo (``NMESC``, ``getAffinityGraphMat``, ``SpectralClustering`` of
nemo/collections/asr/parts/utils/offline_clustering.py -- third-party, unpinned, absent offline).  This module restates the
published algorithm (Park et al., "Auto-tuning spectral clustering for speaker diarization using normalized maximum
eigengap", 2020, as implemented upstream) on torch tensors; it runs where the affinity matrix lives (the GPU when it comes
from ``titanet.multiscale_affinity``).  The dense symmetric eigendecompositions are ``torch.linalg.eigh`` (cuSOLVER / LAPACK,
a plain library call: N is the number of words, at most ``nme_mat_size`` = 512 after NeMo's sub-sampling for the search).
**Parity unpinned**: no NeMo, no vectors in the reference; cluster labels are defined up to a permutation and NeMo's
k-means seeding is not reproduced (k-means++ with a fixed seed, best of ``n_trials`` by inertia).
"""
from __future__ import annotations

from typing import Tuple

import torch


def kneighbors_connections(aff: torch.Tensor, p: int) -> torch.Tensor:
    """getKneighborsConnections: 1 where j is among the p largest entries of row i (stored transposed, as upstream)."""
    n = aff.shape[0]
    idx = torch.argsort(aff, dim=1, descending=True, stable=True)[:, :p]
    out = torch.zeros_like(aff)
    out[idx.T, torch.arange(n, device=aff.device)] = 1.0
    return out


def affinity_graph(aff: torch.Tensor, p: int) -> torch.Tensor:
    """getAffinityGraphMat: symmetrised p-nearest-neighbour graph (p <= 0: the raw affinity)."""
    x = aff if p <= 0 else kneighbors_connections(aff, p)
    return 0.5 * (x + x.T)


def laplacian(x: torch.Tensor) -> torch.Tensor:
    """getLaplacian: unnormalised graph Laplacian of the off-diagonal weights."""
    x = x.clone()
    x.fill_diagonal_(0)
    return torch.diag(x.abs().sum(1)) - x


def estimate_num_speakers(graph: torch.Tensor, max_num_speakers: int) -> Tuple[int, torch.Tensor, torch.Tensor]:
    """estimateNumofSpeakers: position of the largest gap among the smallest Laplacian eigenvalues."""
    lambdas = torch.linalg.eigvalsh(laplacian(graph).double()).sort()[0]
    gaps = lambdas[1:] - lambdas[:-1]
    n = int(torch.argmax(gaps[:min(max_num_speakers, gaps.shape[0])]).item()) + 1
    return n, lambdas, gaps


def is_fully_connected(graph: torch.Tensor) -> bool:
    n = graph.shape[0]
    reach = torch.zeros(n, dtype=torch.bool, device=graph.device)
    reach[0] = True
    adj = graph > 0
    for _ in range(n):
        new = reach | (adj[reach].any(0))
        if bool((new == reach).all()):
            break
        reach = new
    return bool(reach.all())


def nmesc(aff: torch.Tensor, max_num_speakers: int = 8, max_rp_threshold: float = 0.06, sparse_search_volume: int = 30,
          nme_mat_size: int = 512) -> Tuple[int, int]:
    """NMESC.forward: -> (estimated number of speakers, p_hat).  For every candidate neighbour count p the graph is binarised,
    g_p = (p / N) / (largest normalised eigengap) is evaluated, and the p with the smallest g_p wins."""
    n_full = aff.shape[0]
    ratio = max(1, n_full // nme_mat_size) if n_full > nme_mat_size else 1            # subsampleAffinityMat
    mat = aff[::ratio, ::ratio].contiguous()
    n = mat.shape[0]
    max_n = max(int(n * max_rp_threshold), 2)
    steps = min(max_n, sparse_search_volume)
    p_list = sorted(set(int(v) for v in torch.linspace(1, max_n, steps).to(torch.int).tolist()))
    best = None
    est = {}
    for p in p_list:
        k, lambdas, gaps = estimate_num_speakers(affinity_graph(mat, p), max_num_speakers)
        max_gap = gaps[:max_num_speakers].max() / (lambdas.max() + 1e-10)
        g_p = (p / n) / (float(max_gap) + 1e-10)
        est[p] = k
        if best is None or g_p < best[0]:
            best = (g_p, p)
    p_hat = best[1]
    if not is_fully_connected(affinity_graph(mat, p_hat)):
        # getMinimumConnection [upstream]: walk the candidate list itself (not every integer) until the graph is connected or p
        # exceeds max_N; the speaker count is the one already estimated for that candidate (est_spk_n_dict)
        for p in p_list:
            p_hat = p
            if is_fully_connected(affinity_graph(mat, p)) or p > max_n:
                break
    k = est[p_hat]
    return k, ratio * p_hat


def _kmeans(x: torch.Tensor, k: int, n_trials: int = 10, iters: int = 100, seed: int = 0) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    n = x.shape[0]
    best_labels, best_inertia = None, None
    for _ in range(n_trials):
        centers = x[int(torch.randint(n, (1,), generator=g))][None].clone()
        for _c in range(1, k):                                                         # k-means++ seeding
            d2 = torch.cdist(x, centers).min(1)[0] ** 2
            pr = (d2 / d2.sum().clamp_min(1e-30)).cpu()
            centers = torch.cat([centers, x[int(torch.multinomial(pr, 1, generator=g))][None]], 0)
        for _i in range(iters):
            labels = torch.cdist(x, centers).argmin(1)
            new = torch.stack([x[labels == c].mean(0) if bool((labels == c).any()) else centers[c] for c in range(k)])
            if torch.allclose(new, centers):
                break
            centers = new
        inertia = float((torch.cdist(x, centers).min(1)[0] ** 2).sum())
        if best_inertia is None or inertia < best_inertia:
            best_labels, best_inertia = labels, inertia
    return best_labels


def spectral_clustering(graph: torch.Tensor, n_clusters: int) -> torch.Tensor:
    """SpectralClustering.forward: k-means on the eigenvectors of the n_clusters smallest Laplacian eigenvalues."""
    if n_clusters <= 1:
        return torch.zeros(graph.shape[0], dtype=torch.long, device=graph.device)
    _, vec = torch.linalg.eigh(laplacian(graph).double())
    return _kmeans(vec[:, :n_clusters].float(), n_clusters)


def run_clustering(raw_affinity: torch.Tensor, max_num_speakers: int = 8, max_rp_threshold: float = 0.06,
                   sparse_search_volume: int = 30):
    """word_based_diarization.py:32-50: NMESC -> p-neighbour graph of the full matrix -> spectral clustering; int labels [n]."""
    aff = raw_affinity.float()
    if aff.shape[0] == 1:
        return torch.zeros(1, dtype=torch.long).numpy()
    k, p_hat = nmesc(aff, max_num_speakers, max_rp_threshold, sparse_search_volume)
    labels = spectral_clustering(affinity_graph(aff, p_hat), k)
    return labels.cpu().numpy()


def nmesc_backend(emb: torch.Tensor, cfg=None):
    """Clustering backend for ``diarization.set_clustering_backend``: emb [n_words, n_scales, D] -> labels, with the affinity
    of word_based_diarization.py:171-177 computed by the CUDA kernels of csrc/titanet.cu."""
    from .titanet import multiscale_affinity
    # word_based_diarization.py:171: the embeddings are rounded to fp16 before the affinity (which NeMo then evaluates in fp32)
    return run_clustering(multiscale_affinity(emb.half().float()))
