"""CPU-side checks: the C-ABI library loads and exports every symbol include/nsf_b200.h declares (no compute
calls without a GPU), and the host logic of notsofar_b200 (segment plan, permutation chain, segment weights,
weight packing, reference-compatible config) matches the oracle / the reference's conventions."""
import dataclasses
import os
import re

import numpy as np
import pytest

from oracle import css_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def N():
    import notsofar_b200
    return notsofar_b200


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "nsf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nsf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(N):
    lib = N._cabi.load()
    syms = _header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/nsf_b200.h but not exported"
    assert set(syms) == set(N._cabi.SIGNATURES), "ctypes signature table out of sync with the header"
    assert b"sm_100a" in lib.nsf_version()
    assert lib.nsf_num_frames(48000) == 186 and lib.nsf_num_frames(511) == 0 and lib.nsf_num_frames(512) == 1


def test_library_argument_errors_without_gpu(N):
    lib = N._cabi.load()
    rc = lib.nsf_stft_mc(None, 1000, 7, None, 10, 2, None)
    assert rc == -1 and b"null" in lib.nsf_last_error()
    dims = N._cabi.ConformerDims(d_model=100, n_heads=8, d_ff=1024, n_blocks=18, kernel_size=33, in_features=1799,
                                 n_out=1028, maxlen=1000, T=186, gemm_engine=1)
    assert lib.nsf_conformer_num_offsets(dims) == 12 + 34 * 18
    # the entry points added for the diarization / ASR rows validate their arguments before touching the device
    import ctypes as C
    from notsofar_b200.titanet import TitanetDims
    from notsofar_b200.whisper import _RulesStruct
    td = TitanetDims()
    td.feat_in, td.n_blocks, td.att_ch, td.emb = 80, 2, 128, 192
    td.filters[0], td.repeat[0], td.kernel[0], td.residual[0] = 1024, 1, 3, 0
    td.filters[1], td.repeat[1], td.kernel[1], td.residual[1] = 1024, 3, 7, 1
    assert lib.nsf_titanet_num_offsets(C.byref(td)) == (4 * 1 + 2) + (4 * 3 + 2 + 3) + 11
    assert lib.nsf_titanet_workspace_bytes(C.byref(td), 8, 64) > 8 * 64 * 1024 * 4
    td.precision = 1                                                       # fp16 operands: same blob entries
    assert lib.nsf_titanet_num_offsets(C.byref(td)) == (4 * 1 + 2) + (4 * 3 + 2 + 3) + 11
    td.precision = 2
    assert lib.nsf_titanet_num_offsets(C.byref(td)) == 0
    td.precision = 0
    td.kernel[1] = 8                                                       # even kernel: unsupported
    assert lib.nsf_titanet_num_offsets(C.byref(td)) == 0 and lib.nsf_titanet_workspace_bytes(C.byref(td), 8, 64) == 0
    assert lib.nsf_titanet_features(None, None, 1, 100, 16, None, 80, None, None, None, None, None) == -1
    assert b"null" in lib.nsf_last_error()
    assert lib.nsf_cos_affinity_accum(None, 192, 192, 5, 1.0, None, None, None, None, None) == -1
    assert lib.nsf_whisper_alignment_workspace_bytes(2, 3, 10, 1500) >= 2 * 11 * 1501 * 5 and lib.nsf_whisper_alignment_workspace_bytes(0, 3, 10, 1500) == 0
    assert lib.nsf_whisper_alignment(None, 1, 1, 1, 1, 1, None, None, None, None, 0, None) == -1
    r = _RulesStruct(0, 450, 449, 400, -1, 0, 0)                            # sample_begin 0: invalid
    assert lib.nsf_whisper_logit_rules(None, 1, 600, None, 4, None, C.byref(r), None, None, None) == -1


def test_sass_has_blackwell_tensor_core_and_tma_instructions(N):
    """The mask-network GEMM must be a tcgen05 / TMA kernel, not a recompiled mma.sync one."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", N._cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
    assert "HMMA." not in sass.replace("UTCHMMA", "")


def test_cfg_is_field_compatible_with_reference(N):
    fields = [f.name for f in dataclasses.fields(N.CssCfg)]
    expected = ["segment_size_sec", "hop_size_sec", "normalize_segment_power", "stitching_loss", "stitching_input",
                "seg_weight_m0_sec", "seg_weight_m1_sec", "activity_th", "activity_dilation_sec", "activity_erosion_sec",
                "device", "show_progressbar", "checkpoint_sc", "checkpoint_mc", "device_id", "num_spks", "mc_mvdr",
                "mc_mask_floor_db", "sc_mask_floor_db", "pass_through_ch0", "slice_audio_for_debug"]
    assert fields == expected                       # css/css.py:24-48
    c = N.CssCfg()
    assert (c.segment_size_sec, c.hop_size_sec, c.activity_th, c.num_spks, c.mc_mvdr, c.mc_mask_floor_db) == (3., 1.5, 0.4, 3, True, 0.)
    from oracle import reference_shim as R
    if R.available():
        ref = R.load().css.CssCfg
        assert [f.name for f in dataclasses.fields(ref)] == fields
        assert dataclasses.asdict(ref()) == dataclasses.asdict(c)


@pytest.mark.parametrize("n", [512, 20000, 48127, 48128, 60000, 160000, 28_800_000])
def test_segment_plan_matches_oracle(N, n):
    a = N.plan_segments(n, 16000, N.CssCfg())
    b = O.plan_segments(n, 16000, O.OracleCfg())
    assert dataclasses.asdict(a) == {k: getattr(b, k) for k in dataclasses.asdict(a)}


def test_segment_weights_and_sum(N):
    from notsofar_b200 import css as ncss
    plan = N.plan_segments(160000, 16000, N.CssCfg())
    seg_w, wsum = ncss._segment_weights(plan)
    assert seg_w.shape == (plan.num_segments, 186) and wsum.shape == (plan.mix_frames,)
    assert np.abs(seg_w[0] - O.calc_segment_weight(186, 9, 18, is_first=True)).max() < 1e-7
    assert np.abs(seg_w[1] - O.calc_segment_weight(186, 9, 18)).max() < 1e-7
    assert np.abs(seg_w[-1] - O.calc_segment_weight(186, 9, 18, is_last=True)).max() < 1e-7
    assert (wsum > 1e-5).all()
    # a single 3-s segment leaves zero weight at the right edge: the reference asserts (css.py:297), so do we
    short = N.plan_segments(48127, 16000, N.CssCfg())
    assert short.num_segments == 1
    _, ws = ncss._segment_weights(short)
    assert not (ws > 1e-5).all()


def test_permutation_chain_equals_sequential_hungarian(N):
    """costs on the ORIGINAL channel order + host chain == the reference's in-place sequential alignment."""
    rng = np.random.default_rng(0)
    n_seg, S, F_, T, ov = 12, 3, 17, 20, 10
    masks = rng.random((n_seg, S, F_, T)).astype(np.float32)
    # make neighbours similar up to a random permutation so that the assignment is well separated
    true = [rng.permutation(S) for _ in range(n_seg)]
    base = rng.random((S, F_, T + (n_seg - 1) * (T - ov))).astype(np.float32)
    for i in range(n_seg):
        st = i * (T - ov)
        masks[i] = base[true[i], :, st:st + T] + 0.01 * masks[i]
    costs = np.zeros((n_seg, S, S), np.float32)
    for i in range(1, n_seg):
        costs[i] = O.pit_cost_l1(masks[i - 1][:, :, T - ov:].transpose(1, 2, 0), masks[i][:, :, :ov].transpose(1, 2, 0))
    perms = N.permutation_chain(costs)
    seq = masks.copy()
    ref = [np.arange(S)]
    for i in range(1, n_seg):
        c = O.pit_cost_l1(seq[i - 1][:, :, T - ov:].transpose(1, 2, 0), seq[i][:, :, :ov].transpose(1, 2, 0))
        p = O.assign(c)
        seq[i] = seq[i][p]
        ref.append(p)
    assert np.array_equal(perms, np.stack(ref))
    assert not np.array_equal(perms, np.tile(np.arange(S), (n_seg, 1)))


def test_permutation_chain_continues_chunk_by_chunk(N):
    """The progressive tail walks the chain one chunk of segments at a time: (perms, state) of a prefix + the rest with
    prev_state == the one-shot walk, wherever the cut is (also inside / at the ends of the blocked prefix composition)."""
    rng = np.random.default_rng(3)
    for n_seg, S in ((1, 3), (2, 3), (7, 3), (37, 2), (400, 3), (101, 4)):
        costs = rng.random((n_seg, S, S)).astype(np.float32)
        costs[0] = 0
        full = N.permutation_chain(costs)
        for cuts in ([1], [n_seg // 2], [n_seg - 1], [n_seg // 3, 2 * n_seg // 3]):
            cuts = sorted({c for c in cuts if 0 < c < n_seg})
            if not cuts:
                continue
            parts, state, lo = [], None, 0
            for hi in cuts + [n_seg]:
                p, state = N.permutation_chain(costs[lo:hi], prev_state=state, return_state=True)
                parts.append(p)
                lo = hi
            assert np.array_equal(np.concatenate(parts), full), (n_seg, S, cuts)


def test_permutation_chain_is_equivariant_under_relabelling(N):
    """What the sharded progressive read-back rests on: a chain that starts from the identity at segment s is the global
    chain up to one relabelling of the output slots, tau = perms[s] -- perms[i][k] == local[i - s][tau[k]] for every
    i >= s -- because the best order after a previous order q is sigma o q (sums of three float32 costs are exact in
    float64, so the arg-min does not depend on the labelling).  An exact tie between two assignments breaks it; the
    product detects that by this very comparison and falls back."""
    rng = np.random.default_rng(11)
    for n_seg, S in ((40, 3), (25, 2), (30, 4)):
        costs = rng.random((n_seg, S, S)).astype(np.float32)
        costs[0] = 0
        full = N.permutation_chain(costs)
        assert len({tuple(p) for p in full}) > 1
        for s in (1, 7, n_seg - 2):
            loc = costs[s:].copy()
            loc[0] = 0                                  # a rank's first local segment has no predecessor
            local = N.permutation_chain(loc)
            tau = full[s]
            assert np.array_equal(local[0], np.arange(S))
            assert np.array_equal(full[s:], local[:, tau])
    # a tie: two assignments with identical totals -> the first in enumeration order wins, whatever the labels are
    costs = np.zeros((3, 3, 3), np.float32)
    costs[1] = np.array([[1, 0, 2], [0, 1, 2], [2, 2, 0]], np.float32)      # unambiguous: order (1, 0, 2) costs 0
    costs[2] = 1.0                                                         # every order costs 3: a six-way tie
    full = N.permutation_chain(costs)
    loc = costs[1:].copy()
    loc[0] = 0
    local = N.permutation_chain(loc)
    tau = full[1]
    assert np.array_equal(full[1], [1, 0, 2])
    assert not np.array_equal(full[1:], local[:, tau])                      # the tie at the last segment is what the check catches


def test_plan_batches_progressive():
    """Chunks of the streaming path: short first chunk, then full (wave-filling) chunks, a short remainder merged into
    the last one; nothing changes for resident recordings, short sessions or with the progressive tail switched off."""
    from notsofar_b200.css import plan_batches
    assert plan_batches(1209, 1280, streaming=True, progressive=356) == [(0, 176), (176, 356), (532, 356), (888, 321)]
    assert plan_batches(1209, 1280, streaming=True) == [(0, 176), (176, 1033)]
    assert plan_batches(1209, 1280, streaming=False, progressive=356) == [(0, 1209)]
    assert plan_batches(241, 1280, streaming=True, progressive=356) == [(0, 241)]
    assert plan_batches(176 + 356 + 80, 1280, streaming=True, progressive=356) == [(0, 176), (176, 436)]
    for n in (352, 353, 800, 5000, 9677):
        chunks = plan_batches(n, 1280, streaming=True, progressive=356)
        assert chunks[0] == (0, 176) and sum(c for _, c in chunks) == n
        assert all(a[0] + a[1] == b[0] for a, b in zip(chunks, chunks[1:]))
        assert all(c == 356 for _, c in chunks[1:-1]) and 0 < chunks[-1][1] <= 356 + 356 // 4


def test_pack_weights_layout(N, small_weights):
    dims, blob, offsets, extra = N.pack_weights(small_weights, T=186, gemm_engine=N.GEMM_TC_3XTF32)
    assert (dims.d_model, dims.n_heads, dims.d_ff, dims.n_blocks, dims.kernel_size) == (128, 2, 256, 2, 33)
    assert (dims.in_features, dims.n_out, dims.maxlen) == (1799, 1028, 1000)
    assert len(offsets) == 12 + 34 * 2 and np.all(offsets % 64 == 0) and blob.dtype == np.float32
    Kf = 1824
    hi = blob[offsets[0]:offsets[0] + 128 * Kf].reshape(128, Kf)
    lo = blob[offsets[1]:offsets[1] + 128 * Kf].reshape(128, Kf)
    w = small_weights["executor.nnet.conformer.embed.0.weight"]
    assert np.array_equal(hi[:, :1799] + lo[:, :1799], w) and np.all(hi[:, 1799:] == 0) and np.all(lo[:, 1799:] == 0)
    assert np.all(hi.view(np.uint32) & 0x1FFF == 0)
    assert np.abs(lo).max() <= np.abs(w).max() * 2.0 ** -10
    # the DDP "module." prefix of real checkpoints (css/helpers.py:30-36) is accepted
    pref = {"module." + k: v for k, v in small_weights.items()}
    _, blob2, off2, _ = N.pack_weights(pref, T=186, gemm_engine=N.GEMM_TC_3XTF32)
    assert np.array_equal(blob, blob2) and np.array_equal(offsets, off2)


def test_pack_weights_folded_layernorms(N, small_weights, monkeypatch):
    """2xBF16 engine, d_model = 128: the attention and feed_forward_out LayerNorms are folded into the GEMM that follows.
    LN(x) W^T + b == rstd (x (gamma W)^T - mean colsum) + (b + W beta) with the weights / sums / biases exactly as packed."""
    import torch
    import ctypes as C
    lib = N._cabi.load()
    dims, blob, offsets, _ = N.pack_weights(small_weights, T=186, gemm_engine=N.GEMM_TC_2XBF16)
    assert lib.nsf_conformer_ln_fold(C.byref(dims)) == 1 and len(offsets) == 12 + 34 * 2
    d3 = N._cabi.ConformerDims(dims.d_model, dims.n_heads, dims.d_ff, dims.n_blocks, dims.kernel_size, dims.in_features, dims.n_out,
                               dims.maxlen, dims.T, N.GEMM_TC_3XTF32)
    assert lib.nsf_conformer_ln_fold(C.byref(d3)) == 0                      # other engines keep the LayerNorm kernels
    d = 128

    def bf16_pairs(off, rows, cols, src=None):
        raw = (blob if src is None else src)[off:off + rows * cols // 2].view(np.int16).copy()
        return torch.from_numpy(raw).view(torch.bfloat16).to(torch.float64).numpy().reshape(rows, cols)

    rng = np.random.default_rng(5)
    x = (rng.standard_normal((7, d)) * 1.7 + 0.6).astype(np.float64)            # rows with a mean
    mu, var = x.mean(1, keepdims=True), x.var(1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + 1e-5)
    P = "executor.nnet.conformer.encoders.1."
    base = 12 + 34
    for name, o_w, o_b, o_cs, n_out, W, b, g, beta in (
            ("qkv", 10, 12, 32, 3 * d,
             np.concatenate([small_weights[P + f"self_attn.linear_{c}.weight"] for c in "qkv"], 0),
             np.concatenate([small_weights[P + f"self_attn.linear_{c}.bias"] for c in "qkv"]),
             small_weights[P + "self_attn.layer_norm.weight"], small_weights[P + "self_attn.layer_norm.bias"]),
            ("ffo", 24, 26, 33, 256,
             small_weights[P + "feed_forward_out.net.0.weight"], small_weights[P + "feed_forward_out.net.0.bias"],
             small_weights[P + "feed_forward_out.layer_norm.weight"], small_weights[P + "feed_forward_out.layer_norm.bias"])):
        Wst = bf16_pairs(offsets[base + o_w], n_out, d) + bf16_pairs(offsets[base + o_w + 1], n_out, d)
        bias = blob[offsets[base + o_b]:offsets[base + o_b] + n_out].astype(np.float64)
        cs = blob[offsets[base + o_cs]:offsets[base + o_cs] + n_out].astype(np.float64)
        ref = ((x - mu) * rstd * g.astype(np.float64) + beta.astype(np.float64)) @ W.astype(np.float64).T + b
        got = rstd * (x @ Wst.T - mu * cs[None, :]) + bias[None, :]
        err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        assert err < 2e-5, (name, err)                                         # bf16 head + remainder: 2^-17 per weight
    # NSF_LN_FOLD=0: the plain layout (no scaling, zero column sums)
    monkeypatch.setenv("NSF_LN_FOLD", "0")
    dims0, blob0, offsets0, _ = N.pack_weights(small_weights, T=186, gemm_engine=N.GEMM_TC_2XBF16)
    assert lib.nsf_conformer_ln_fold(C.byref(dims0)) == 0
    assert np.all(blob0[offsets0[base + 32]:offsets0[base + 32] + 3 * d] == 0)
    W0 = bf16_pairs(offsets0[base + 10], 3 * d, d, blob0) + bf16_pairs(offsets0[base + 11], 3 * d, d, blob0)
    Wq = np.concatenate([small_weights[P + f"self_attn.linear_{c}.weight"] for c in "qkv"], 0)
    assert np.abs(W0 - Wq).max() <= np.abs(Wq).max() * 2.0 ** -15


def test_product_does_not_import_oracle():
    """The product path must not route through the oracle (or the reference)."""
    pkg = os.path.join(ROOT, "notsofar1-challenge_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("# oracle", "") or fn == "__init__.py" and False, fn
            assert "/root/reference" not in src, fn
