"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Build-container only (needs /root/reference): imports the unmodified reference through
oracle/reference_shim.py, runs css.css.separate_and_stitch on a 4.6 s excerpt of the
reference's bundled 7-channel sample (sample_data/css_train_set/*.mixture) with a small
seeded mask network of the reference architecture, and records the inputs and outputs of
every stage (by wrapping, not editing, separator.separate and make_mvdr).

    python tests/golden/make_golden.py

Outputs (committed):
    css_golden_small.npz   -- stage-wise I/O of the reference run (float32 / complex64 as the
                              reference produced them; float16-free; int outputs exact)
The mask-network weights are NOT stored: they are oracle.css_oracle.random_weights(seed=1,
d_model=128, n_heads=2, d_ff=256, n_blocks=2), a numpy-PCG64 stream that regenerates
identically on the GPU box.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import css_oracle as O            # noqa: E402
from oracle import reference_shim as R        # noqa: E402

SMALL_NET = dict(seed=1, d_model=128, n_heads=2, d_ff=256, n_blocks=2)
# Short segments keep the fixture small (the reference takes the segment length from its cfg,
# css.py:142-147): 1 s segments -> 61 frames, hop 30, overlap 31; 36 200 samples -> 140 frames ->
# 4 segments, the last one zero-padded.  The production shape (3 s -> 186 frames) is pinned
# against the live reference by tests/test_oracle_pinned.py in the build container.
SEGMENT_SEC, HOP_SEC = 1.0, 0.5
N_SAMPLES = 36200
OFFSET = 30000
ACTIVITY_TH = 0.0          # replaced below by the median of the activity so the gate is non-trivial
# The mask network has a fixed output order, so adjacent segments never need re-ordering.  To
# exercise the permutation chain the separator plug-in used for the golden run shuffles its
# speaker channels per call (separators are plug-ins: README.md:229-232).
CHANNEL_SHUFFLES = [(0, 1, 2), (2, 0, 1), (1, 0, 2), (2, 1, 0)]


def main():
    import torch
    ns = R.load()
    x, raw, scale = R.sample_mixture(N_SAMPLES, OFFSET)
    w = O.random_weights(**SMALL_NET)
    sep = R.build_separator(w)
    def make_cfg(th):
        return ns.css.CssCfg(show_progressbar=False, activity_th=th, segment_size_sec=SEGMENT_SEC,
                             hop_size_sec=HOP_SEC)

    rec = {"sep_in": [], "masks": [], "mvdr": [], "mvdr64": [], "pit": []}
    orig_separate = sep.separate

    def separate_rec(stft):
        out = orig_separate(stft)
        if stft.shape[-1] == 7 and len(rec["masks"]) < len(CHANNEL_SHUFFLES):
            sh = list(CHANNEL_SHUFFLES[len(rec["masks"])])
            out = {"spk_masks": out["spk_masks"][..., sh].contiguous(), "noise_masks": out["noise_masks"]}
        rec["sep_in"].append(stft.detach().cpu().numpy()[0].copy())                     # [F, T, C]
        m = torch.cat([out["spk_masks"], out["noise_masks"]], -1)[0]
        rec["masks"].append(m.detach().cpu().numpy().transpose(2, 0, 1).copy())         # [4, F, T]
        return out

    sep.separate = separate_rec
    orig_mvdr = ns.css.make_mvdr

    def mvdr_rec(spk, noise, mix_wav=None, mix_stft=None, return_stft=False):
        res = orig_mvdr(spk, noise, mix_wav=mix_wav, mix_stft=mix_stft, return_stft=return_stft)
        rec["mvdr"].append(np.stack(res).copy())
        res64 = orig_mvdr(spk.astype(np.float64), noise.astype(np.float64),
                          mix_stft=mix_stft.astype(np.complex128), return_stft=True)     # fp64-lifted reference
        rec["mvdr64"].append(np.stack(res64).copy())
        return res

    ns.css.make_mvdr = mvdr_rec
    orig_pit_forward = ns.losses.PitWrapper.forward

    def pit_rec(self, preds, targets):
        loss, perms = orig_pit_forward(self, preds, targets)
        rec["pit"].append(np.asarray(perms[0]).copy())
        return loss, perms

    ns.losses.PitWrapper.forward = pit_rec
    try:
        # pass 1: find a threshold inside the activity spread; pass 2: the recorded run
        _, side = ns.css.separate_and_stitch(x[None], sep, 16000, torch.device("cpu"), make_cfg(0.3))
        act = side["mask_stitched"].mean(dim=1)[0].numpy()
        th = float(np.round(np.median(act), 3))
        gap = np.abs(act - th).min()
        print("activity quantiles", np.quantile(act, [0, .25, .5, .75, 1]), "th", th, "closest frame", gap)
        assert gap > 1e-5, "threshold is on a knife edge; pick another"
        for k in rec:
            rec[k].clear()
        cfg = make_cfg(th)
        wavs, side = ns.css.separate_and_stitch(x[None], sep, 16000, torch.device("cpu"), cfg)
    finally:
        ns.css.make_mvdr = orig_mvdr
        ns.losses.PitWrapper.forward = orig_pit_forward
        sep.separate = orig_separate

    with torch.no_grad():
        stft_ref = sep.stft(torch.from_numpy(x[None])).numpy()[0]                        # [F, T, C]
        # feature extractor on segment 0 exactly as separate() calls it
        st = torch.from_numpy(rec["sep_in"][0][None]).moveaxis(3, 1).contiguous()
        _, _, feat = sep.executor.extractor(mix=None, mag=st.abs(), pha=st.angle())
        feat0 = feat[0].numpy().T.copy()                                                 # [T, 1799]
        # iSTFT stage on the stitched, gated STFT is implied by wavs; also keep a direct pair
        rng = np.random.default_rng(7)
        s_in = (rng.standard_normal((2, 257, 40)) + 1j * rng.standard_normal((2, 257, 40))).astype(np.complex64)
        s_out = sep.istft(torch.from_numpy(s_in)).numpy()

    seg_w = np.stack([ns.css.calc_segment_weight(186, 9, 18).numpy(),
                      ns.css.calc_segment_weight(186, 9, 18, is_first_seg=True).numpy(),
                      ns.css.calc_segment_weight(186, 9, 18, is_last_seg=True).numpy()])
    # the reference's own known-answer vectors on this path (numpy_utils.py:16-22)
    morph_in = np.array([1, 1, 0, 1, 1, 1, 0, 0, 0, 1, 1, 0, 0], dtype=bool)
    morph_er = ns.numpy_utils.erode(morph_in, 1)
    morph_di = ns.numpy_utils.dilate(morph_in, 1)

    out = dict(
        mixture_int16=raw, mixture_scale=np.float64(scale), activity_th=np.float64(th),
        segment_size_sec=np.float64(SEGMENT_SEC), hop_size_sec=np.float64(HOP_SEC),
        channel_shuffles=np.array(CHANNEL_SHUFFLES, dtype=np.int64),
        stft=stft_ref.astype(np.complex64),                               # [F, T_long, C]; segments are slices of it
        feat0=feat0.astype(np.float32),                                   # features of segment 0 [T, 1799]
        masks=np.stack(rec["masks"]).astype(np.float32),                  # [4, 4, F, T] (speaker channels shuffled)
        mvdr=np.stack(rec["mvdr"])[1:3].astype(np.complex64),             # segments 1,2  reference fp32
        mvdr64=np.stack(rec["mvdr64"])[1:3].astype(np.complex64),         # segments 1,2  reference lifted to fp64 (stored c64)
        perms=np.stack(rec["pit"]).astype(np.int64),                      # [3, 3]
        morph_in=morph_in, morph_erode=morph_er, morph_dilate=morph_di,
        mask_stitched=side["mask_stitched"].numpy().astype(np.float32),   # [1, F, T_long, 3]
        activity_b=side["activity_b"].numpy(),
        activity_final=side["activity_final"].numpy(),
        segment_frames=np.int64(side["segment_frames"]),
        wavs=np.stack(wavs).astype(np.float32),                           # [3, N']
        istft_in=s_in, istft_out=s_out.astype(np.float32),
        seg_weights=seg_w.astype(np.float32),
    )
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "css_golden_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")
    print("perms", out["perms"].tolist(), "activity_b frac", out["activity_b"].mean(),
          "activity_final frac", out["activity_final"].mean())


if __name__ == "__main__":
    main()
