"""ctypes binding of libnsf_b200.so (include/nsf_b200.h).  torch owns every buffer; this module only
passes raw device pointers, sizes and the current CUDA stream, and turns error codes into exceptions.
There is no CPU fallback: if the library is missing or a call fails, it raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnsf_b200.so")

GEMM_SIMT_FP32, GEMM_TC_3XTF32, GEMM_TC_TF32, GEMM_TC_2XBF16, GEMM_TC_2XF16, GEMM_TC_BF16 = 0, 1, 2, 3, 4, 5
SPLIT_TF32, SPLIT_BF16, SPLIT_F16, SPLIT_BF16_1, SPLIT_FP32 = 0, 1, 2, 3, 4
NSF_OK, NSF_ERR_INVALID_ARG, NSF_ERR_CUDA, NSF_ERR_UNSUPPORTED = 0, -1, -2, -3      # include/nsf_b200.h
F16_ACT_SCALE, F16_WEIGHT_SCALE = 16.0, 256.0        # csrc/common.cuh kF16ActScale / kF16WeightScale


def split_fmt_of_engine(engine: int) -> int:
    return {GEMM_TC_2XBF16: SPLIT_BF16, GEMM_TC_2XF16: SPLIT_F16, GEMM_TC_BF16: SPLIT_BF16_1}.get(engine, SPLIT_TF32)

c_f32p = C.c_void_p
i64 = C.c_int64
i32 = C.c_int


class ConformerDims(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("d_model", "n_heads", "d_ff", "n_blocks", "kernel_size", "in_features",
                                        "n_out", "maxlen", "T", "gemm_engine")]


# name -> (restype, argtypes); must list every symbol include/nsf_b200.h declares
SIGNATURES = {
    "nsf_last_error": (C.c_char_p, []),
    "nsf_version": (C.c_char_p, []),
    "nsf_launch_count": (i64, []),
    "nsf_prof_enable": (i32, [i32]),
    "nsf_prof_num_classes": (i32, []),
    "nsf_prof_class_name": (C.c_char_p, [i32]),
    "nsf_prof_collect": (i32, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i64), i32]),
    "nsf_num_frames": (i64, [i64]),
    "nsf_stft_mc": (i32, [c_f32p, i64, i32, c_f32p, i64, i64, C.c_void_p]),
    "nsf_css_features": (i32, [c_f32p, i64, i64, i32, i64, i32, i32, i32, c_f32p, c_f32p, c_f32p, c_f32p, i64, i32, C.c_void_p]),
    "nsf_conformer_create": (i32, [C.POINTER(ConformerDims), c_f32p, i64, C.POINTER(i64), i32, C.POINTER(C.c_void_p)]),
    "nsf_conformer_destroy": (None, [C.c_void_p]),
    "nsf_conformer_num_offsets": (i64, [C.POINTER(ConformerDims)]),
    "nsf_conformer_ln_fold": (i32, [C.POINTER(ConformerDims)]),
    "nsf_conformer_workspace_bytes": (i64, [C.POINTER(ConformerDims), i32]),
    "nsf_conformer_forward": (i32, [C.c_void_p, c_f32p, c_f32p, i64, i32, c_f32p, C.c_void_p, i64, C.c_void_p]),
    "nsf_mvdr": (i32, [c_f32p, i32, i32, c_f32p, i64, i64, i32, i64, i32, i32, i32, i32, C.c_float, c_f32p, C.c_void_p]),
    "nsf_mvdr_utterance_workspace_bytes": (i64, [i32, i64, i32]),
    "nsf_mvdr_utterance": (i32, [c_f32p, i32, i32, c_f32p, i64, i32, i32, C.c_float, c_f32p, C.c_void_p, i64, C.c_void_p]),
    "nsf_pit_cost": (i32, [C.c_void_p, i32, i32, i32, i32, i32, i32, i32, i32, c_f32p, C.c_void_p]),
    "nsf_pit_cost_range": (i32, [C.c_void_p, i32, i32, i32, i32, i32, i32, i32, i32, i32, c_f32p, C.c_void_p]),
    "nsf_stitch_masks": (i32, [c_f32p, i32, C.c_void_p, c_f32p, c_f32p, i32, i32, i32, i32, i32, i64, c_f32p, c_f32p, C.c_void_p]),
    "nsf_activity": (i32, [c_f32p, i64, i32, C.c_float, i32, i32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsf_stitch_stft": (i32, [c_f32p, C.c_void_p, c_f32p, c_f32p, C.c_void_p, i32, i32, i32, i32, i32, i64, c_f32p, C.c_void_p]),
    "nsf_istft": (i32, [c_f32p, i32, i64, c_f32p, C.c_void_p]),
    "nsf_istft_range": (i32, [c_f32p, i32, i64, c_f32p, i64, i64, C.c_void_p]),
    "nsf_stitch_progress": (i32, [c_f32p, i32, c_f32p, C.c_void_p, c_f32p, c_f32p, i32, i32, i32, i32, i32, i32, i32, i64, C.c_float, i32, i32,
                                  c_f32p, c_f32p, C.c_void_p, C.c_void_p, C.c_void_p, c_f32p, c_f32p, C.POINTER(C.c_int64), C.c_void_p]),
    "nsf_peaknorm_pcm16": (i32, [c_f32p, i32, i64, c_f32p, C.c_void_p, C.c_void_p]),
    "nsf_pcm16_to_float_interleaved": (i32, [C.c_void_p, i32, i64, c_f32p, C.c_void_p]),
    "nsf_attention_test_workspace_bytes": (i64, [i32, i32, i32, i32]),
    "nsf_attention_test": (i32, [c_f32p, c_f32p, c_f32p, c_f32p, i32, i32, i32, i32, c_f32p, C.c_void_p, i64, C.c_void_p]),
    "nsf_mask_apply": (i32, [c_f32p, i32, i32, c_f32p, i64, i64, i32, i64, i32, i32, i32, i32, C.c_float, c_f32p, C.c_void_p]),
    "nsf_segment_power_norm": (i32, [c_f32p, i32, c_f32p, i64, i64, i32, i64, i32, i32, i32, i32, i64, c_f32p, C.c_void_p]),
    "nsf_gather_crops": (i32, [C.c_void_p, i32, i64, C.c_void_p, C.c_void_p, C.c_void_p, i32, i64, c_f32p, C.c_void_p]),
    "nsf_whisper_encoder_num_offsets": (i64, [C.c_void_p]),
    "nsf_whisper_encoder_create": (i32, [C.c_void_p, c_f32p, i64, C.POINTER(i64), i32, C.POINTER(C.c_void_p)]),
    "nsf_whisper_encoder_destroy": (None, [C.c_void_p]),
    "nsf_whisper_encoder_workspace_bytes": (i64, [C.c_void_p, i32]),
    "nsf_whisper_mel_plane_elems": (i64, [i32, i32]),
    "nsf_whisper_logmel": (i32, [c_f32p, i32, i64, c_f32p, i32, c_f32p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsf_whisper_logmel_recording": (i32, [c_f32p, i64, c_f32p, i32, i64, c_f32p, C.c_void_p, C.c_void_p]),
    "nsf_whisper_mel_windows": (i32, [c_f32p, i64, C.c_void_p, i32, C.c_void_p, C.c_void_p, i32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsf_whisper_encoder_forward": (i32, [C.c_void_p, C.c_void_p, C.c_void_p, i32, c_f32p, C.c_void_p, C.c_void_p, i64, C.c_void_p]),
    "nsf_whisper_decoder_num_offsets": (i64, [C.c_void_p]),
    "nsf_whisper_decoder_create": (i32, [C.c_void_p, c_f32p, i64, C.POINTER(i64), i32, C.POINTER(C.c_void_p)]),
    "nsf_whisper_decoder_destroy": (None, [C.c_void_p]),
    "nsf_whisper_decoder_state_bytes": (i64, [C.c_void_p, i32]),
    "nsf_whisper_decoder_prefill_cross": (i32, [C.c_void_p, C.c_void_p, i32, C.c_void_p, i64, C.c_void_p]),
    "nsf_whisper_decoder_step_dev": (i32, [C.c_void_p, C.c_void_p, C.c_void_p, i32, C.c_void_p, i64, C.c_void_p, i32, i32, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsf_whisper_decoder_forward": (i32, [C.c_void_p, C.c_void_p, C.c_void_p, i32, C.c_void_p, i64, c_f32p, C.c_void_p]),
    "nsf_whisper_decoder_reorder_scratch_bytes": (i64, [C.c_void_p, i32]),
    "nsf_whisper_decoder_reorder": (i32, [C.c_void_p, C.c_void_p, C.c_void_p, i32, C.c_void_p, i64, C.c_void_p, i64, C.c_void_p]),
    "nsf_whisper_decoder_step": (i32, [C.c_void_p, C.c_void_p, i32, i32, C.c_void_p, i64, c_f32p, C.c_void_p, C.c_void_p]),
    "nsf_flash_attention_test_workspace_bytes": (i64, [i32, i32, i32]),
    "nsf_flash_attention_test": (i32, [c_f32p, c_f32p, c_f32p, i32, i32, i32, c_f32p, C.c_void_p, i64, C.c_void_p]),
    "nsf_attention16_test": (i32, [c_f32p, c_f32p, c_f32p, c_f32p, i32, i32, i32, i32, c_f32p, C.c_void_p, i64, C.c_void_p]),
    "nsf_whisper_logit_rules": (i32, [c_f32p, i32, i32, C.c_void_p, i32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsf_whisper_decoder_step_rules": (i32, [C.c_void_p, C.c_void_p, C.c_void_p, i32, C.c_void_p, i64, C.c_void_p, i32, i32, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, i32, C.c_void_p]),
    "nsf_whisper_alignment_workspace_bytes": (i64, [i32, i32, i32, i32]),
    "nsf_whisper_alignment": (i32, [c_f32p, i32, i32, i32, i32, i32, C.c_void_p, C.c_void_p, c_f32p, C.c_void_p, i64, C.c_void_p]),
    "nsf_titanet_num_offsets": (i64, [C.c_void_p]),
    "nsf_titanet_create": (i32, [C.c_void_p, c_f32p, i64, C.POINTER(i64), i32, C.POINTER(C.c_void_p)]),
    "nsf_titanet_destroy": (None, [C.c_void_p]),
    "nsf_titanet_workspace_bytes": (i64, [C.c_void_p, i32, i32]),
    "nsf_titanet_features": (i32, [c_f32p, C.c_void_p, i32, i64, i32, c_f32p, i32, c_f32p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsf_titanet_forward": (i32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, i32, i32, c_f32p, C.c_void_p, i64, C.c_void_p]),
    "nsf_cos_affinity_accum": (i32, [c_f32p, i64, i32, i32, C.c_float, c_f32p, c_f32p, C.c_void_p, c_f32p, C.c_void_p]),
    "nsf_gemm_test": (i32, [i32, c_f32p, c_f32p, c_f32p, c_f32p, i32, i32, i32, C.c_void_p, i64, C.c_void_p]),
}

_lib = None


class NsfError(RuntimeError):
    pass


def load(build_if_missing: bool = True):
    """dlopen the library (building it first if it is absent and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise NsfError(f"{LIB_PATH} is missing; run `python __graft_entry__.py build`")
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().nsf_last_error().decode(errors="replace")
        raise NsfError(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    """device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def prof_collect():
    """{class name: (ms, work, brackets)} since the previous collect (synchronises the device)."""
    lib = load()
    n = lib.nsf_prof_num_classes()
    ms, work, cnt = (C.c_double * n)(), (C.c_double * n)(), (i64 * n)()
    check(lib.nsf_prof_collect(ms, work, cnt, n), "nsf_prof_collect")
    return {lib.nsf_prof_class_name(i).decode(): (ms[i], work[i], cnt[i]) for i in range(n)}
