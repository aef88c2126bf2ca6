"""notsofar_b200 -- B200-native CSS hot path of the NOTSOFAR-1 baseline.

The directory is named ``notsofar1-challenge_b200`` (repository contract); it is importable as
``notsofar_b200`` through the shim package of that name at the repository root.
"""
from .css import (CssCfg, css_inference, separate_and_stitch, calc_segment_weight, plan_segments, plan_batches,
                  permutation_chain, load_css_model, load_audio, write_wav)
from .separator import ConformerCssB200, pack_weights
from .diarization import DiarizationCfg, diarization_inference
from .asr import WhisperAsrCfg, asr_inference
from ._cabi import NsfError, GEMM_SIMT_FP32, GEMM_TC_3XTF32, GEMM_TC_TF32, GEMM_TC_2XBF16, GEMM_TC_2XF16, GEMM_TC_BF16

__all__ = ["CssCfg", "css_inference", "separate_and_stitch", "calc_segment_weight", "plan_segments",
           "permutation_chain", "load_css_model", "load_audio", "write_wav", "ConformerCssB200", "pack_weights",
           "DiarizationCfg", "diarization_inference", "WhisperAsrCfg", "asr_inference", "NsfError", "GEMM_SIMT_FP32", "GEMM_TC_3XTF32", "GEMM_TC_TF32", "GEMM_TC_2XBF16", "GEMM_TC_2XF16", "GEMM_TC_BF16"]
