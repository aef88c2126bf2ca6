/*
 * nsf_b200.h -- C ABI of the B200-native NOTSOFAR CSS hot path (libnsf_b200.so).
 *
 * The reference (microsoft/NOTSOFAR1-Challenge) is pure Python and has no FFI of its own;
 * its drop-in boundary is the Python plug-in `css_inference` / `separate_and_stitch`
 * (css/css.py:51,110).  This header is the thin native layer *below* that boundary: one
 * entry point per stage of the reference's hot path, each citing the reference code it
 * replaces.  `notsofar_b200/css.py` mirrors the reference's Python signatures on top of it.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (HBM) owned by the caller unless marked `host`;
 *     nothing is allocated or freed inside except by nsf_conformer_create/destroy
 *     (a small host-side handle; weights stay in the caller's blob);
 *   - all calls are stream-ordered on `stream` (a cudaStream_t passed as void*), do not
 *     synchronise, and are re-entrant;
 *   - return value: 0 = ok, negative = error (NSF_ERR_*); nsf_last_error() gives the text
 *     (thread-local);
 *   - complex data is interleaved (re, im) float32 ("c64"), as torch.complex64 / numpy
 *     complex64 store it;
 *   - F = 257 bins (frame 512, hop 256), C = 7 microphones, S = 3 speakers, Nn noise masks.
 *
 * HBM layouts (row-major, last index fastest)
 *   audio      x      [N][C]                    f32   (reference speech_mix[0], css.py:110)
 *   mixture    X      [F][T_long][C]            c64   (reference stft_mix[0], css.py:155)
 *   features   feat   [n_seg*T][ldf]            f32   (time-major rows; ldf >= 257*C)
 *   masks      masks  [n_seg][S+Nn][F][T]       f32   (reference: tuple of [1,F,T], conformer.py:308)
 *   separated  Y      [n_seg][S][F][T]          c64   (reference separated_seg, css.py:227)
 *   stitched   mask_st[F][T_long][S]            f32   (reference side_info['mask_stitched'][0])
 *              S_st   [S][T_long][F]            c64   (frame-major, feeds the iSTFT)
 *   waveforms  wav    [S][N_out]                f32   N_out = (T_long-1)*256 + 512
 */
#ifndef NSF_B200_H_
#define NSF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSF_OK                 0
#define NSF_ERR_INVALID_ARG   -1
#define NSF_ERR_CUDA          -2
#define NSF_ERR_UNSUPPORTED   -3

#define NSF_NUM_BINS   257
#define NSF_FRAME_LEN  512
#define NSF_FRAME_HOP  256

/* GEMM engines of the mask network */
#define NSF_GEMM_SIMT_FP32   0   /* CUDA-core fp32 FMA (bit-faithful fp32 products; debugging / cross-check) */
#define NSF_GEMM_TC_3XTF32   1   /* tcgen05 kind::tf32, error-compensated 3-pass split (fp32-parity mode, default) */
#define NSF_GEMM_TC_TF32     2   /* tcgen05 kind::tf32, single pass (throughput mode, ~1e-3) */
#define NSF_GEMM_TC_2XBF16   3   /* tcgen05 kind::f16 on bf16 head + remainder pairs, 3 MMAs per product (~2^-17, fp32 range) */
#define NSF_GEMM_TC_2XF16    4   /* tcgen05 kind::f16 on power-of-two-scaled fp16 head + remainder pairs, 3 MMAs per product
                                    (~2^-22: fp32-grade at twice the 3xTF32 rate); scaled magnitudes saturate at 65504 */

#define NSF_GEMM_TC_BF16     5   /* tcgen05 kind::f16, plain bf16 operands, one MMA per product (the Whisper encoder's engine) */

const char* nsf_last_error(void);
const char* nsf_version(void);
/* number of kernels this library has launched in the process so far (bench.py's gpu_launches) */
int64_t nsf_launch_count(void);

/* Optional kernel-class profiler (bench.py's roofline): when enabled, every entry point brackets its launches
 * with CUDA events on the launching stream.  nsf_prof_collect synchronises the device and returns, per class,
 * the summed event time [ms], the summed algorithmic work (bytes for HBM-bound classes, flops for GEMMs) and
 * the number of brackets since the previous collect. */
int nsf_prof_enable(int on);
int nsf_prof_num_classes(void);
const char* nsf_prof_class_name(int cls);
int nsf_prof_collect(double* ms /*host*/, double* work /*host*/, int64_t* count /*host*/, int n_classes);

/* Number of STFT frames of an n_samples signal: conv1d stride 256, kernel 512, no padding
 * (css_with_conformer/executor/feature.py:105). */
int64_t nsf_num_frames(int64_t n_samples);

/* Multichannel STFT.  Replaces ConformerCssWrapper.stft (css/training/conformer_wrapper.py:106-129)
 * -> STFT.forward (feature.py:88-128, kernel init_kernel feature.py:19-45):
 *   X[k][t][c] = sum_n x[256 t + n][c] * hann_periodic(n) * exp(-2 pi i k n / 512),  no padding, scale 1.
 * The DC and Nyquist bins carry the residue the reference's (mag, atan2) -> th.polar round trip
 * leaves: imag = re * sin(pi_f32) when re < 0, else +0.
 * Computes frames 0 .. n_frames-1 of x (n_frames <= nsf_num_frames(n_samples)) into X[:, 0..n_frames-1, :];
 * T_long is the frame pitch of X (>= n_frames).  A rank that owns a frame range passes offset pointers. */
int nsf_stft_mc(const float* x, int64_t n_samples, int n_ch,
                float* X, int64_t T_long, int64_t n_frames, void* stream);

/* Segment features.  Replaces the front half of ConformerCssWrapper.separate
 * (conformer_wrapper.py:91-94) + FeatureExtractor.forward (feature.py:543-569): MVN magnitude of
 * mic 0 (feature.py:496-507) and mean-normalised IPD v1 of mics 1..C-1 vs mic 0 (feature.py:214-221).
 * Segment i (i < n_seg) covers frames [(seg_first + i)*hop, +T) of X; frames >= T_valid read as zeros
 * (the zero-padded tail of the last segment, css.py:185-190; T_valid <= T_long = frame pitch of X).
 * Row (i*T + t), column (m*257 + f).  If in_bias/in_scale are non-NULL the network's input
 * normalisation (f + bias) * scale (conformer.py:297-299) is fused.  If feat_lo is non-NULL the
 * value is stored split for the tensor-core GEMM in format split_fmt: NSF_SPLIT_TF32 -> two fp32 arrays (TF32 head,
 * exact remainder); NSF_SPLIT_BF16 / NSF_SPLIT_F16 -> feat and feat_lo are uint16 arrays of pitch ldf holding bf16 /
 * (x 2^4-scaled) fp16 head and remainder, for the NSF_GEMM_TC_2XBF16 / _2XF16 engines.
 * Columns [257*n_ch, ldf) of every row (the K padding of the first GEMM) are written as zeros. */
#define NSF_SPLIT_TF32 0
#define NSF_SPLIT_BF16 1
#define NSF_SPLIT_F16  2
int nsf_css_features(const float* X, int64_t T_long, int64_t T_valid, int n_ch, int64_t seg_first, int n_seg,
                     int T, int hop, const float* in_bias, const float* in_scale,
                     float* feat, float* feat_lo, int64_t ldf, int split_fmt, void* stream);

/* Mask network (ConformerCSS.forward, css_with_conformer/nnet/conformer.py:287-310, eval mode).
 * The handle only records dimensions and the offsets of each tensor inside the caller's device blob
 * (layout documented in notsofar_b200/separator.py::pack_weights).  */
typedef struct nsf_conformer nsf_conformer;
typedef struct {
    int d_model, n_heads, d_ff, n_blocks, kernel_size, in_features, n_out, maxlen;
    int T;            /* frames per segment */
    int gemm_engine;  /* NSF_GEMM_* */
} nsf_conformer_dims;

int nsf_conformer_create(const nsf_conformer_dims* dims, const float* blob, int64_t blob_floats,
                         const int64_t* offsets /*host*/, int n_offsets, nsf_conformer** out);
void nsf_conformer_destroy(nsf_conformer* h);
int64_t nsf_conformer_num_offsets(const nsf_conformer_dims* dims);
/* 1 if a handle with these dimensions runs with two LayerNorms of every block folded into the GEMMs on either side
 * (the attention LayerNorm into the QKV projection, the feed_forward_out LayerNorm into its first linear layer;
 * conformer.py:153-156,180-182): the blob must then hold gamma-scaled weights, bias + W beta and the column sums of the
 * scaled weights for those two layers (pack_weights does; 2xBF16 engine, d_model = 128 * {1,2,4}; NSF_LN_FOLD=0 disables). */
int nsf_conformer_ln_fold(const nsf_conformer_dims* dims);
/* bytes of scratch HBM nsf_conformer_forward needs for a batch of n_seg segments */
int64_t nsf_conformer_workspace_bytes(const nsf_conformer_dims* dims, int n_seg);
/* feat (+feat_lo): [n_seg*T][ldf] already input-normalised, in the split format of the handle's engine
 * (see nsf_css_features; ldf counts elements);
 * masks: [n_seg][n_out/257][257][T] = sigmoid(linear(...)). */
int nsf_conformer_forward(nsf_conformer* h, const float* feat, const float* feat_lo, int64_t ldf, int n_seg,
                          float* masks, void* workspace, int64_t workspace_bytes, void* stream);

/* Mask-weighted MVDR.  Replaces make_mvdr(..., return_stft=True)
 * (css_with_conformer/utils/mvdr_util.py:5-47: make_wta :50-55, get_mask_scm :58-66,
 * calc_bfcoeffs :69-75, get_bf :78-80) and the floored-mask multiply of css.py:223-227.
 * Arithmetic: spatial covariances, solve and beamformer in fp64 on chip (the reference's
 * complex64 solve is only ~1e-2 accurate in ill-conditioned bins; parity is against the
 * fp64-lifted reference).  mask_floor = 10^(floor_dB/20); 1.0 => pure MVDR.
 * Segment geometry as in nsf_css_features.  X is [n_bins][T_long][C]; masks [n_seg][S+Nn][n_bins][T];
 * Y [n_seg][S][n_bins][T].  Bin 0 gets the den += 1e-15 of mvdr_util.py:73.
 * Built for S = 2..4 speaker masks and C = 7 microphones (NSF_ERR_UNSUPPORTED otherwise; the reference's covariance code
 * is itself written for 7 microphones, mvdr_util.py:63); any T and hop.  When T == 2 hop (50 % overlap: every shipped
 * configuration) and hop <= 96 the streaming kernel runs -- a warp walks a run of consecutive segments of one bin and forms
 * every outer product once for the two segments that share the frame -- otherwise one warp per (segment, bin). */
int nsf_mvdr(const float* masks, int n_spk, int n_noise, const float* X, int64_t T_long, int64_t T_valid, int n_ch,
             int64_t seg_first, int n_seg, int T, int hop, int n_bins, float mask_floor,
             float* Y, void* stream);

/* The same beamformer on ONE utterance of any length (make_mvdr accepts any T, mvdr_util.py:5-47; BASELINE config 5): one covariance
 * set per bin over all T frames, accumulated by T-chunks in parallel (partial sums -> per-bin solve -> apply).
 * masks [S+Nn][n_bins][T], X [n_bins][T][C], Y [S][n_bins][T]; workspace: nsf_mvdr_utterance_workspace_bytes, 16-byte aligned. */
int64_t nsf_mvdr_utterance_workspace_bytes(int n_spk, int64_t T, int n_bins);
int nsf_mvdr_utterance(const float* masks, int n_spk, int n_noise, const float* X, int64_t T, int n_ch, int n_bins, float mask_floor,
                       float* Y, void* workspace, int64_t workspace_bytes, void* stream);

/* Separation without the beamformer (single-channel input, or CssCfg.mc_mvdr = False): css.py:218-227,
 *   Y[seg][s][f][t] = X[f][(seg_first+seg)*hop + t][channel 0] * max(masks[seg][s][f][t], mask_floor).
 * Layouts and segment geometry as in nsf_mvdr; X may have any number of channels (the reference channel is 0). */
int nsf_mask_apply(const float* masks, int n_spk, int n_noise, const float* X, int64_t T_long, int64_t T_valid, int n_ch,
                   int64_t seg_first, int n_seg, int T, int hop, int n_bins, float mask_floor, float* Y, void* stream);

/* CssCfg.normalize_segment_power (css.py:233-247): every segment of Y [n_seg][S][n_bins][T] is scaled in place by
 *   sqrt(mean_{f,t<t_seg} |X_ref|^2) / sqrt(mean_{f,t<t_seg} |sum_s Y_s|^2),   t_seg = min(T, mix_frames - segment start),
 * mix_frames = frames of the (possibly zero-padded) long-form STFT (css.py:159-169).  ratio [n_seg] f32 scratch / output. */
int nsf_segment_power_norm(float* Y, int n_spk, const float* X, int64_t T_long, int64_t T_valid, int n_ch, int64_t seg_first,
                           int n_seg, int T, int hop, int n_bins, int64_t mix_frames, float* ratio, void* stream);

/* PIT stitching costs.  Replaces PitWrapper._opt_perm_loss (css/training/losses.py:50-71) as called
 * at css.py:276: cost[i][a][b] = mean_{f,t} loss(left_a, right_b) over the `overlap` trailing frames
 * of segment i-1 and leading frames of segment i, for i = 1..n_seg-1 (cost[0] is zero).
 * input_kind 0: masks [n_seg][n_ch_total][F][T] f32 (first n_spk channels); 1: |Y| with Y [n_seg][n_spk][F][T] c64.
 * loss_kind 0: l1 (losses.py:104), 1: mse (losses.py:100).  The (tiny, sequential) permutation chain
 * itself runs on the host. */
int nsf_pit_cost(const void* in, int input_kind, int loss_kind, int n_seg, int n_ch_total, int n_spk,
                 int n_bins, int T, int overlap, float* cost, void* stream);
/* The same for segments [seg_begin, seg_end) only (`in` and `cost` still address segment 0; segment seg_begin - 1 must exist):
 * the costs of a chunk of segments as soon as its masks are there, while later chunks are still in the network. */
int nsf_pit_cost_range(const void* in, int input_kind, int loss_kind, int seg_begin, int seg_end, int n_ch_total, int n_spk,
                       int n_bins, int T, int overlap, float* cost, void* stream);

/* Weighted overlap-add of the permuted segment masks (css.py:254-299) and the activity mean
 * (css.py:304).  seg_w [n_seg][T] f32 trapezoid weights (calc_segment_weight css.py:341-390, rows
 * already specialised for first/last segment), wsum [T_long] their overlap-added sum,
 * perms [n_seg][S] int32 (new channel k <- old channel perms[i][k]).
 * mask_st [F][T_long][S]; activity [T_long][S] = mean_f mask_st. */
int nsf_stitch_masks(const float* masks, int n_ch_total, const int32_t* perms, const float* seg_w,
                     const float* wsum, int n_seg, int n_spk, int n_bins, int T, int hop, int64_t T_long,
                     float* mask_st, float* activity, void* stream);

/* Activity gate: (activity >= th) -> dilate(dil) -> erode(ero) per speaker
 * (css.py:305-309, utils/numpy_utils.py:4-13).  act_b / act_final: [T_long][S] uint8; tmp same size. */
int nsf_activity(const float* activity, int64_t T_long, int n_spk, float th, int dil, int ero,
                 uint8_t* act_b, uint8_t* tmp, uint8_t* act_final, void* stream);

/* Weighted overlap-add of the permuted separated STFTs, normalisation, activity gating
 * (css.py:287-299, 312) and the [B*S, F, T] re-layout of css.py:316.  S_st [S][T_long][F] c64. */
int nsf_stitch_stft(const float* Y, const int32_t* perms, const float* seg_w, const float* wsum,
                    const uint8_t* act_final, int n_seg, int n_spk, int n_bins, int T, int hop,
                    int64_t T_long, float* S_st, void* stream);

/* Inverse STFT.  Replaces ConformerCssWrapper.istft (conformer_wrapper.py:131-146) -> iSTFT.forward
 * (feature.py:138-167): y[n] = sum_t g[n - 256 t] * Re sum_{k<=256} S[k][t] exp(+2 pi i k (n-256t)/512),
 * g = sqrt(hann_periodic)/16; no one-sided doubling, no window-sum normalisation.
 * S_st [n_streams][T_long][F] c64 -> wav [n_streams][(T_long-1)*256+512] f32. */
int nsf_istft(const float* S_st, int n_streams, int64_t T_long, float* wav, void* stream);
/* Hops (256-sample blocks of the output) [hop_begin, hop_end) only; reads frames hop_begin-1 .. hop_end-1.  hop_begin must be a
 * multiple of 8 (the kernel's tile: the launch then pairs frames, and rounds, exactly like nsf_istft). */
int nsf_istft_range(const float* S_st, int n_streams, int64_t T_long, float* wav, int64_t hop_begin, int64_t hop_end, void* stream);

/* Progressive tail of separate_and_stitch (css.py:287-338 behind the permutation chain): nsf_stitch_masks, nsf_activity,
 * nsf_stitch_stft and nsf_istft restricted to what became final when the segments [seg_prev, seg_done) were added to
 * [0, seg_prev) -- every stage is local in time, so the calls for seg_done = s_1 < s_2 < ... < n_seg together write bit for bit
 * what the four whole-recording calls write.  perms must hold the rows of segments < seg_done; all buffers are the
 * whole-recording ones (same layouts as above).  hops_written (host, int64[2]) receives the range of waveform hops
 * (256-sample blocks of every stream of wav) this call produced: the caller can start their device -> host copy while later
 * segments are still in the mask network. */
int nsf_stitch_progress(const float* masks, int n_ch_total, const float* Y, const int32_t* perms, const float* seg_w,
                        const float* wsum, int n_seg, int seg_prev, int seg_done, int n_spk, int n_bins, int T, int hop,
                        int64_t T_long, float th, int dil, int ero, float* mask_st, float* activity, uint8_t* act_b,
                        uint8_t* tmp, uint8_t* act_final, float* S_st, float* wav, int64_t* hops_written, void* stream);

/* File boundary.  Replaces write_wav's peak normalisation (utils/audio_utils.py:44-45) and
 * libsndfile's float -> PCM_16 conversion: q = rint(x * 0.99 / (max|x| + 1e-7) * 32767).
 * peak [n_streams] f32 scratch/output (max |x| per stream). */
int nsf_peaknorm_pcm16(const float* wav, int n_streams, int64_t n, float* peak, int16_t* pcm, void* stream);

/* File boundary, input side.  Replaces load_audio (css/helpers.py:40-65: soundfile.read of 7 mono files, np.stack(axis=-1)) for
 * 16-bit files: pcm [n_ch][n] int16 (one row per channel file) -> out [n][n_ch] f32 = pcm / 32768 (libsndfile's int16 -> float32
 * scaling, exact).  n_ch <= 8. */
int nsf_pcm16_to_float_interleaved(const int16_t* pcm, int n_ch, int64_t n, float* out, void* stream);

/* CSS -> diarization hand-off without the disk round trip.  Replaces read_wav(normalize=True) (utils/audio_utils.py:10-34:
 * int16 / 32767) + the per-(word, scale) slicing and pad_sequence of extract_speaker_embedding_for_words
 * (diarization/word_based_diarization.py:78-104): out[i][j] = pcm[stream_id[i]][start[i] + j] / 32767 for j < len[i], else 0.
 * pcm [n_streams][n] int16 (the output of nsf_peaknorm_pcm16), stream_id / start / len [n_crops] on the device,
 * out [n_crops][max_len] f32.  The integer crop plan itself is host logic (notsofar_b200.diarization.word_crop_plan). */
int nsf_gather_crops(const int16_t* pcm, int n_streams, int64_t n, const int32_t* stream_id, const int64_t* start,
                     const int32_t* len, int n_crops, int64_t max_len, float* out, void* stream);

/* Test hook: C[M][N] = A[M][K] * W[N][K]^T + bias with the selected engine (row-major, K contiguous;
 * K % 32 == 0, lda/ldw multiples of 4).  Used by the parity tests to check the tcgen05 GEMM
 * against the CUDA-core one. */
int nsf_gemm_test(int engine, const float* A, const float* W, const float* bias, float* Cout,
                  int M, int N, int K, void* workspace, int64_t workspace_bytes, void* stream);

/* Test hook: the fused relative-position attention of the mask network (conformer.py:66-92) on plain fp32 inputs.
 * q, k, v [n_seg*n_heads][T][64], pe [2*maxlen][64] (Embedding table of RelativePositionalEncoding, conformer.py:18)
 *   out[seg*T + t1][head*64 + d] = sum_t2 softmax_t2((q[t1].k[t2] + q[t1].pe[maxlen + t1 - t2]) / 8) v[t2][d].
 * T <= 192.  Used by the parity tests to check the tcgen05 attention kernel against a float64 restatement. */
int64_t nsf_attention_test_workspace_bytes(int n_seg, int n_heads, int T, int maxlen);
int nsf_attention_test(const float* q, const float* k, const float* v, const float* pe, int maxlen, int n_seg, int n_heads,
                       int T, float* out, void* workspace, int64_t workspace_bytes, void* stream);

/* The same test hook for the bf16-pair attention kernel of the NSF_GEMM_TC_2XBF16 engine (attention16.cu); workspace as above. */
int nsf_attention16_test(const float* q, const float* k, const float* v, const float* pe, int maxlen, int n_seg, int n_heads,
                         int T, float* out, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- Whisper audio encoder (row a15: asr/asr.py:69-74 calls openai-whisper, third-party and unpinned; the published
 * algorithm of whisper/audio.py::log_mel_spectrogram and whisper/model.py::AudioEncoder is restated here, parity against
 * the transformers implementation of the same model -- see csrc/whisper.cu).  30-s chunks of 480 000 samples. */
typedef struct nsf_whisper_encoder nsf_whisper_encoder;
typedef struct { int n_mels, n_ctx, d_model, n_heads, n_layers, d_ff; } nsf_whisper_dims;
int64_t nsf_whisper_encoder_num_offsets(const nsf_whisper_dims* dims);
/* blob layout: notsofar_b200/whisper.py::pack_whisper_encoder (bf16 planes packed two per float word, fp32 biases / LN / pos) */
int nsf_whisper_encoder_create(const nsf_whisper_dims* dims, const float* blob, int64_t blob_floats, const int64_t* offsets /*host*/,
                               int n_offsets, nsf_whisper_encoder** out);
void nsf_whisper_encoder_destroy(nsf_whisper_encoder* h);
int64_t nsf_whisper_encoder_workspace_bytes(const nsf_whisper_dims* dims, int n_batch);
/* elements of one bf16 plane of the time-major, zero-framed log-mel input: n_batch * 3002 * n_mels */
int64_t nsf_whisper_mel_plane_elems(int n_mels, int n_batch);
/* audio [n_batch][480000] f32 -> log-mel as bf16 head / remainder planes [n_batch][3002][n_mels] (row 0 and 3001 zero).
 * filters [n_mels][201] f32 (slaney mel filterbank); log_spec [n_batch][n_mels][3000] f32 and gmax [n_batch] u32 are scratch. */
int nsf_whisper_logmel(const float* audio, int n_batch, int64_t n_samples, const float* filters, int n_mels, float* log_spec,
                       uint32_t* gmax, void* mel_hi, void* mel_lo, void* stream);
/* Whole-recording front end of transcribe() (whisper/audio.py log_mel_spectrogram with padding = 30 s, whisper/transcribe.py
 * [upstream]; call site asr/asr.py:74): audio [n_samples] f32 (the recording followed by 30 s of zeros) -> log_spec [n_mels][n_frames]
 * f32 = log10(max(mel power, 1e-10)) of the n_frames = n_samples / 160 centred frames (reflect padding at both ends) and gmax
 * [1] = its order-encoded maximum; nsf_whisper_mel_windows then cuts n_windows 30-s windows starting at frames seeks[] (device
 * int32) holding sizes[] content frames each (NULL: 3000; the rest of a window is zero like pad_or_trim) and applies
 * max(., gmax - 8), (. + 4) / 4: the recording-wide normalisation of the reference, not a per-window one. */
int nsf_whisper_logmel_recording(const float* audio, int64_t n_samples, const float* filters, int n_mels, int64_t n_frames,
                                 float* log_spec, uint32_t* gmax, void* stream);
int nsf_whisper_mel_windows(const float* log_spec, int64_t n_frames, const uint32_t* gmax, int n_mels, const int32_t* seeks,
                            const int32_t* sizes, int n_windows, void* mel_hi, void* mel_lo, void* stream);
/* out [n_batch * 1500][d_model] f32 = ln_post(encoder(mel)); out_bf16 (optional): the same as a bf16 plane, the operand
 * nsf_whisper_decoder_prefill_cross reads */
int nsf_whisper_encoder_forward(nsf_whisper_encoder* h, const void* mel_hi, const void* mel_lo, int n_batch, float* out, void* out_bf16,
                                void* workspace, int64_t workspace_bytes, void* stream);

/* ---- Whisper text decoder, greedy step (csrc/whisper_dec.cu; whisper/model.py TextDecoder [upstream, unpinned]).
 * state: one caller-owned device buffer of nsf_whisper_decoder_state_bytes (cross / self key-value caches in bf16,
 * activations).  prefill_cross projects the audio features (bf16 plane [n_batch*n_audio_ctx][d_model], written by
 * nsf_whisper_encoder_forward) into every layer's cross-attention caches; step consumes one token per sequence at
 * position pos (0-based, the initial prompt is fed token by token) and returns the arg-max next tokens
 * (and, if logits_out != NULL, the fp32 logits [n_batch][vocab]). */
typedef struct nsf_whisper_decoder nsf_whisper_decoder;
typedef struct { int vocab, n_text_ctx, d_model, n_heads, n_layers, d_ff, n_audio_ctx; } nsf_whisper_dec_dims;
int64_t nsf_whisper_decoder_num_offsets(const nsf_whisper_dec_dims* dims);
int nsf_whisper_decoder_create(const nsf_whisper_dec_dims* dims, const float* blob, int64_t blob_floats, const int64_t* offsets /*host*/,
                               int n_offsets, nsf_whisper_decoder** out);
void nsf_whisper_decoder_destroy(nsf_whisper_decoder* h);
int64_t nsf_whisper_decoder_state_bytes(const nsf_whisper_dec_dims* dims, int n_batch);
int nsf_whisper_decoder_prefill_cross(nsf_whisper_decoder* h, const void* enc_bf16, int n_batch, void* state, int64_t state_bytes, void* stream);
int nsf_whisper_decoder_step(nsf_whisper_decoder* h, const int32_t* tokens, int pos, int n_batch, void* state, int64_t state_bytes,
                             float* logits_out, int32_t* next_tokens, void* stream);

/* One decoder position without the sampling bookkeeping: logits_out [n_batch][vocab] f32 of tokens[b] at position *pos_dev (the
 * self-attention keys / values of that position are appended to the caches).  Beam search and temperature sampling
 * (whisper/decoding.py BeamSearchDecoder / GreedyDecoder [upstream]; the reference decodes with beam_size = 5, asr/asr.py:17-21,52-56)
 * choose the next tokens from these logits on the host side of the API; nsf_whisper_decoder_reorder then makes sequence b
 * continue the hypothesis of slot src[b] (rearrange_kv_cache [upstream]): rows [0, *n_pos_dev) of the self-attention caches are
 * permuted through `scratch` (nsf_whisper_decoder_reorder_scratch_bytes). */
int nsf_whisper_decoder_forward(nsf_whisper_decoder* h, const int32_t* tokens, const int32_t* pos_dev, int n_batch, void* state,
                                int64_t state_bytes, float* logits_out, void* stream);
int64_t nsf_whisper_decoder_reorder_scratch_bytes(const nsf_whisper_dec_dims* dims, int n_batch);
int nsf_whisper_decoder_reorder(nsf_whisper_decoder* h, const int32_t* src, const int32_t* n_pos_dev, int n_batch, void* state,
                                int64_t state_bytes, void* scratch, int64_t scratch_bytes, void* stream);
/* The same step with all loop state on the device, so that one captured CUDA graph can be replayed for every position:
 * consumes cur_tokens [n_batch] at position *pos_dev, then (one bookkeeping kernel) records the arg-max in
 * argmaxes[b][pos+1], chooses the token fed next -- forced[b][pos+1] if >= 0 (prompt / teacher forcing), eot once the
 * sequence is done, else the arg-max -- into out_tokens[b][pos+1] and cur_tokens, marks sequences that produced eot in
 * done, and increments *pos_dev.  forced / out_tokens / argmaxes: [n_batch][total_len] int32. */
int nsf_whisper_decoder_step_dev(nsf_whisper_decoder* h, int32_t* cur_tokens, int32_t* pos_dev, int n_batch, void* state, int64_t state_bytes,
                                 const int32_t* forced, int total_len, int eot, int32_t* out_tokens, int32_t* argmaxes, uint8_t* done,
                                 void* stream);

/* Logit filters of the decoding loop [upstream whisper/decoding.py: SuppressBlank, SuppressTokens, ApplyTimestampRules; the
 * reference decodes with timestamps, asr/asr.py:52-56 -> whisper.transcribe], applied in place to logits [n_batch][vocab]
 * before the arg-max.  tokens [n_batch][total_len] holds the fed tokens up to position *pos_dev; sample_begin = length of the
 * prompt (sot sequence); timestamp_begin < 0 switches the timestamp rules off; max_initial_timestamp_index < 0: none;
 * suppress [n_suppress] token ids always suppressed, suppress_first [n_suppress_first] (blank tokens + eot) at the first
 * sampled position.  Same arithmetic as transformers' WhisperTimeStampLogitsProcessor (the pin used by the tests). */
typedef struct { int sample_begin, timestamp_begin, no_timestamps, eot, max_initial_timestamp_index, n_suppress, n_suppress_first; } nsf_whisper_rules;
int nsf_whisper_logit_rules(float* logits, int n_batch, int vocab, const int32_t* tokens, int total_len, const int32_t* pos_dev,
                            const nsf_whisper_rules* rules, const int32_t* suppress, const int32_t* suppress_first, void* stream);
/* nsf_whisper_decoder_step_dev with the filters between the logits and the arg-max (rules may be NULL) and, optionally, the
 * capture of the cross-attention softmax rows of the alignment heads (graph-replayable like it): xattn_probs
 * [n_batch][n_align][total_len][n_audio_ctx] f32 gets row *pos_dev of every head h of layer l with align_map[l][h] = slot >= 0
 * (align_map [n_layers][n_heads] int32 on the device, -1 elsewhere). */
int nsf_whisper_decoder_step_rules(nsf_whisper_decoder* h, int32_t* cur_tokens, int32_t* pos_dev, int n_batch, void* state,
                                   int64_t state_bytes, const int32_t* forced, int total_len, int eot, int32_t* out_tokens,
                                   int32_t* argmaxes, uint8_t* done, const nsf_whisper_rules* rules, const int32_t* suppress,
                                   const int32_t* suppress_first, float* xattn_probs, const int32_t* align_map, int n_align, void* stream);
/* Token-level timestamps from the captured weights [upstream whisper/timing.py: find_alignment / median_filter / dtw; the
 * reference transcribes with word_timestamps=True, asr/asr.py:52-56]: per (sequence, head, audio position) normalisation over
 * the tokens, median filter of width 7 along the audio axis, mean over the heads, dynamic time warping of the negated matrix;
 * start_frame [n_batch][n_tokens] int32 = audio position (20 ms units) at which the path enters each token.
 * weights [n_batch][n_heads][n_tokens][n_frames] f32 (rows of the tokens to align only); the first m_valid audio positions take
 * part (num_frames // 2 upstream); n_tokens_per_seq [n_batch] or NULL; cost_out [n_batch][n_tokens][m_valid] or NULL. */
int64_t nsf_whisper_alignment_workspace_bytes(int n_batch, int n_heads, int n_tokens, int n_frames);
int nsf_whisper_alignment(const float* weights, int n_batch, int n_heads, int n_tokens, int n_frames, int m_valid,
                          const int32_t* n_tokens_per_seq, int32_t* start_frame, float* cost_out, void* workspace,
                          int64_t workspace_bytes, void* stream);

/* ---- TitaNet speaker-embedding forward + multi-scale cosine affinity (row a16: diarization/word_based_diarization.py:26
 * loads NeMo's EncDecSpeakerLabelModel "titanet_large", :105 calls spk_model.forward(input_signal, input_signal_length) under
 * autocast, :171-177 build the per-scale affinity with NeMo's getCosAffinityMatrix and average it.  NeMo is third-party and
 * unpinned; the published architecture is restated in oracle/titanet_oracle.py -- see csrc/titanet.cu).
 * dims: Jasper blocks (filters / repeat / kernel / residual per block), attention channels, embedding size.
 * blob layout: notsofar_b200/titanet.py::pack_titanet (BatchNorms folded; GEMM weights as bf16 head / remainder planes). */
typedef struct nsf_titanet nsf_titanet;
typedef struct {
    int feat_in, n_blocks, att_ch, emb; int filters[8], repeat[8], kernel[8], residual[8];
    int precision;   /* 0: fp32-grade GEMMs (bf16 head + remainder planes, three MMAs per product; the blob holds both planes);
                        1: the reference's autocast() arithmetic -- fp16 operands, fp32 accumulation, one MMA per product (the blob's
                           "head" entries hold fp16 planes, the "remainder" entries are ignored) */
} nsf_titanet_dims;
int64_t nsf_titanet_num_offsets(const nsf_titanet_dims* dims);
int nsf_titanet_create(const nsf_titanet_dims* dims, const float* blob, int64_t blob_floats, const int64_t* offsets /*host*/,
                       int n_offsets, nsf_titanet** out);
void nsf_titanet_destroy(nsf_titanet* h);
int64_t nsf_titanet_workspace_bytes(const nsf_titanet_dims* dims, int n_crops, int t_pad);
/* crops [n_crops][max_len] f32 zero padded (the output of nsf_gather_crops), lengths [n_crops] -> normalised log-mel features
 * as bf16 head / remainder planes [n_crops][t_pad][n_mels] (frames >= n_frames zero) and n_frames [n_crops] = len / 160 + 1.
 * mel_filters [257][n_mels] f32 (bin major), lm_scratch [n_crops][t_pad][n_mels] f32, t_pad >= max_len / 160 + 1. */
int nsf_titanet_features(const float* crops, const int32_t* lengths, int n_crops, int64_t max_len, int t_pad,
                         const float* mel_filters, int n_mels, float* lm_scratch, void* feat_hi, void* feat_lo, int32_t* n_frames,
                         void* stream);
/* features -> emb [n_crops][dims.emb] f32 (the second output of spk_model.forward) */
int nsf_titanet_forward(nsf_titanet* h, const void* feat_hi, const void* feat_lo, const int32_t* n_frames, int n_crops, int t_pad,
                        float* emb, void* workspace, int64_t workspace_bytes, void* stream);
/* acc [n][n] += scale * minmax(cos_similarity(emb)) for one scale: rows of emb (pitch row_pitch floats) normalised by
 * (norm + 3.5e-4), unit diagonal, global min-max scaling (getCosAffinityMatrix [upstream]); scale = 1 / n_scales gives the mean
 * of word_based_diarization.py:177.  en_scratch [n][dim], sim_scratch [n][n], minmax [2] u32. */
int nsf_cos_affinity_accum(const float* emb, int64_t row_pitch, int dim, int n, float scale, float* en_scratch, float* sim_scratch,
                           uint32_t* minmax, float* acc, void* stream);

/* Test hook: non-causal multi-head attention with online softmax (flash_attn.cu, the Whisper encoder's attention) on fp32
 * inputs that are rounded to bf16 inside.  q, k, v [n_batch*n_heads][T][64] (already scaled), out [n_batch*T][n_heads*64] f32:
 *   out[b*T + t1][h*64 + d] = sum_t2 softmax_t2(q[t1].k[t2]) v[t2][d]. */
int64_t nsf_flash_attention_test_workspace_bytes(int n_batch, int n_heads, int T);
int nsf_flash_attention_test(const float* q, const float* k, const float* v, int n_batch, int n_heads, int T, float* out,
                             void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NSF_B200_H_ */
