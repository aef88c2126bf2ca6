"""The transcriber behind ``asr_inference``: ``model.transcribe(wav, task='transcribe', language=..., word_timestamps=True,
beam_size=5, hallucination_silence_threshold=2.0)`` of the reference (asr/asr.py:52-56,69-74) on the in-tree Whisper kernels.

The reference's implementation is the third-party package openai-whisper (requirements.txt, unpinned HEAD; absent offline together
with its weights and vocabulary -- SURVEY.md 8c: **parity unpinned**).  What follows restates the published algorithm of
whisper/tokenizer.py, whisper/decoding.py (DecodingTask, BeamSearchDecoder, GreedyDecoder, MaximumLikelihoodRanker),
whisper/timing.py (add_word_timestamps, merge_punctuations) and whisper/transcribe.py (the seek loop, temperature fallback, the
hallucination-silence rules) [upstream]; the arithmetic runs in libnsf_b200.so (log-mel, encoder, decoder step, logit rules,
alignment / DTW), the control flow -- a few scalars per decoded token -- on the host.

    tok = WhisperTokenizerLite.from_tiktoken_file("multilingual.tiktoken", num_languages=100)      # whisper/assets/*.tiktoken
    model = WhisperB200(state_dict)                                                                  # notsofar_b200.whisper
    set_transcriber(WhisperB200Transcriber(model, tok, alignment_heads=[(l, h), ...]))

or, with NSF_WHISPER_CKPT / NSF_WHISPER_VOCAB set, nothing at all: ``asr_inference`` builds the transcriber itself.
"""
from __future__ import annotations

import base64
import os
import string
import zlib
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi
from .whisper import HOP, N_FRAMES, WhisperB200, WhisperRules, apply_logit_rules, token_alignment

SAMPLE_RATE = 16000
FRAMES_PER_SECOND = SAMPLE_RATE // HOP            # 100 mel frames per second
TOKENS_PER_SECOND = FRAMES_PER_SECOND // 2        # 50 audio positions per second (20 ms)
LANGUAGES = ("en zh de es ru ko fr ja pt tr pl ca nl ar sv it id hi fi vi he uk el ms cs ro da hu ta no th ur hr bg lt la mi ml cy sk te fa lv bn "
             "sr az sl kn et mk br eu is hy ne mn bs kk sq sw gl mr pa si km sn yo so af oc ka be tg sd gu am yi lo uz fo ht ps tk nn mt sa lb my bo "
             "tl mg as tt haw ln ha ba jw su yue").split()


# ------------------------------------------------------------------------------------------------------------------- tokenizer
class WhisperTokenizerLite:
    """whisper/tokenizer.py [upstream] without tiktoken: byte-level BPE ranks (token bytes -> id), the special tokens appended in
    the published order, decoding (a join of token bytes), the word splitting used by the word timestamps, and the BPE *encoding*
    of short strings (needed for the non-speech / blank suppression lists only)."""

    def __init__(self, token_bytes: Sequence[bytes], multilingual: bool = True, num_languages: int = 99, language: Optional[str] = "en",
                 task: Optional[str] = "transcribe"):
        self.token_bytes = list(token_bytes)
        self.ranks = {b: i for i, b in enumerate(self.token_bytes)}
        n = len(self.token_bytes)
        langs = LANGUAGES[:num_languages]
        specials = ["<|endoftext|>", "<|startoftranscript|>", *[f"<|{l}|>" for l in langs], "<|translate|>", "<|transcribe|>", "<|startoflm|>",
                    "<|startofprev|>", "<|nospeech|>", "<|notimestamps|>", *[f"<|{i * 0.02:.2f}|>" for i in range(1501)]]
        self.special = {s: n + i for i, s in enumerate(specials)}
        self.special_by_id = {v: k for k, v in self.special.items()}
        self.n_vocab = n + len(specials)
        self.eot, self.sot = self.special["<|endoftext|>"], self.special["<|startoftranscript|>"]
        self.translate, self.transcribe = self.special["<|translate|>"], self.special["<|transcribe|>"]
        self.sot_lm, self.sot_prev = self.special["<|startoflm|>"], self.special["<|startofprev|>"]
        self.no_speech, self.no_timestamps = self.special["<|nospeech|>"], self.special["<|notimestamps|>"]
        self.timestamp_begin = self.special["<|0.00|>"]
        self.multilingual = multilingual
        self.language = (language or "en") if multilingual else None
        self.task = (task or "transcribe") if multilingual else None
        seq = [self.sot]
        if self.language is not None:
            seq.append(self.sot + 1 + list(langs).index(self.language))
        if self.task is not None:
            seq.append(self.transcribe if self.task == "transcribe" else self.translate)
        self.sot_sequence = tuple(seq)

    @classmethod
    def from_tiktoken_file(cls, path: str, **kw) -> "WhisperTokenizerLite":
        """whisper/assets/{multilingual,gpt2}.tiktoken: one ``base64(token bytes) rank`` pair per line."""
        pairs = [line.split() for line in open(path) if line.strip()]
        toks = [b""] * len(pairs)
        for tok, rank in pairs:
            toks[int(rank)] = base64.b64decode(tok)
        return cls(toks, **kw)

    # -- decoding
    def decode(self, tokens: Sequence[int]) -> str:
        out = []
        for t in tokens:
            t = int(t)
            if t >= self.timestamp_begin:
                continue
            out.append(self.token_bytes[t] if t < len(self.token_bytes) else self.special_by_id[t].encode())
        return b"".join(out).decode("utf-8", errors="replace")

    def decode_with_timestamps(self, tokens: Sequence[int]) -> str:
        out = []
        for t in tokens:
            t = int(t)
            if t >= self.timestamp_begin:
                out.append(f"<|{(t - self.timestamp_begin) * 0.02:.2f}|>".encode())
            else:
                out.append(self.token_bytes[t] if t < len(self.token_bytes) else self.special_by_id[t].encode())
        return b"".join(out).decode("utf-8", errors="replace")

    # -- encoding of short strings (one pre-tokenizer piece): merge the adjacent pair of lowest rank until none is left
    def encode_piece(self, text: str) -> List[int]:
        parts = [bytes([b]) for b in text.encode("utf-8")]
        while len(parts) > 1:
            best, best_rank = None, None
            for i in range(len(parts) - 1):
                r = self.ranks.get(parts[i] + parts[i + 1])
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = i, r
            if best is None:
                break
            parts[best:best + 2] = [parts[best] + parts[best + 1]]
        return [self.ranks[p] for p in parts if p in self.ranks]

    @property
    def non_speech_tokens(self) -> Tuple[int, ...]:
        symbols = list('"#()*+/:;<=>@[\\]^_`{|}~「」『』')
        symbols += "<< >> <<< >>> -- --- -( -[ (' (\" (( )) ((( ))) [[ ]] {{ }} ♪♪ ♪♪♪".split()
        miscellaneous = set("♩♪♫♬♭♮♯")
        result = set()
        for s0 in (" -", " '"):
            e = self.encode_piece(s0)
            if e:
                result.add(e[0])
        for symbol in symbols + list(miscellaneous):
            for toks in (self.encode_piece(symbol), self.encode_piece(" " + symbol)):
                if toks and (len(toks) == 1 or symbol in miscellaneous):
                    result.add(toks[0])
        return tuple(sorted(result))

    # -- word splitting
    def split_tokens_on_unicode(self, tokens: Sequence[int]):
        decoded_full = self.decode_with_timestamps(tokens)
        replacement = "�"
        words, word_tokens, current, offset = [], [], [], 0
        for t in tokens:
            current.append(int(t))
            decoded = self.decode_with_timestamps(current)
            if replacement not in decoded or decoded_full[offset + decoded.index(replacement)] == replacement:
                words.append(decoded)
                word_tokens.append(current)
                current = []
                offset += len(decoded)
        return words, word_tokens

    def split_tokens_on_spaces(self, tokens: Sequence[int]):
        subwords, subword_tokens = self.split_tokens_on_unicode(tokens)
        words, word_tokens = [], []
        for sw, st in zip(subwords, subword_tokens):
            special = st[0] >= self.eot
            with_space = sw.startswith(" ")
            punctuation = sw.strip() in string.punctuation
            if special or with_space or punctuation or len(words) == 0:
                words.append(sw)
                word_tokens.append(list(st))
            else:
                words[-1] = words[-1] + sw
                word_tokens[-1].extend(st)
        return words, word_tokens

    def split_to_word_tokens(self, tokens: Sequence[int]):
        if self.language in {"zh", "ja", "th", "lo", "my", "yue"}:
            return self.split_tokens_on_unicode(tokens)
        return self.split_tokens_on_spaces(tokens)


# ------------------------------------------------------------------------------------------------------------------- decoding
@dataclass
class DecodingResult:
    tokens: List[int] = field(default_factory=list)
    text: str = ""
    avg_logprob: float = float("nan")
    no_speech_prob: float = float("nan")
    temperature: float = float("nan")
    compression_ratio: float = float("nan")


def compression_ratio(text: str) -> float:
    b = text.encode("utf-8")
    return len(b) / len(zlib.compress(b))


def suppress_lists(tok: WhisperTokenizerLite):
    """DecodingTask._get_suppress_tokens with suppress_tokens = '-1' (the default) and SuppressBlank's list [upstream]."""
    sup = set(tok.non_speech_tokens)
    sup.update([tok.transcribe, tok.translate, tok.sot, tok.sot_prev, tok.sot_lm, tok.no_speech])
    blank = tok.encode_piece(" ") + [tok.eot]
    return tuple(sorted(sup)), tuple(blank)


class WhisperDecoder:
    """DecodingTask.run for ONE 30-s window [upstream whisper/decoding.py]: beam search (temperature 0, beam_size > 1), greedy arg-max,
    or temperature sampling, on top of ``WhisperB200.step_logits`` with the timestamp / suppression rules applied on the device."""

    def __init__(self, model: WhisperB200, tok: WhisperTokenizerLite, seed: int = 0):
        self.model, self.tok = model, tok
        sup, blank = suppress_lists(tok)
        self.rules = WhisperRules(eot=tok.eot, timestamp_begin=tok.timestamp_begin, no_timestamps=tok.no_timestamps,
                                  max_initial_timestamp_index=50, suppress=sup, suppress_first=blank)
        self.n_ctx = model.dec_dims.n_text_ctx
        self.sample_len = self.n_ctx // 2
        self.seed = seed
        self.gen = torch.Generator(device=model.device).manual_seed(seed)

    def reseed(self):
        """The temperature-fallback sampler restarts from the same seed for every recording: a stream transcribes to the same
        result whatever was transcribed before it (upstream draws from torch's global generator)."""
        self.gen.manual_seed(self.seed)

    def initial_tokens(self, prompt: Optional[Sequence[int]]) -> List[int]:
        tokens = list(self.tok.sot_sequence)
        if prompt:
            tokens = [self.tok.sot_prev] + [int(t) for t in prompt][-(self.n_ctx // 2 - 1):] + tokens
        return tokens

    @torch.no_grad()
    def run(self, enc_bf16: torch.Tensor, temperature: float, beam_size: Optional[int], best_of: Optional[int], prompt) -> DecodingResult:
        """enc_bf16 int16 [1, 1500, d]: the encoder output of the window."""
        tok, model, dev = self.tok, self.model, self.model.device
        n_group = (beam_size or 1) if temperature == 0 else (best_of or 1)
        init = self.initial_tokens(prompt)
        sample_begin, sot_index = len(init), init.index(tok.sot)
        model.begin_sequences(enc_bf16.expand(n_group, -1, -1).contiguous())
        tokens = torch.tensor([init] * n_group, dtype=torch.int32, device=dev)
        sum_logprobs = torch.zeros(n_group, dtype=torch.float32, device=dev)
        no_speech_prob = float("nan")
        logits = None
        for p, t in enumerate(init):                                   # the prompt, one position at a time (keys / values are cached)
            logits = model.step_logits(tokens[:, p])
            if p == sot_index:
                no_speech_prob = float(torch.softmax(logits[0].float(), -1)[tok.no_speech])
        beam = _BeamSearch(n_group, tok.eot) if (temperature == 0 and n_group > 1) else None
        completed = False
        for i in range(self.sample_len):
            lg = apply_logit_rules(logits, tokens, sample_begin, self.rules)
            if beam is not None:
                tokens, src, completed = beam.update(tokens, lg, sum_logprobs)
                model.reorder_sequences(src)
            else:
                logprobs = torch.log_softmax(lg.float(), -1)
                if temperature == 0:
                    nxt = lg.argmax(-1)
                else:
                    nxt = torch.multinomial(torch.softmax(lg.float() / temperature, -1), 1, generator=self.gen)[:, 0]
                cur = logprobs.gather(1, nxt[:, None])[:, 0]
                alive = tokens[:, -1] != tok.eot
                sum_logprobs += cur * alive
                nxt = torch.where(alive, nxt, torch.full_like(nxt, tok.eot))
                tokens = torch.cat([tokens, nxt[:, None].to(torch.int32)], 1)
                completed = bool((tokens[:, -1] == tok.eot).all())
            if completed or tokens.shape[1] > self.n_ctx:
                break
            logits = model.step_logits(tokens[:, -1])
        if beam is not None:
            cands, lps = beam.finalize(tokens, sum_logprobs)
        else:
            cands, lps = [r.tolist() + [tok.eot] for r in tokens], sum_logprobs.tolist()
        seqs = []
        for c in cands:
            c = list(c[sample_begin:])
            seqs.append(c[:c.index(tok.eot)] if tok.eot in c else c)
        # MaximumLikelihoodRanker with length_penalty = None: the highest log-probability per token
        best = int(np.argmax([lp / max(len(s), 1) for lp, s in zip(lps, seqs)]))
        toks = seqs[best]
        text = tok.decode([t for t in toks if t < tok.eot]).strip()
        return DecodingResult(tokens=toks, text=text, avg_logprob=float(lps[best]) / (len(toks) + 1), no_speech_prob=no_speech_prob,
                              temperature=temperature, compression_ratio=compression_ratio(text) if text else 0.0)


class _BeamSearch:
    """whisper/decoding.py::BeamSearchDecoder [upstream] for one audio window: ``beam`` hypotheses, patience 1."""

    def __init__(self, beam: int, eot: int):
        self.beam, self.eot = beam, eot
        self.max_candidates = beam
        self.finished: Dict[tuple, float] = {}

    def update(self, tokens: torch.Tensor, logits: torch.Tensor, sum_logprobs: torch.Tensor):
        logprobs = torch.log_softmax(logits.float(), -1)
        top_lp, top_tok = logprobs.topk(self.beam + 1, dim=-1)
        top_lp, top_tok = top_lp.tolist(), top_tok.tolist()
        prefixes, sums = tokens.tolist(), sum_logprobs.tolist()
        scores, sources = {}, {}
        for j in range(self.beam):
            for lp, t in zip(top_lp[j], top_tok[j]):
                seq = tuple(prefixes[j] + [t])
                scores[seq] = sums[j] + lp
                sources[seq] = j
        nxt, src, new_sums, newly = [], [], [], {}
        for seq in sorted(scores, key=scores.get, reverse=True):
            if seq[-1] == self.eot:
                newly[seq] = scores[seq]
            else:
                new_sums.append(scores[seq])
                nxt.append(seq)
                src.append(sources[seq])
                if len(nxt) == self.beam:
                    break
        for seq in sorted(newly, key=newly.get, reverse=True):
            if len(self.finished) >= self.max_candidates:
                break
            self.finished[seq] = newly[seq]
        dev = tokens.device
        sum_logprobs.copy_(torch.tensor(new_sums, dtype=torch.float32, device=dev))
        completed = len(self.finished) >= self.max_candidates
        return torch.tensor(nxt, dtype=torch.int32, device=dev), torch.tensor(src, dtype=torch.int32, device=dev), completed

    def finalize(self, tokens: torch.Tensor, sum_logprobs: torch.Tensor):
        if len(self.finished) < self.beam:                              # not enough finished hypotheses: the best unfinished ones + eot
            sums = sum_logprobs.cpu().numpy()
            rows = tokens.tolist()
            for i in np.argsort(sums)[::-1]:
                seq = tuple(rows[int(i)] + [self.eot])
                self.finished[seq] = float(sums[int(i)])
                if len(self.finished) >= self.beam:
                    break
        return [list(s) for s in self.finished.keys()], list(self.finished.values())


# ------------------------------------------------------------------------------------------------------------------- word timestamps
@dataclass
class WordTiming:
    word: str
    tokens: List[int]
    start: float
    end: float
    probability: float


def merge_punctuations(alignment: List[WordTiming], prepended: str, appended: str):
    """whisper/timing.py::merge_punctuations [upstream]."""
    i, j = len(alignment) - 2, len(alignment) - 1
    while i >= 0:
        previous, following = alignment[i], alignment[j]
        if previous.word.startswith(" ") and previous.word.strip() in prepended:
            following.word = previous.word + following.word
            following.tokens = previous.tokens + following.tokens
            previous.word, previous.tokens = "", []
        else:
            j = i
        i -= 1
    i, j = 0, 1
    while j < len(alignment):
        previous, following = alignment[i], alignment[j]
        if not previous.word.endswith(" ") and following.word in appended:
            previous.word = previous.word + following.word
            previous.tokens = previous.tokens + following.tokens
            following.word, following.tokens = "", []
        else:
            i = j
        j += 1


def words_from_alignment(tok: WhisperTokenizerLite, text_tokens: Sequence[int], start_frames: np.ndarray, token_probs: Sequence[float]) -> List[WordTiming]:
    """The tail of whisper/timing.py::find_alignment [upstream]: start_frames[i] = audio position (20 ms units) at which the DTW path
    enters the row that predicts text token i (row len(text_tokens): the eot row) -> words with start / end seconds."""
    words, word_tokens = tok.split_to_word_tokens(list(text_tokens) + [tok.eot])
    if len(word_tokens) <= 1:
        return []
    bounds = np.pad(np.cumsum([len(t) for t in word_tokens[:-1]]), (1, 0))
    jump_times = np.asarray(start_frames, dtype=np.float64) / TOKENS_PER_SECOND
    starts, ends = jump_times[bounds[:-1]], jump_times[bounds[1:]]
    probs = [float(np.mean(token_probs[i:j])) if j > i else 0.0 for i, j in zip(bounds[:-1], bounds[1:])]
    return [WordTiming(w, list(t), float(s), float(e), p) for w, t, s, e, p in zip(words, word_tokens, starts, ends, probs)]


# ------------------------------------------------------------------------------------------------------------------- transcribe
class WhisperB200Transcriber:
    """``(stream, cfg, options) -> whisper result dict``: the callable ``asr.set_transcriber`` expects.  ``stream`` is a WAV path or
    an int16 / float32 CUDA tensor (the device-resident separated stream of the CSS stage)."""

    PREPEND, APPEND = "\"'“¿([{-", "\"'.。,，!！?？:：”)]}、"

    def __init__(self, model: WhisperB200, tokenizer: WhisperTokenizerLite, alignment_heads: Optional[Sequence[Tuple[int, int]]] = None,
                 temperatures: Sequence[float] = (0.0, 0.2, 0.4, 0.6, 0.8, 1.0), compression_ratio_threshold: float = 2.4,
                 logprob_threshold: float = -1.0, no_speech_threshold: float = 0.6, condition_on_previous_text: bool = True):
        self.model, self.tok = model, tokenizer
        D = model.dec_dims
        # default of whisper/model.py: all heads of the last half of the decoder layers [upstream]; released checkpoints ship their own list
        self.alignment_heads = list(alignment_heads) if alignment_heads is not None else \
            [(l, h) for l in range(D.n_layers // 2, D.n_layers) for h in range(D.n_heads)]
        self.temperatures = tuple(temperatures)
        self.cr_th, self.lp_th, self.ns_th = compression_ratio_threshold, logprob_threshold, no_speech_threshold
        self.condition_on_previous_text = condition_on_previous_text
        self.decoder = WhisperDecoder(model, tokenizer)

    # -- input
    def _audio(self, stream) -> torch.Tensor:
        dev = self.model.device
        if isinstance(stream, torch.Tensor):
            a = stream.to(dev)
            return (a.to(torch.float32) / 32768.0) if a.dtype == torch.int16 else a.to(torch.float32)      # whisper.load_audio: int16 / 32768
        import scipy.io.wavfile as wf
        from .css import flush_wav_writes
        flush_wav_writes([stream])
        sr, data = wf.read(str(stream))
        if sr != SAMPLE_RATE:
            raise _cabi.NsfError(f"{stream}: {sr} Hz; the separated streams are 16 kHz (whisper.load_audio resamples with ffmpeg: not rebuilt)")
        if data.ndim > 1:
            data = data.mean(axis=1)
        t = torch.from_numpy(np.ascontiguousarray(data)).to(dev)
        return t.to(torch.float32) / 32768.0 if data.dtype == np.int16 else t.to(torch.float32)

    def decode_with_fallback(self, enc16: torch.Tensor, options: dict, prompt) -> DecodingResult:
        result = None
        for t in self.temperatures:
            beam = options.get("beam_size") if t == 0 else None          # beam search only at temperature 0, best_of only above
            best_of = options.get("best_of") if t > 0 else None
            result = self.decoder.run(enc16, t, beam, best_of, prompt)
            needs_fallback = False
            if self.cr_th is not None and result.compression_ratio > self.cr_th:
                needs_fallback = True                                    # too repetitive
            if self.lp_th is not None and result.avg_logprob < self.lp_th:
                needs_fallback = True                                    # average log probability is too low
            if self.ns_th is not None and result.no_speech_prob > self.ns_th and self.lp_th is not None and result.avg_logprob < self.lp_th:
                needs_fallback = False                                   # silence
            if not needs_fallback:
                break
        return result

    def _add_word_timestamps(self, segments: List[dict], enc16: torch.Tensor, num_frames: int, last_speech_timestamp: float) -> float:
        """whisper/timing.py::add_word_timestamps [upstream]."""
        tok, model = self.tok, self.model
        if not segments:
            return last_speech_timestamp
        per_seg = [[t for t in s["tokens"] if t < tok.eot] for s in segments]
        text_tokens = [t for ts in per_seg for t in ts]
        alignment: List[WordTiming] = []
        if text_tokens:
            prompt = list(tok.sot_sequence) + [tok.no_timestamps]
            forced = torch.tensor([text_tokens + [tok.eot]], dtype=torch.int32, device=model.device)
            n0 = len(tok.sot_sequence)
            # teacher-forced pass 1 (graph-replayed step): cross-attention rows of the alignment heads.  The sequence is padded with
            # eot to a multiple of 32 positions so that a handful of captured graphs serve every window (the extra rows are ignored)
            n_new = len(text_tokens) + 1
            n_pad = min(-(-(len(prompt) + n_new) // 32) * 32, model.dec_dims.n_text_ctx) - len(prompt)
            if n_pad > n_new:
                forced = torch.cat([forced, torch.full((1, n_pad - n_new), tok.eot, dtype=torch.int32, device=model.device)], 1)
            _, probs = model.decode_greedy(enc16, prompt, max_new_tokens=max(n_pad, n_new), forced_tokens=forced, align_heads=self.alignment_heads)
            mv = max(1, num_frames // 2)
            rows = probs[:, :, n0:n0 + len(text_tokens) + 1, :mv]
            rows = rows / rows.sum(-1, keepdim=True).clamp_min(1e-30)                      # softmax over the cropped positions [upstream]
            start = token_alignment(rows.contiguous())[0].cpu().numpy()
            # teacher-forced pass 2: probability of every text token given its prefix (softmax over the text vocabulary)
            model.begin_sequences(enc16)
            seq = prompt + text_tokens
            tprobs = []
            for p, t in enumerate(seq):
                lg = model.step_logits(torch.tensor([t], dtype=torch.int32, device=model.device))
                k = p - (len(prompt) - 1)
                if 0 <= k < len(text_tokens):
                    tprobs.append(float(torch.softmax(lg[0, :tok.eot].float(), -1)[text_tokens[k]]))
            alignment = words_from_alignment(tok, text_tokens, start, tprobs)
        durations = np.array([w.end - w.start for w in alignment])
        durations = durations[durations.nonzero()]
        median_duration = min(0.7, float(np.median(durations))) if len(durations) > 0 else 0.0
        max_duration = median_duration * 2
        if len(durations) > 0:                                           # truncate long words at sentence boundaries
            marks = ".。!！?？"
            for i in range(1, len(alignment)):
                if alignment[i].end - alignment[i].start > max_duration:
                    if alignment[i].word in marks:
                        alignment[i].end = alignment[i].start + max_duration
                    elif alignment[i - 1].word in marks:
                        alignment[i].start = alignment[i].end - max_duration
        merge_punctuations(alignment, self.PREPEND, self.APPEND)
        time_offset = segments[0]["seek"] * HOP / SAMPLE_RATE
        wi = 0
        for seg, toks in zip(segments, per_seg):
            saved, words = 0, []
            while wi < len(alignment) and saved < len(toks):
                w = alignment[wi]
                if w.word:
                    words.append(dict(word=w.word, start=round(time_offset + w.start, 2), end=round(time_offset + w.end, 2), probability=w.probability))
                saved += len(w.tokens)
                wi += 1
            if words:
                if words[0]["end"] - last_speech_timestamp > median_duration * 4 and (
                        words[0]["end"] - words[0]["start"] > max_duration or (len(words) > 1 and words[1]["end"] - words[0]["start"] > max_duration * 2)):
                    if len(words) > 1 and words[1]["end"] - words[1]["start"] > max_duration:
                        boundary = max(words[1]["end"] / 2, words[1]["end"] - max_duration)
                        words[0]["end"] = words[1]["start"] = boundary
                    words[0]["start"] = max(0, words[0]["end"] - max_duration)
                if seg["start"] < words[0]["end"] and seg["start"] - 0.5 > words[0]["start"]:
                    words[0]["start"] = max(0, min(words[0]["end"] - median_duration, seg["start"]))
                else:
                    seg["start"] = words[0]["start"]
                if seg["end"] > words[-1]["start"] and seg["end"] + 0.5 < words[-1]["end"]:
                    words[-1]["end"] = max(words[-1]["start"] + median_duration, seg["end"])
                else:
                    seg["end"] = words[-1]["end"]
                last_speech_timestamp = seg["end"]
            seg["words"] = words
        return last_speech_timestamp

    @torch.no_grad()
    def transcribe(self, stream, options: dict) -> dict:
        """whisper/transcribe.py::transcribe [upstream] for one stream."""
        tok, model = self.tok, self.model
        audio = self._audio(stream)
        if hasattr(self.decoder, "reseed"):
            self.decoder.reseed()
        log_spec, gmax, content_frames = model.log_mel_recording(audio.contiguous())
        content_duration = content_frames * HOP / SAMPLE_RATE
        word_timestamps = bool(options.get("word_timestamps", False))
        hst = options.get("hallucination_silence_threshold")
        time_precision = 0.02
        seek, all_tokens, all_segments, prompt_reset_since, last_speech_timestamp = 0, [], [], 0, 0.0
        punctuation = "\"'“¿([{-\"'.。,，!！?？:：”)]}、"

        def word_anomaly_score(w):
            probability, duration, score = w.get("probability", 0.0), w["end"] - w["start"], 0.0
            if probability < 0.15:
                score += 1.0
            if duration < 0.133:
                score += (0.133 - duration) * 15
            if duration > 2.0:
                score += duration - 2.0
            return score

        def is_segment_anomaly(s):
            if s is None or not s["words"]:
                return False
            words = [w for w in s["words"] if w["word"] not in punctuation][:8]
            score = sum(word_anomaly_score(w) for w in words)
            return score >= 3 or score + 0.01 >= len(words)

        def next_words_segment(segs):
            return next((s for s in segs if s["words"]), None)

        def get_end(segs):
            return next((w["end"] for s in reversed(segs) for w in reversed(s["words"])), segs[-1]["end"] if segs else None)

        while seek < content_frames:
            time_offset = seek * HOP / SAMPLE_RATE
            window_end_time = (seek + N_FRAMES) * HOP / SAMPLE_RATE
            segment_size = min(N_FRAMES, content_frames - seek)
            segment_duration = segment_size * HOP / SAMPLE_RATE
            hi, lo = model.mel_windows(log_spec, gmax, [seek], [segment_size])
            _, enc16 = model.encode(hi, lo)
            prompt = all_tokens[prompt_reset_since:] if self.condition_on_previous_text else None
            result = self.decode_with_fallback(enc16, options, prompt)
            tokens = list(result.tokens)
            if self.ns_th is not None:
                should_skip = result.no_speech_prob > self.ns_th
                if self.lp_th is not None and result.avg_logprob > self.lp_th:
                    should_skip = False                                  # the log-probability is high enough despite the no-speech probability
                if should_skip:
                    seek += segment_size
                    continue
            previous_seek = seek
            current: List[dict] = []

            def new_segment(start, end, toks):
                return dict(seek=previous_seek, start=start, end=end, text=tok.decode([t for t in toks if t < tok.eot]), tokens=list(toks),
                            temperature=result.temperature, avg_logprob=result.avg_logprob, compression_ratio=result.compression_ratio,
                            no_speech_prob=result.no_speech_prob)

            is_ts = [t >= tok.timestamp_begin for t in tokens]
            single_timestamp_ending = is_ts[-2:] == [False, True]
            consecutive = [i + 1 for i in range(len(tokens) - 1) if is_ts[i] and is_ts[i + 1]]
            if consecutive:
                slices = list(consecutive)
                if single_timestamp_ending:
                    slices.append(len(tokens))
                last = 0
                for cur in slices:
                    sl = tokens[last:cur]
                    current.append(new_segment(time_offset + (sl[0] - tok.timestamp_begin) * time_precision,
                                               time_offset + (sl[-1] - tok.timestamp_begin) * time_precision, sl))
                    last = cur
                if single_timestamp_ending:
                    seek += segment_size                                 # no speech after the last timestamp
                else:
                    seek += (tokens[last - 1] - tok.timestamp_begin) * 2
            else:
                duration = segment_duration
                ts = [t for t in tokens if t >= tok.timestamp_begin]
                if ts and ts[-1] != tok.timestamp_begin:
                    duration = (ts[-1] - tok.timestamp_begin) * time_precision
                current.append(new_segment(time_offset, time_offset + duration, tokens))
                seek += segment_size

            if word_timestamps:
                self._add_word_timestamps(current, enc16, segment_size, last_speech_timestamp)
                if not single_timestamp_ending:
                    lwe = get_end(current)
                    if lwe is not None and lwe > time_offset:
                        seek = round(lwe * FRAMES_PER_SECOND)
                if hst is not None:                                       # skip silence before possible hallucinations
                    if not single_timestamp_ending:
                        lwe = get_end(current)
                        if lwe is not None and lwe > time_offset:
                            seek = round(lwe * FRAMES_PER_SECOND) if window_end_time - lwe > hst else previous_seek + segment_size
                    first = next_words_segment(current)
                    if first is not None and is_segment_anomaly(first):
                        gap = first["start"] - time_offset
                        if gap > hst:
                            seek = previous_seek + round(gap * FRAMES_PER_SECOND)
                            continue
                    hal_last_end = last_speech_timestamp
                    for si in range(len(current)):
                        s = current[si]
                        if not s["words"]:
                            continue
                        if is_segment_anomaly(s):
                            nxt = next_words_segment(current[si + 1:])
                            hal_next_start = nxt["words"][0]["start"] if nxt is not None else time_offset + segment_duration
                            silence_before = s["start"] - hal_last_end > hst or s["start"] < hst or s["start"] - time_offset < 2.0
                            silence_after = hal_next_start - s["end"] > hst or is_segment_anomaly(nxt) or window_end_time - s["end"] < 2.0
                            if silence_before and silence_after:
                                seek = round(max(time_offset + 1, s["start"]) * FRAMES_PER_SECOND)
                                if content_duration - s["end"] < hst:
                                    seek = content_frames
                                current[si:] = []
                                break
                        hal_last_end = s["end"]
                lwe = get_end(current)
                if lwe is not None:
                    last_speech_timestamp = lwe

            for s in current:                                            # instantaneous or empty segments carry nothing
                if s["start"] == s["end"] or s["text"].strip() == "":
                    s["text"], s["tokens"], s["words"] = "", [], []
            all_segments.extend({"id": i, **s} for i, s in enumerate(current, start=len(all_segments)))
            all_tokens.extend(t for s in current for t in s["tokens"])
            if not self.condition_on_previous_text or result.temperature > 0.5:
                prompt_reset_since = len(all_tokens)                     # no prompt from a high-temperature window
            if seek <= previous_seek:                                    # the seek must advance (a window that closes at <|0.00|>)
                seek = previous_seek + segment_size
        if word_timestamps:
            for s in all_segments:
                s.setdefault("words", [])
        return dict(text=tok.decode([t for t in all_tokens if t < tok.eot]), segments=all_segments, language=tok.language)

    def __call__(self, stream, cfg, options: dict) -> dict:
        return self.transcribe(stream, options)


def transcriber_from_env() -> Optional[WhisperB200Transcriber]:
    """NSF_WHISPER_CKPT (an openai-whisper ``.pt`` with 'dims' / 'model_state_dict', or a plain state dict in either naming) and
    NSF_WHISPER_VOCAB (the matching ``.tiktoken`` file) -> a transcriber on the current CUDA device, else None."""
    ckpt, vocab = os.environ.get("NSF_WHISPER_CKPT"), os.environ.get("NSF_WHISPER_VOCAB")
    if not ckpt or not vocab:
        return None
    obj = torch.load(ckpt, map_location="cpu", weights_only=False)
    sd = obj.get("model_state_dict", obj) if isinstance(obj, dict) else obj
    dims = obj.get("dims", {}) if isinstance(obj, dict) else {}
    n_vocab = int(dims.get("n_vocab", 51865))
    multilingual = n_vocab >= 51865
    tok = WhisperTokenizerLite.from_tiktoken_file(vocab, multilingual=multilingual, num_languages=n_vocab - 51765 - int(multilingual),
                                                  language=os.environ.get("NSF_WHISPER_LANGUAGE", "en"))
    heads = None
    if "alignment_heads" in sd:                                          # a dense [n_layers, n_heads] mask in released checkpoints
        m = sd.pop("alignment_heads")
        m = m.to_dense() if hasattr(m, "to_dense") else m
        heads = [(int(l), int(h)) for l, h in torch.nonzero(torch.as_tensor(m)).tolist()]
    return WhisperB200Transcriber(WhisperB200(sd), tok, alignment_heads=heads)
