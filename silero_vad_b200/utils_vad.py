"""Host-side utilities with the reference's signatures (src/silero_vad/utils_vad.py:211-655).

get_speech_timestamps: the chunk loop of the reference (utils_vad.py:323-336, one model call and one
`.item()` sync per 32 ms) becomes one fused bulk launch; the hysteresis automaton runs in C++
(csrc/svad_segments.cpp through the C ABI).  Any other duck-typed model (e.g. the reference's own
TorchScript object) is still accepted and driven chunk by chunk.
"""
import math
import warnings
from typing import Callable, List

import numpy as np
import torch

from . import _cabi
from .model import SileroVADB200

languages = ['ru', 'en', 'de', 'es']


def _segment_params(sampling_rate, threshold, neg_threshold, min_speech_duration_ms, max_speech_duration_s,
                    min_silence_duration_ms, speech_pad_ms, min_silence_at_max_speech, use_max_poss_sil_at_max_speech):
    p = _cabi.SegmentParams()
    _cabi.lib().svad_segment_params_default(p)
    p.sampling_rate = sampling_rate
    p.threshold = threshold
    p.neg_threshold = math.nan if neg_threshold is None else neg_threshold
    p.min_speech_duration_ms = min_speech_duration_ms
    p.max_speech_duration_s = max_speech_duration_s
    p.min_silence_duration_ms = min_silence_duration_ms
    p.speech_pad_ms = speech_pad_ms
    p.min_silence_at_max_speech_ms = min_silence_at_max_speech
    p.use_max_poss_sil_at_max_speech = 1 if use_max_poss_sil_at_max_speech else 0
    return p


def _finish(segs, sampling_rate, audio_length_samples, return_seconds, time_resolution, step):
    speeches = [{'start': a, 'end': b} for a, b in segs]
    if return_seconds:
        audio_length_seconds = audio_length_samples / sampling_rate
        for d in speeches:
            d['start'] = max(round(d['start'] / sampling_rate, time_resolution), 0)
            d['end'] = min(round(d['end'] / sampling_rate, time_resolution), audio_length_seconds)
    elif step > 1:
        for d in speeches:
            d['start'] *= step
            d['end'] *= step
    return speeches


def _probs_chunk_by_chunk(audio, model, sampling_rate, window, progress_tracking_callback):
    n = len(audio)
    probs = []
    for start in range(0, n, window):
        chunk = audio[start:start + window]
        if len(chunk) < window:
            chunk = torch.nn.functional.pad(chunk, (0, int(window - len(chunk))))
        probs.append(model(chunk, sampling_rate).item())
        if progress_tracking_callback:
            progress_tracking_callback(min(start + window, n) / n * 100)
    return np.asarray(probs, np.float32)


@torch.no_grad()
def get_speech_timestamps(audio: torch.Tensor,
                          model,
                          threshold: float = 0.5,
                          sampling_rate: int = 16000,
                          min_speech_duration_ms: int = 250,
                          max_speech_duration_s: float = float('inf'),
                          min_silence_duration_ms: int = 100,
                          speech_pad_ms: int = 30,
                          return_seconds: bool = False,
                          time_resolution: int = 1,
                          visualize_probs: bool = False,
                          progress_tracking_callback: Callable[[float], None] = None,
                          neg_threshold: float = None,
                          window_size_samples: int = 512,
                          min_silence_at_max_speech: int = 98,
                          use_max_poss_sil_at_max_speech: bool = True):
    """Split a 1-D audio tensor into speech segments; parameters, defaults, units, warnings and errors as in
    the reference (utils_vad.py:212-455).  Returns a list of {'start', 'end'} dicts."""
    if not torch.is_tensor(audio):
        try:
            audio = torch.Tensor(audio)
        except Exception:
            raise TypeError("Audio cannot be casted to tensor. Cast it manually")
    if len(audio.shape) > 1:
        for _ in range(len(audio.shape)):
            audio = audio.squeeze(0)
        if len(audio.shape) > 1:
            raise ValueError("More than one dimension in audio. Are you trying to process audio with 2 channels?")
    fused = isinstance(model, SileroVADB200)
    raw_audio, raw_rate = audio, sampling_rate
    if sampling_rate > 16000 and (sampling_rate % 16000 == 0):
        step = sampling_rate // 16000
        sampling_rate = 16000
        if not fused:
            audio = audio[::step]
        warnings.warn('Sampling rate is a multiply of 16000, casting to 16000 manually!')
    else:
        step = 1
    if sampling_rate not in [8000, 16000]:
        raise ValueError("Currently silero VAD models support 8000 and 16000 (or multiply of 16000) sample rates")
    window = 512 if sampling_rate == 16000 else 256
    audio_length_samples = -(-len(raw_audio) // step) if fused else len(audio)    # = len(audio[::step])

    model.reset_states()
    if fused and audio_length_samples > 0:
        # the kernel decimates (reads every step-th sample) and zero-pads the tail itself; a clip shorter than one window is
        # padded up to it here, as the reference pads every chunk before the model call (utils_vad.py:323-327)
        x = raw_audio
        if audio_length_samples < window:
            x = torch.nn.functional.pad(raw_audio, (0, window * step - len(raw_audio)))
        probs = model.audio_forward(x.unsqueeze(0), raw_rate)[0].numpy()
        if progress_tracking_callback:
            for start in range(0, audio_length_samples, window):
                progress_tracking_callback(min(start + window, audio_length_samples) / audio_length_samples * 100)
    else:
        probs = _probs_chunk_by_chunk(audio, model, sampling_rate, window, progress_tracking_callback)

    p = _segment_params(sampling_rate, threshold, neg_threshold, min_speech_duration_ms, max_speech_duration_s,
                        min_silence_duration_ms, speech_pad_ms, min_silence_at_max_speech, use_max_poss_sil_at_max_speech)
    segs = _cabi.speech_segments(probs.reshape(1, -1), np.asarray([audio_length_samples], np.int64), p)[0]
    speeches = _finish(segs, sampling_rate, audio_length_samples, return_seconds, time_resolution, step)
    if visualize_probs:
        make_visualization(probs.tolist(), window / sampling_rate)
    return speeches


@torch.no_grad()
def get_speech_timestamps_batch(audio, model, lengths=None, sampling_rate: int = 16000, threshold: float = 0.5,
                                min_speech_duration_ms: int = 250, max_speech_duration_s: float = float('inf'),
                                min_silence_duration_ms: int = 100, speech_pad_ms: int = 30, return_seconds: bool = False,
                                time_resolution: int = 1, neg_threshold: float = None, min_silence_at_max_speech: int = 98,
                                use_max_poss_sil_at_max_speech: bool = True):
    """B independent streams at once: audio f32[B, L] (rows zero-padded to a common L), lengths[B] the true
    sample counts.  One fused launch for all probabilities, one C++ pass for all segment lists.  Each
    returned list equals get_speech_timestamps() on that row alone (rows are independent in the reference:
    utils_vad.py:65-76)."""
    if sampling_rate not in [8000, 16000]:
        raise ValueError("Currently silero VAD models support 8000 and 16000 (or multiply of 16000) sample rates")
    audio = torch.as_tensor(audio)
    if audio.dim() != 2:
        raise ValueError("audio must be [B, L]")
    B, L = audio.shape
    lengths = np.full(B, L, np.int64) if lengths is None else np.asarray(lengths, np.int64)
    window = 512 if sampling_rate == 16000 else 256
    if L < window:   # clips shorter than one window: padded like the reference pads every chunk (utils_vad.py:323-327)
        audio = torch.nn.functional.pad(audio, (0, window - L))
    probs = model.audio_forward(audio, sampling_rate).numpy()
    p = _segment_params(sampling_rate, threshold, neg_threshold, min_speech_duration_ms, max_speech_duration_s,
                        min_silence_duration_ms, speech_pad_ms, min_silence_at_max_speech, use_max_poss_sil_at_max_speech)
    segs = _cabi.speech_segments(probs, lengths, p)
    return [_finish(s, sampling_rate, int(n), return_seconds, time_resolution, 1) for s, n in zip(segs, lengths)]


class VADIterator:
    """Streaming start/end event emitter, one chunk per call (utils_vad.py:458-549)."""

    def __init__(self, model, threshold: float = 0.5, sampling_rate: int = 16000, min_silence_duration_ms: int = 100,
                 speech_pad_ms: int = 30):
        self.model = model
        self.threshold = threshold
        self.sampling_rate = sampling_rate
        if sampling_rate not in [8000, 16000]:
            raise ValueError('VADIterator does not support sampling rates other than [8000, 16000]')
        self.min_silence_samples = sampling_rate * min_silence_duration_ms / 1000
        self.speech_pad_samples = sampling_rate * speech_pad_ms / 1000
        self.reset_states()

    def reset_states(self):
        self.model.reset_states()
        self.triggered = False
        self.temp_end = 0
        self.current_sample = 0

    @torch.no_grad()
    def __call__(self, x, return_seconds=False, time_resolution: int = 1):
        if not torch.is_tensor(x):
            try:
                x = torch.Tensor(x)
            except Exception:
                raise TypeError("Audio cannot be casted to tensor. Cast it manually")
        window = len(x[0]) if x.dim() == 2 else len(x)
        self.current_sample += window
        prob = self.model(x, self.sampling_rate).item()
        rising = prob >= self.threshold
        if rising and self.temp_end:
            self.temp_end = 0
        if rising and not self.triggered:
            self.triggered = True
            start = max(0, self.current_sample - self.speech_pad_samples - window)
            return {'start': int(start) if not return_seconds else round(start / self.sampling_rate, time_resolution)}
        if prob < self.threshold - 0.15 and self.triggered:
            if not self.temp_end:
                self.temp_end = self.current_sample
            if self.current_sample - self.temp_end < self.min_silence_samples:
                return None
            end = self.temp_end + self.speech_pad_samples - window
            self.temp_end = 0
            self.triggered = False
            return {'end': int(end) if not return_seconds else round(end / self.sampling_rate, time_resolution)}
        return None


class VADIteratorBatch:
    """B live streams multiplexed through one model call per chunk period (the telephony case: README.md:132 of the
    reference).  Each row behaves exactly like its own `VADIterator` (utils_vad.py:458-549):
    `__call__(x[B, n])` returns a list of B entries, each None / {'start': ..} / {'end': ..}.  Rows can be parked and
    resumed with `model.get_states()` / `model.set_states()`; `reset_rows(mask)` restarts individual streams."""

    def __init__(self, model, batch_size: int, threshold: float = 0.5, sampling_rate: int = 16000,
                 min_silence_duration_ms: int = 100, speech_pad_ms: int = 30):
        if sampling_rate not in [8000, 16000]:
            raise ValueError('VADIterator does not support sampling rates other than [8000, 16000]')
        self.model, self.B = model, batch_size
        self.threshold, self.sampling_rate = threshold, sampling_rate
        self.min_silence_samples = sampling_rate * min_silence_duration_ms / 1000
        self.speech_pad_samples = sampling_rate * speech_pad_ms / 1000
        self.reset_states()

    def reset_states(self):
        self.model.reset_states()
        self.triggered = np.zeros(self.B, bool)
        self.temp_end = np.zeros(self.B, np.int64)
        self.current_sample = np.zeros(self.B, np.int64)

    def reset_rows(self, mask):
        """Restart the streams selected by the boolean mask (new call on that line): their LSTM state and audio context
        are zeroed on the device, their event automaton is cleared; the other rows are untouched."""
        mask = np.asarray(mask, bool)
        self.triggered[mask] = False
        self.temp_end[mask] = 0
        self.current_sample[mask] = 0
        st = getattr(self.model, "_state", None)
        if st is not None and mask.any():
            m = torch.as_tensor(mask, device=st.device)
            st[:, m, :] = 0
            self.model._context[m, :] = 0

    @torch.no_grad()
    def __call__(self, x, return_seconds=False, time_resolution: int = 1):
        x = torch.as_tensor(x)
        if x.dim() != 2 or x.shape[0] != self.B:
            raise ValueError(f"expected [{self.B}, n] audio chunks")
        window = x.shape[1]
        self.current_sample += window
        prob = self.model(x, self.sampling_rate).reshape(-1).cpu().numpy().astype(np.float64)
        rising = prob >= self.threshold
        self.temp_end[rising & (self.temp_end != 0)] = 0
        out = [None] * self.B
        start_rows = rising & ~self.triggered
        for b in np.nonzero(start_rows)[0]:
            self.triggered[b] = True
            start = max(0, self.current_sample[b] - self.speech_pad_samples - window)
            out[b] = {'start': int(start) if not return_seconds else round(start / self.sampling_rate, time_resolution)}
        falling = (prob < self.threshold - 0.15) & self.triggered & ~start_rows
        for b in np.nonzero(falling)[0]:
            if not self.temp_end[b]:
                self.temp_end[b] = self.current_sample[b]
            if self.current_sample[b] - self.temp_end[b] < self.min_silence_samples:
                continue
            end = self.temp_end[b] + self.speech_pad_samples - window
            self.temp_end[b] = 0
            self.triggered[b] = False
            out[b] = {'end': int(end) if not return_seconds else round(end / self.sampling_rate, time_resolution)}
        return out


def _gather_chunks(tss_rows, wav2d, lengths, drop, model=None):
    """One gather launch (svad_collect_chunks_device) over the segment table of B rows.  wav2d: [B, L] float32 or int16 (CPU
    tensors are moved to the model's GPU).  Returns (flat device tensor, out_offsets[B+1])."""
    eng_model = model if model is not None else _default_model(wav2d)
    dev = eng_model.device
    x = wav2d.to(device=dev).contiguous()
    if x.dtype not in (torch.float32, torch.int16):
        x = x.to(torch.float32)
    B, L = x.shape
    rows = np.asarray([b for b, tss in enumerate(tss_rows) for _ in tss], np.int64)
    bounds = np.asarray([[int(d['start']), int(d['end'])] for tss in tss_rows for d in tss], np.int64).reshape(-1, 2)
    lengths = np.full(B, L, np.int64) if lengths is None else np.asarray(lengths, np.int64)
    eng = eng_model.engine
    es = x.element_size()
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        offs = eng.collect_chunks_device(0, es, B, x.stride(0), lengths, rows, bounds, drop, 0, 0, st)   # sizing pass (host only)
        out = torch.empty(int(offs[-1]), dtype=x.dtype, device=dev)
        if out.numel():
            eng.collect_chunks_device(x.data_ptr(), es, B, x.stride(0), lengths, rows, bounds, drop, out.data_ptr(), out.numel(), st)
    return out, offs


_models_by_device = {}


def _default_model(t):
    """Engine to run the gather on when the caller has none at hand: one per CUDA device, created on first use."""
    idx = t.device.index if t.is_cuda else torch.cuda.current_device()
    if idx not in _models_by_device:
        _models_by_device[idx] = SileroVADB200(device=idx)
    return _models_by_device[idx]


def _chunks(tss, wav, seconds, sampling_rate, drop):
    if seconds and not sampling_rate:
        raise ValueError('sampling_rate must be provided when seconds is True')
    _tss = _seconds_to_samples_tss(tss, sampling_rate) if seconds else tss
    if not drop and len(_tss) == 0:
        return torch.cat([])                       # what the reference's torch.cat of an empty list does (raises)
    if wav.dim() != 1 or any(d['start'] < 0 or d['end'] < 0 for d in _tss):
        # negative (wrap-around) bounds or unusual shapes: plain slicing, nothing to accelerate
        if drop:
            chunks, cur = [], 0
            for i in _tss:
                chunks.append(wav[cur:i['start']])
                cur = i['end']
            chunks.append(wav[cur:])
            return torch.cat(chunks)
        return torch.cat([wav[i['start']:i['end']] for i in _tss])
    out, _ = _gather_chunks([_tss], wav.unsqueeze(0), None, drop)
    out = out.to(wav.dtype) if out.dtype != wav.dtype else out
    return out if wav.is_cuda else out.to(wav.device)


def collect_chunks(tss: List[dict], wav: torch.Tensor, seconds: bool = False, sampling_rate: int = None) -> torch.Tensor:
    """Concatenate the audio inside the given segments (utils_vad.py:552-597): one gather launch over the segment table;
    the result lives where `wav` lives (a CUDA `wav` never leaves the GPU)."""
    return _chunks(tss, wav, seconds, sampling_rate, drop=False)


def drop_chunks(tss: List[dict], wav: torch.Tensor, seconds: bool = False, sampling_rate: int = None) -> torch.Tensor:
    """Concatenate the audio outside the given segments (utils_vad.py:600-646), same mechanism."""
    return _chunks(tss, wav, seconds, sampling_rate, drop=True)


def collect_chunks_batch(tss_rows, wav: torch.Tensor, lengths=None, drop: bool = False, model=None):
    """B rows at once: tss_rows[b] = segment list of row b of wav[B, L] (float32 or int16 PCM).  Returns the list of B
    per-row results as views of ONE device buffer filled by ONE gather launch."""
    out, offs = _gather_chunks(tss_rows, wav, lengths, drop, model)
    return [out[int(offs[b]):int(offs[b + 1])] for b in range(len(tss_rows))]


def _seconds_to_samples_tss(tss: List[dict], sampling_rate: int) -> List[dict]:
    return [{'start': round(c['start'] * sampling_rate), 'end': round(c['end'] * sampling_rate)} for c in tss]


def read_audio(path: str, sampling_rate: int = 16000) -> torch.Tensor:
    """Mono float32 [-1, 1) tensor from a PCM WAV file (int16 / 32768, the convention of
    examples/cpp/wav.h:95-136 and examples/onnx_sequence/run.py:104-119).  The reference goes through
    torchaudio/torchcodec (utils_vad.py:138-169), which this image cannot do; other containers and
    resampling by non-integer factors are out of scope here."""
    import wave
    with wave.open(str(path), "rb") as w:
        if w.getsampwidth() != 2 or w.getcomptype() != "NONE":
            raise RuntimeError(f"{path}: only uncompressed PCM16 WAV is supported")
        sr, ch = w.getframerate(), w.getnchannels()
        pcm = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
    wav = torch.from_numpy(pcm.astype(np.float32) / 32768.0)
    if ch > 1:
        wav = wav.view(-1, ch).mean(dim=1)
    if sr != sampling_rate:
        if sr % sampling_rate == 0:
            wav = wav[::sr // sampling_rate]
        else:
            raise RuntimeError(f"{path}: {sr} Hz cannot be brought to {sampling_rate} Hz by integer decimation")
    return wav


def save_audio(path: str, tensor: torch.Tensor, sampling_rate: int = 16000):
    """16-bit PCM WAV writer (utils_vad.py:162-185 writes bits_per_sample=16 through torchaudio)."""
    import wave
    t = tensor.detach().cpu().flatten().clamp(-1.0, 1.0)
    pcm = (t * 32767.0).round().to(torch.int16).numpy()
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sampling_rate)
        w.writeframes(pcm.astype("<i2").tobytes())


def make_visualization(probs, step):
    import pandas as pd
    pd.DataFrame({'probs': probs}, index=[x * step for x in range(len(probs))]).plot(
        figsize=(16, 8), kind='area', ylim=[0, 1.05], xlim=[0, len(probs) * step], xlabel='seconds',
        ylabel='speech probability', colormap='tab20')


# ---- constructors of the reference's two model classes (utils_vad.py:10-31, 194-199), for callers that build the model
# themselves instead of going through load_silero_vad().  Both return the CUDA engine; the model file path is accepted and
# ignored (the weights are the ones of silero_vad.jit), and there is no CPU execution to force.
def init_jit_model(model_path: str = None, device=None):
    """Reference: utils_vad.py:194-199 (torch.jit.load + eval).  `device`: a CUDA device / index, default cuda:0."""
    del model_path
    if isinstance(device, torch.device):
        if device.type != "cuda":
            raise RuntimeError("silero_vad_b200 runs on a CUDA device only (no CPU fallback)")
        device = torch.cuda.current_device() if device.index is None else device.index
    return SileroVADB200(device=device)


class OnnxWrapper(SileroVADB200):
    """Reference: utils_vad.py:10-31.  Same call protocol as the TorchScript wrapper; `force_onnx_cpu` has no meaning here."""

    def __init__(self, path=None, force_onnx_cpu=False):
        del force_onnx_cpu
        super().__init__()
        if path is not None and "16k" in str(path):
            self.sample_rates = [16000]   # silero_vad_16k_op15.onnx is 16 kHz only (utils_vad.py:27-29)
