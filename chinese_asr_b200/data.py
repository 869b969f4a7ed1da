"""Feature front end: the reference's data.py hot-path surface on B200.

Same names and argument meaning as the reference (data.py:21-57 create_fb_matrix, :84-106 MelScale,
:109-121 fast_read, :167-280 get_log_mel, :371-382 AudioBase), but get_log_mel runs the fused
CUDA kernels (csrc/features.cu) through the C ABI and returns a CUDA tensor.  The host only builds
the constant tables (filterbank, window, delta taps) once, exactly as AudioBase does.

Batched front end: AudioDst / AudioLoader (data.py:385-540, eval and
infer modes) hand whole batches of 16-bit PCM to the device - the samples travel as int16 and the
log-mel kernel converts them, the instance normalisation of batch_audio (eps 1e-7, data.py:517) is
fused into the feature kernels."""
import math
import os
import pickle
import wave

import numpy as np
import torch

from .gpd import gpd


def create_fb_matrix(n_stft, f_min, f_max, n_mels):
    """Triangular mel filterbank [n_stft, n_mels] (data.py:21-57): HTK mel scale, un-normalised,
    with the reference's quirk that STFT bin j is assigned frequency linspace(f_min, f_max)[j]."""
    def to_mel(f):
        return 2595. * torch.log10(torch.tensor(1.) + (f / 700.))

    def to_hz(m):
        return 700. * (10 ** (m / 2595.) - 1.)

    freqs = torch.linspace(f_min, f_max, n_stft)
    m_lo = 0. if f_min == 0 else to_mel(f_min)
    pts = to_hz(torch.linspace(m_lo, to_mel(f_max), n_mels + 2))
    gap = pts[1:] - pts[:-1]
    delta = pts.unsqueeze(0) - freqs.unsqueeze(1)
    down = (-1. * delta[:, :-2]) / gap[:-1]
    up = delta[:, 2:] / gap[1:]
    return torch.max(torch.tensor(0.), torch.min(down, up))


class MelScale(object):
    """Holder of the filterbank (data.py:84-106).  Calling it on a spectrogram is not part of the
    B200 path (the mel projection is fused into the log-mel kernel); `fb` is what is consumed."""

    def __init__(self, n_mels=128, sr=16000, f_max=None, f_min=0., n_stft=None):
        self.n_mels = n_mels
        self.sr = sr
        self.f_max = f_max if f_max is not None else sr // 2
        self.f_min = f_min
        self.fb = create_fb_matrix(n_stft, self.f_min, self.f_max, self.n_mels) if n_stft is not None else None


def delta_filter_stack():
    """[3, 9] identity / delta / delta-delta taps, L2-normalised per filter (data.py:138-149)."""
    d = np.array([2, 1, 0, -1, -2])
    dd = np.convolve(d, d, "full")
    f = np.array([[0] * 4 + [1] + [0] * 4, [0] * 2 + list(d) + [0] * 2, list(dd)], dtype=np.float32)
    f /= np.sqrt(np.sum(f ** 2, axis=1, keepdims=True))
    return np.ascontiguousarray(f, dtype=np.float32)


def fast_read(path):
    """16-bit or 32-bit integer PCM WAV -> float32 in [-1, 1) (data.py:109-121; soundfile is replaced by
    the stdlib wave module, which reads integer PCM only: an IEEE-float WAV raises wave.Error)."""
    with wave.open(path, 'rb') as w:
        rate, width, ch, n = w.getframerate(), w.getsampwidth(), w.getnchannels(), w.getnframes()
        raw = w.readframes(n)
    if width == 2:
        data = np.frombuffer(raw, dtype='<i2').astype(np.float32) / 32768.0
    elif width == 4:
        data = np.frombuffer(raw, dtype='<i4').astype(np.float32) / 2147483648.0
    else:
        raise ValueError(f"unsupported sample width {width} in {path}")
    if ch > 1:
        data = data.reshape(-1, ch)[:, 0].copy()
    if rate != gpd['sample_rate']:
        print(f'[WARN] rate={rate}, dtype={data.dtype}, path={path}')
    return data


def read_pcm(path):
    """WAV samples as stored: int16 for 16-bit files (converted to float on the device, same values
    as fast_read's x / 32768), float32 otherwise.  Channel 0 of multi-channel files."""
    with wave.open(path, 'rb') as w:
        rate, width, ch, n = w.getframerate(), w.getsampwidth(), w.getnchannels(), w.getnframes()
        raw = w.readframes(n)
    if rate != gpd['sample_rate']:
        print(f'[WARN] rate={rate}, path={path}')
    if width == 2:
        data = np.frombuffer(raw, dtype='<i2')
        return data.reshape(-1, ch)[:, 0].copy() if ch > 1 else data
    return fast_read(path)


def convert_pcm(samples, rate, norm_db=-1.0):
    """The two steps of the reference's convert_audio (main.py:19-24: ffmpeg -ar 16000 -ac 1, sox --norm=-1) on
    the device: `samples` [n] or [n, channels], int16 or float32 in [-1, 1) at `rate` Hz -> 16 kHz mono int16 with
    its peak at norm_db dBFS.  Builder-defined filter (include/asr_b200.h: asr_convert_audio)."""
    import ctypes as C
    from . import _cabi
    a = np.asarray(samples)
    if a.ndim == 1:
        a = a[:, None]
    if a.dtype == np.int16:
        fmt = _cabi.PCM_S16
    elif np.issubdtype(a.dtype, np.floating):
        a, fmt = a.astype(np.float32), _cabi.PCM_F32
    else:
        raise TypeError(f"waveform dtype {a.dtype}: only float (in [-1, 1)) or int16 PCM is accepted")
    a = np.ascontiguousarray(a)
    n, ch = a.shape
    cap = int(_cabi.lib.asr_convert_audio_length(n, int(rate)))
    out = np.empty(max(cap, 1), dtype=np.int16)
    n_out = C.c_int64(0)
    _cabi.check(_cabi.lib.asr_convert_audio(a.ctypes.data_as(C.c_void_p), fmt, n, ch, int(rate), float(norm_db),
                                            out.ctypes.data_as(C.POINTER(C.c_int16)), cap, C.byref(n_out), None),
                "asr_convert_audio")
    return out[:n_out.value]


def convert_audio(path, out_path='tmp.wav'):
    """main.py:19-24: `path` -> 16 kHz mono 16-bit, peak-normalised to -1 dBFS, written to tmp.wav (the reference's
    fixed output name); returns that path.  Input: integer-PCM WAV of any rate / channel count - the reference
    hands `path` to ffmpeg, which also decodes compressed containers; that decoder is not part of this package."""
    with wave.open(path, 'rb') as w:
        rate, width, ch, n = w.getframerate(), w.getsampwidth(), w.getnchannels(), w.getnframes()
        raw = w.readframes(n)
    if width == 2:
        a = np.frombuffer(raw, dtype='<i2').reshape(-1, ch)
    elif width == 4:
        a = (np.frombuffer(raw, dtype='<i4').astype(np.float32) / 2147483648.0).reshape(-1, ch)
    else:
        raise ValueError(f"unsupported sample width {width} in {path}")
    pcm = convert_pcm(a, rate, -1.0)
    if os.path.exists(out_path):
        os.remove(out_path)
    with wave.open(out_path, 'wb') as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(gpd['sample_rate'])
        w.writeframes(pcm.astype('<i2').tobytes())
    return out_path


class AudioBase(object):
    """Vocabulary + feature constants (data.py:371-382).  dict.pkl is read from `dict_path`
    (default: ./dict.pkl like the reference, then $ASR_DICT_PKL)."""

    def __init__(self, dict_path=None):
        path = dict_path or ('dict.pkl' if os.path.exists('dict.pkl') else os.environ.get('ASR_DICT_PKL', 'dict.pkl'))
        with open(path, 'rb') as f:
            self.word2int, self.int2word = pickle.load(f)
        self.ms = MelScale(n_mels=gpd['n_mels'], sr=gpd['sample_rate'], f_max=7600, f_min=80, n_stft=257)
        self.window = torch.hann_window(int(gpd['window_len'] * gpd['sample_rate']))


def feature_consts(ms=None, window=None):
    """numpy constant tables handed to asr_create (kept alive by the caller)."""
    if ms is None:
        ms = MelScale(n_mels=gpd['n_mels'], sr=gpd['sample_rate'], f_max=7600, f_min=80, n_stft=257)
    if window is None:
        window = torch.hann_window(int(gpd['window_len'] * gpd['sample_rate']))
    return {
        "mel_fb": np.ascontiguousarray(ms.fb.numpy(), dtype=np.float32),
        "window": np.ascontiguousarray(window.numpy(), dtype=np.float32),
        "taps": delta_filter_stack(),
        "preemphasis": float(gpd['preemphasis']),
    }


_default_engine = None


def set_default_engine(engine):
    """The Model whose handle get_log_mel uses (set by Model.load / Model.load_state)."""
    global _default_engine
    _default_engine = engine


def get_log_mel(training, file_path, ms, window, data_aug=False, engine=None):
    """data.py:167-280 on the device: WAV path (or float32 array) -> CUDA tensor [L, 720]
    (un-normalised, like the reference; main.py:37 applies the CMVN).  training / data_aug must
    be False: dither and augmentation are training-only (data.py:181-200) and out of scope."""
    if training or data_aug:
        raise NotImplementedError("training-time dither / augmentation is outside the inference path")
    eng = engine or _default_engine
    if eng is None:
        raise RuntimeError("get_log_mel needs a loaded Model (Model.load) to own the device handle")
    return eng.features([_waveform(file_path)], normalise=False)[0]


def _waveform(src):
    """WAV path or array -> int16 (as stored; converted on the device) or float32 samples."""
    if isinstance(src, str):
        return read_pcm(src)
    a = src.detach().cpu().numpy() if isinstance(src, torch.Tensor) else np.asarray(src)
    if a.dtype == np.int16:
        return a
    if not np.issubdtype(a.dtype, np.floating):
        raise TypeError(f"waveform dtype {a.dtype}: only float (in [-1, 1)) or int16 PCM is accepted")
    return a.astype(np.float32)


# ---------------------------------------------------------------------------------------------
# batched front end: AudioDst / AudioLoader (data.py:385-540), eval and infer modes

class AudioDst(object):
    """data.py:392-459 without the training branch.  Items are (waveform, text_int) or (waveform,):
    the waveform stays raw (int16 as stored) because the features of a whole batch are computed in
    one pass on the device by AudioLoader; `feature(idx)` gives the reference's per-item result."""

    def __init__(self, audio_base, mode='train', dev_or_test='dev', path_list=None, text_list=None):
        # same signature and defaults as the reference (data.py:393); the default mode is the one this package
        # does not implement, so it has to be spelled out: AudioDst(ab, 'eval', 'dev', paths, texts) / 'infer'
        assert mode in ('train', 'eval', 'infer'), "mode must be train, eval or infer"
        if dev_or_test is not None:
            assert dev_or_test in ('dev', 'test'), "dev_or_test must be dev or test"
        if mode == 'train':
            raise NotImplementedError("training data pipeline (augmentation, TrainSampler) is outside the inference path")
        if path_list is None:
            raise ValueError("path_list is required (the AISHELL manifests of AudioBase are not part of this package)")
        if text_list is not None:
            assert len(text_list) == len(path_list)
        else:
            assert mode == 'infer'
        self.path_list = path_list
        self.text_list = text_list
        self.word2int = audio_base.word2int
        self.data_aug = False
        self.audio_base = audio_base
        self.mode = mode

    def __len__(self):
        return len(self.path_list)

    def _text_int(self, idx):
        text = self.text_list[idx]
        return [self.word2int.get(ele, self.word2int['<unk>']) for ele in text]      # data.py:455

    def __getitem__(self, idx):
        p = self.path_list[idx]
        pcm = _waveform(p)
        if self.text_list is not None:
            return pcm, self._text_int(idx)
        return (pcm,)

    def feature(self, idx, engine=None):
        """What the reference's __getitem__ returns first: get_log_mel of item idx (un-normalised)."""
        return get_log_mel(False, self.path_list[idx], self.audio_base.ms, self.audio_base.window, False, engine)


class AudioLoader(object):
    """data.py:462-540 for eval / infer: `.loader` iterates (t, lens, text) batches of
    gpd['eval_batch_size'] items in order - t: list of normalised [L_i, 720] CUDA tensors,
    lens: IntTensor, text: list of int lists or None."""

    def __init__(self, dst, engine=None, batch_size=None):
        self.dst = dst
        self.mode = dst.mode
        self.bsz = batch_size or gpd.get('eval_batch_size', 256)
        self.engine = engine
        self.loader = self

    def __len__(self):
        return (len(self.dst) + self.bsz - 1) // self.bsz

    def __iter__(self):
        eng = self.engine or _default_engine
        if eng is None:
            raise RuntimeError("AudioLoader needs a loaded Model (Model.load) to own the device handle")
        for b0 in range(0, len(self.dst), self.bsz):
            items = [self.dst[i] for i in range(b0, min(len(self.dst), b0 + self.bsz))]
            pcms = [it[0] for it in items]
            if not all(p.dtype == np.int16 for p in pcms):
                pcms = [p.astype(np.float32) / 32768.0 if p.dtype == np.int16 else np.asarray(p, np.float32) for p in pcms]
            # get_log_mel of every item + batch_audio's normalisation (eps 1e-7), one pass on the device
            t = eng.features(pcms, normalise=bool(gpd['normalize']), eps=1e-7)
            lens = torch.IntTensor([x.size(0) for x in t])
            text = [it[1] for it in items] if len(items[0]) == 2 else None
            yield t, lens, text

    @staticmethod
    def collate_fn(batch, engine=None):
        """data.py:496-507 on items that already carry features: [(feature, text_int)] | [(feature,)]."""
        t, lens = AudioLoader.batch_audio([ele[0] for ele in batch], engine)
        if len(batch[0]) == 2:
            return t, lens, [ele[1] for ele in batch]
        return t, lens, None

    @staticmethod
    def batch_audio(batch, engine=None):
        """data.py:509-518 (RNN encoders: the batch stays a list): instance normalisation on the device."""
        eng = engine or _default_engine
        if eng is None:
            raise RuntimeError("batch_audio needs a loaded Model (Model.load) to own the device handle")
        lens = torch.IntTensor([t.size(0) for t in batch])
        if gpd['normalize']:
            batch = eng.cmvn(batch, eps=1e-7)
        return batch, lens
