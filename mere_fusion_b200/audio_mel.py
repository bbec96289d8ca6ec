"""Wav2Lip mel front-end on the host: numpy/scipy restatement of wav2lip/audio.py:45-51 (+ :20-23,
57-61, 92-122) with the constants of wav2lip/hparams.py:41-73.

The reference calls librosa (requirements.txt:6, unpinned, absent from this image) at audio.py:61
(`librosa.stft`) and audio.py:100 (`librosa.filters.mel`).  Both are restated from their published
algorithms, librosa >= 0.10 semantics: centred STFT with ZERO padding (`pad_mode="constant"`),
periodic Hann window, Slaney-scale mel filterbank with Slaney area normalisation in float32.
PARITY UNPINNED: no golden vector of the reference exists for this stage.
"""
import numpy as np
from scipy import signal

NUM_MELS = 80
N_FFT = 800
HOP_SIZE = 200
WIN_SIZE = 800
SAMPLE_RATE = 16000
PREEMPHASIS = 0.97
MIN_LEVEL_DB = -100
REF_LEVEL_DB = 20
FMIN = 55
FMAX = 7600
MAX_ABS_VALUE = 4.0

_mel_basis = None


def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr=SAMPLE_RATE, n_fft=N_FFT, n_mels=NUM_MELS, fmin=FMIN, fmax=FMAX):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with htk=False, norm='slaney', dtype float32"""
    fftfreqs = np.linspace(0, float(sr) / 2, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights.astype(np.float32)


def stft(y, n_fft=N_FFT, hop=HOP_SIZE, win=WIN_SIZE):
    """librosa.stft(y, n_fft, hop_length, win_length), center=True, pad_mode='constant', window='hann'"""
    y = np.asarray(y)
    w = signal.get_window("hann", win, fftbins=True)
    if win < n_fft:
        lp = (n_fft - win) // 2
        w = np.pad(w, (lp, n_fft - win - lp))
    yp = np.pad(y, n_fft // 2, mode="constant")
    n_frames = 1 + (len(yp) - n_fft) // hop
    idx = np.arange(n_fft)[:, None] + hop * np.arange(n_frames)[None, :]
    return np.fft.rfft(yp[idx] * w[:, None], axis=0)


def melspectrogram(wav):
    """audio.py:45-51: float[n] -> [80, n // 200 + 1]"""
    global _mel_basis
    wav = signal.lfilter([1, -PREEMPHASIS], [1], wav)                          # audio.py:20-23
    D = stft(wav)
    if _mel_basis is None:
        _mel_basis = mel_filterbank()
    min_level = np.exp(MIN_LEVEL_DB / 20 * np.log(10))                          # audio.py:103-105
    S = 20 * np.log10(np.maximum(min_level, np.dot(_mel_basis, np.abs(D)))) - REF_LEVEL_DB
    return np.clip((2 * MAX_ABS_VALUE) * ((S - MIN_LEVEL_DB) / (-MIN_LEVEL_DB)) - MAX_ABS_VALUE,
                   -MAX_ABS_VALUE, MAX_ABS_VALUE)                                # audio.py:110-114
