"""CPU: the oracle's restatement of the load_audio front end (PCM decode, mono mix, torchaudio sinc resampler) against
golden vectors generated from the installed torchaudio (tests/golden/make_golden_io.py), plus the WAV helpers."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden_io import CASES, make_input  # noqa: E402
from oracle import audio_io  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_io_v1.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("i", range(len(CASES)))
def test_oracle_resample_matches_torchaudio_golden(gold, i):
    orig, new, C, N, seed = (int(v) for v in gold[f"case{i}_meta"])
    assert (orig, new, C, N) == CASES[i]
    x = make_input(seed, C, N)
    y = audio_io.load_front_end(x, orig, new)
    ref = gold[f"case{i}_y"]
    assert y.shape == ref.shape == (1, -(-new * N // orig))
    assert np.abs(y - ref).max() <= 2e-6        # fp32 accumulation order differs from conv1d, nothing else


def test_taps_match_torchaudio_formula():
    import torchaudio.functional.functional as F
    for orig, new in [(44100, 22050), (48000, 22050), (16000, 22050)]:
        import math
        g = math.gcd(orig, new)
        k, w = F._get_sinc_resample_kernel(orig, new, g)
        taps, width, o, n = audio_io.sinc_resample_taps(orig, new)
        assert w == width and tuple(k.shape) == (n, 1, taps.shape[1])
        assert np.abs(k[:, 0].numpy() - taps).max() <= 1e-7


def test_identity_rate_and_pcm16():
    x = make_input(5, 2, 100)
    assert audio_io.resample(x, 22050, 22050) is not None and np.array_equal(audio_io.resample(x, 22050, 22050), x)
    pcm = np.array([[0, -32768], [32767, 1], [-1, 16384]], dtype=np.int16)
    f = audio_io.pcm16_to_float(pcm)
    assert f.shape == (2, 3) and f.dtype == np.float32
    assert f[1, 0] == -1.0 and f[0, 1] == np.float32(32767 / 32768) and f[1, 2] == 0.5


def test_wav_helpers_round_trip(tmp_path):
    from ml_audio_restoration_b200.audio_processing import save_audio, _read_wav, load_audio
    a = torch.from_numpy(make_input(9, 2, 1000))
    p32 = str(tmp_path / "f32.wav")
    save_audio(p32, a, 44100)                       # 32-bit float, as torchaudio.save writes float tensors
    b, sr = _read_wav(p32)
    assert sr == 44100 and torch.equal(a, b)
    p16 = str(tmp_path / "p16.wav")
    save_audio(p16, a, 22050, encoding="pcm16")
    c, sr = _read_wav(p16)
    assert sr == 22050 and float((a.clamp(-1, 1) - c).abs().max()) <= 2.0 / 32768     # quantisation + the 32767/32768 scale pair
    mono, sr = load_audio(p32, sample_rate=22050)   # host path: mean + torchaudio resample == oracle front end
    ref = audio_io.load_front_end(a.numpy(), 44100, 22050)
    assert sr == 22050 and mono.shape == ref.shape and float(np.abs(mono.numpy() - ref).max()) <= 2e-6


# ----------------------------------------------------------------------------- WAV sample encodings (host logic + oracle)
FORMATS = [audio_io.PCM_U8, audio_io.PCM_S16, audio_io.PCM_S24, audio_io.PCM_S32, audio_io.PCM_F32, audio_io.PCM_F64]


def encoding_samples(fmt, n, ch, seed):
    """[n, ch] samples of the encoding's numpy type, full range incl. the extreme codes."""
    rng = np.random.default_rng(seed)
    if fmt == audio_io.PCM_U8:
        a = rng.integers(0, 256, size=(n, ch)).astype(np.uint8)
        a[0, 0], a[1, 0] = 0, 255
    elif fmt == audio_io.PCM_S16:
        a = rng.integers(-32768, 32768, size=(n, ch)).astype(np.int16)
        a[0, 0], a[1, 0] = -32768, 32767
    elif fmt == audio_io.PCM_S24:
        a = rng.integers(-(1 << 23), 1 << 23, size=(n, ch)).astype(np.int32)
        a[0, 0], a[1, 0] = -(1 << 23), (1 << 23) - 1
    elif fmt == audio_io.PCM_S32:
        a = rng.integers(-(1 << 31), 1 << 31, size=(n, ch)).astype(np.int32)
        a[0, 0], a[1, 0] = -(1 << 31), (1 << 31) - 1
    elif fmt == audio_io.PCM_F32:
        a = (rng.standard_normal((n, ch)) * 0.3).astype(np.float32)
    else:
        a = rng.standard_normal((n, ch)) * 0.3
    return a


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("extensible", [False, True])
def test_wav_header_parser_and_oracle_decode(tmp_path, fmt, extensible):
    """`wav_info` finds the data chunk of plain and WAVE_FORMAT_EXTENSIBLE files (odd-sized chunk in front); the oracle's
    decode equals scipy's reader for the sample extraction and libsndfile's 2^(bits-1) convention for the scale."""
    from scipy.io import wavfile
    from ml_audio_restoration_b200.audio_processing import wav_info, _read_wav
    n, ch, sr = 1001, 2 if fmt != audio_io.PCM_F64 else 3, 44100
    a = encoding_samples(fmt, n, ch, 7 + fmt)
    path = str(tmp_path / "x.wav")
    data = audio_io.write_wav(path, a, fmt, sr, extensible=extensible, junk=True)
    code, nch, rate, frames, offset = wav_info(path)
    assert (code, nch, rate, frames) == (fmt, ch, sr, n)
    with open(path, "rb") as f:
        f.seek(offset)
        assert f.read(len(data)) == data
    y = audio_io.pcm_to_float(data, fmt, ch)
    assert y.shape == (ch, n) and y.dtype == np.float32
    rate2, raw = wavfile.read(path)                       # scipy: unscaled samples (24-bit left-justified in int32)
    assert rate2 == sr
    scale = {audio_io.PCM_U8: None, audio_io.PCM_S16: 32768.0, audio_io.PCM_S24: 2147483648.0, audio_io.PCM_S32: 2147483648.0}.get(fmt, 1.0)
    ref = (raw.astype(np.float64) - 128.0) / 128.0 if scale is None else raw.astype(np.float64) / scale
    assert np.abs(y.astype(np.float64) - ref.T).max() <= 2.0 ** -24        # one fp32 rounding (32-bit ints, float64 input)
    host, rate3 = _read_wav(path)                          # the package's host reader (stdlib `wave` / float parser)
    assert rate3 == sr and np.array_equal(host.numpy(), y)


def test_wav_header_parser_rejects_other_files(tmp_path):
    from ml_audio_restoration_b200.audio_processing import wav_info
    p = tmp_path / "not.wav"
    p.write_bytes(b"OggS" + b"\0" * 64)
    with pytest.raises(RuntimeError):
        wav_info(str(p))
    import struct
    body = b"WAVE" + struct.pack("<4sIHHIIHH", b"fmt ", 16, 2, 1, 8000, 4000, 256, 4) + struct.pack("<4sI", b"data", 0)   # ADPCM
    p.write_bytes(struct.pack("<4sI", b"RIFF", len(body)) + body)
    with pytest.raises(RuntimeError):
        wav_info(str(p))
