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
