"""GPU: the device front end of load_audio (ar_pcm16_to_float, ar_resample_mono) against the oracle and the torchaudio
golden vectors, through the C-ABI."""
import os
import sys
import wave

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden_io import CASES, make_input  # noqa: E402
from oracle import audio_io  # noqa: E402
from ml_audio_restoration_b200.audio_processing import resample_mono_cuda, load_audio_cuda, load_audio  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_io_v1.npz")


@pytest.mark.parametrize("i", range(len(CASES)))
def test_resample_matches_torchaudio_golden(i):
    gold = np.load(GOLD)
    orig, new, C, N, seed = (int(v) for v in gold[f"case{i}_meta"])
    x = torch.from_numpy(make_input(seed, C, N)).cuda()
    y = resample_mono_cuda(x, orig, new).cpu().numpy()
    ref = gold[f"case{i}_y"]
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() <= 2e-6


@pytest.mark.parametrize("orig,new,C,N", [(44100, 22050, 2, 1_000_003), (48000, 22050, 1, 480_001), (22050, 22050, 2, 5000),
                                           (22050, 22050, 1, 17), (11025, 22050, 1, 1)])
def test_resample_vs_oracle_large_and_edge(orig, new, C, N):
    x = make_input(N % 97, C, N)
    y = resample_mono_cuda(torch.from_numpy(x).cuda(), orig, new).cpu().numpy()
    ref = audio_io.load_front_end(x, orig, new)
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() <= 3e-6


def test_pcm16_wav_load_on_device(tmp_path):
    rng = np.random.default_rng(3)
    pcm = rng.integers(-32768, 32767, size=(30_001, 2), dtype=np.int16)
    path = str(tmp_path / "in.wav")
    with wave.open(path, "wb") as w:
        w.setnchannels(2)
        w.setsampwidth(2)
        w.setframerate(44100)
        w.writeframes(pcm.tobytes())
    got, sr = load_audio_cuda(path, sample_rate=22050)
    assert sr == 22050 and got.is_cuda
    ref = audio_io.load_front_end(audio_io.pcm16_to_float(pcm), 44100, 22050)       # oracle
    host, _ = load_audio(path, sample_rate=22050)                                   # reference-style host path (torchaudio)
    assert got.shape == ref.shape == tuple(host.shape)
    assert np.abs(got.cpu().numpy() - ref).max() <= 3e-6
    assert float((got.cpu() - host).abs().max()) <= 3e-6


def test_resample_errors():
    with pytest.raises(RuntimeError):
        resample_mono_cuda(torch.zeros(1, 10), 44100, 22050)        # host tensor: no CPU fallback


FORMATS = [audio_io.PCM_U8, audio_io.PCM_S16, audio_io.PCM_S24, audio_io.PCM_S32, audio_io.PCM_F32, audio_io.PCM_F64]


@pytest.mark.parametrize("fmt", FORMATS)
def test_every_wav_encoding_decodes_on_device(tmp_path, fmt):
    """`ar_pcm_to_float` is bit-exact against the oracle's decode for every sample encoding (full code range), and a file
    of that encoding -- plain and WAVE_FORMAT_EXTENSIBLE -- loads on the device like the host path (decode, mono mix,
    44.1 -> 22.05 kHz resample)."""
    from test_oracle_io import encoding_samples
    from ml_audio_restoration_b200.audio_processing import decode_pcm_cuda, wav_info
    n, ch = 20_003, 2
    a = encoding_samples(fmt, n, ch, 100 + fmt)
    for extensible in (False, True):
        path = str(tmp_path / f"in_{fmt}_{int(extensible)}.wav")
        data = audio_io.write_wav(path, a, fmt, 44100, extensible=extensible, junk=extensible)
        ref_planar = audio_io.pcm_to_float(data, fmt, ch)
        raw = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
        got_planar = decode_pcm_cuda(raw, fmt, ch, n)
        assert np.array_equal(got_planar.cpu().numpy(), ref_planar)                       # bit-exact
        assert wav_info(path)[:4] == (fmt, ch, 44100, n)
        got, sr = load_audio_cuda(path, sample_rate=22050)
        ref = audio_io.load_front_end(ref_planar, 44100, 22050)
        host, _ = load_audio(path, sample_rate=22050)
        assert sr == 22050 and got.is_cuda and got.shape == ref.shape == tuple(host.shape)
        assert np.abs(got.cpu().numpy() - ref).max() <= 3e-6
        assert float((got.cpu() - host).abs().max()) <= 3e-6


def test_decode_errors():
    from ml_audio_restoration_b200.audio_processing import decode_pcm_cuda
    with pytest.raises(RuntimeError):
        decode_pcm_cuda(torch.zeros(8, dtype=torch.uint8), audio_io.PCM_S16, 1, 4)        # host tensor: no CPU fallback
    with pytest.raises(RuntimeError):
        decode_pcm_cuda(torch.zeros(8, dtype=torch.uint8).cuda(), 99, 1, 4)               # unknown format
