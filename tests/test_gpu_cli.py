"""GPU: `restore_audio()` and the CLI `main()` executed end to end (reference: src/inference.py:17-108, :111-143):
three `{'model_state_dict': ...}` checkpoints on disk -> torch.load -> load_state_dict -> WAV in -> chain -> WAV out."""
import numpy as np
import pytest
import torch

import oracle
from oracle import pipeline as opipe
from oracle.weights import make_input
from ml_audio_restoration_b200 import restore_audio, save_audio, load_audio
from ml_audio_restoration_b200.audio_processing import _read_wav, load_audio_cuda
from ml_audio_restoration_b200.inference import main
from gpu_util import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture()
def checkpoints(state_dicts, tmp_path):
    """Checkpoint files as the reference trainer writes them (trainer.py:727-734): a dict with 'model_state_dict'."""
    paths = {}
    for name in oracle.MODEL_NAMES:
        p = tmp_path / f"{name}.pth"
        torch.save({"epoch": 3, "model_state_dict": state_dicts[name], "best_val_loss": 0.1, "history": {}}, p)
        paths[name] = str(p)
    return paths


def test_restore_audio_float_wav_matches_reference_golden(checkpoints, golden, tmp_path, capsys):
    """Float32 WAV in (bit-faithful samples) -> `restore_audio` (default mode 'whole' = reference semantics) -> float32 WAV
    out == the reference's golden `chain_whole` / `chain_whole_nosr`; progress prints as in the reference."""
    audio = make_input(1, 3001, 1235, scale=0.3)[0]
    src, dst = tmp_path / "in.wav", tmp_path / "out" / "restored.wav"
    save_audio(str(src), audio, 22050)                                  # 32-bit IEEE float, what torchaudio.save writes
    restore_audio(str(src), str(dst), denoiser_checkpoint=checkpoints["denoiser"],
                  super_res_checkpoint=checkpoints["super_resolution"], stereo_checkpoint=checkpoints["stereo"], device="cuda")
    out = capsys.readouterr().out
    for line in ("Processing:", "Loading audio...", "Loading denoiser model...", "Loading super-resolution model...",
                 "Loading stereo separator model...", "Applying denoising...", "Applying stereo separation...",
                 "Restoration complete!", "Output sample rate: 44100Hz (bandwidth extended)"):
        assert line in out
    y, sr = _read_wav(str(dst))
    assert sr == 44100 and y.shape == (2, 6002)
    assert_close(torch.from_numpy(golden["chain_whole"]), y, "restore_audio (files, checkpoints) vs reference golden chain_whole")
    dst2 = tmp_path / "nosr.wav"
    restore_audio(str(src), str(dst2), checkpoints["denoiser"], checkpoints["super_resolution"], checkpoints["stereo"],
                  22050, False, "cuda")                                 # positional, enable_super_resolution=False
    y2, sr2 = _read_wav(str(dst2))
    assert sr2 == 22050
    assert_close(torch.from_numpy(golden["chain_whole_nosr"]), y2, "restore_audio without super-resolution vs golden")


def test_cli_main_pcm16_44k_input_all_modes(checkpoints, state_dicts, tmp_path):
    """`main()` with the reference's flags on a 16-bit PCM 44.1 kHz stereo file (decoded, mixed to mono and resampled on
    the GPU), in the three modes; 'whole' and 'exact' against the oracle's whole-file chain on the same decoded signal,
    'chunked' against the oracle's chunked scheme."""
    g = torch.Generator().manual_seed(8)
    stereo_in = 0.2 * torch.randn(2, 2 * 9000, generator=g)
    src = tmp_path / "side.wav"
    save_audio(str(src), stereo_in, 44100, encoding="pcm16")
    mono, _ = load_audio(str(src), 22050, mono=True)                    # host restatement of the same front end
    common = ["--denoiser", checkpoints["denoiser"], "--super-res", checkpoints["super_resolution"],
              "--stereo", checkpoints["stereo"], "--sample-rate", "22050", "--device", "cuda"]
    outs = {}
    for mode, extra in (("whole", []), ("exact", ["--chunk-size", "2048"]), ("chunked", ["--chunk-size", "2048", "--overlap", "256"])):
        dst = tmp_path / f"{mode}.wav"
        main([str(src), str(dst), "--mode", mode] + extra + common)
        outs[mode], sr = _read_wav(str(dst))
        assert sr == 44100 and outs[mode].shape == (2, 2 * mono.shape[1])
    ref_whole = opipe.restore_whole(state_dicts, mono)
    assert_close(ref_whole, outs["whole"], "CLI --mode whole vs oracle")
    assert_close(ref_whole, outs["exact"], "CLI --mode exact vs oracle whole-file chain")
    assert float((outs["whole"] - outs["exact"]).abs().max()) <= 1e-6
    assert_close(opipe.restore_chunked(state_dicts, mono, chunk_size=2048, overlap=256), outs["chunked"], "CLI --mode chunked vs oracle")


def test_float_wav_round_trip_through_the_gpu_loader(checkpoints, tmp_path):
    """The package's own default output format (32-bit float WAV, format tag 3) loads through `load_audio_cuda` -- the
    stdlib `wave` module rejects it -- so a restored file can be restored again; stereo float input is mixed to mono."""
    a = make_input(2, 5000, 21, scale=0.2)[:, 0]                        # [2, 5000] "stereo"
    p = tmp_path / "f32.wav"
    save_audio(str(p), a, 22050)
    dev_audio, sr = load_audio_cuda(str(p), 22050, "cuda")
    assert sr == 22050 and dev_audio.shape == (1, 5000)
    assert float((dev_audio.cpu() - a.mean(dim=0, keepdim=True)).abs().max()) <= 1e-7
    p44 = tmp_path / "f32_44k.wav"
    save_audio(str(p44), a, 44100)
    dev2, _ = load_audio_cuda(str(p44), 22050, "cuda")
    host2, _ = load_audio(str(p44), 22050, mono=True)
    assert_close(host2, dev2, "float WAV 44.1k -> 22.05k: GPU resampler vs torchaudio", max_abs=3e-6, min_snr=100.0)
    out1, out2 = tmp_path / "r1.wav", tmp_path / "r2.wav"
    kw = dict(denoiser_checkpoint=checkpoints["denoiser"], super_res_checkpoint=checkpoints["super_resolution"],
              stereo_checkpoint=checkpoints["stereo"], device="cuda")
    restore_audio(str(p), str(out1), **kw)
    restore_audio(str(out1), str(out2), sample_rate=44100, enable_super_resolution=False, **kw)   # re-restore the float output
    y2, sr2 = _read_wav(str(out2))
    assert sr2 == 44100 and y2.shape == (2, 10000) and torch.isfinite(y2).all()


def test_restore_audio_error_behaviour(checkpoints, tmp_path):
    src = tmp_path / "x.wav"
    save_audio(str(src), make_input(1, 2000, 1)[0], 22050)
    with pytest.raises(RuntimeError):                                   # no CPU fallback
        restore_audio(str(src), str(tmp_path / "o.wav"), device="cpu")
    with pytest.raises(FileNotFoundError):                              # reference: torch.load of the default checkpoint path
        restore_audio(str(src), str(tmp_path / "o.wav"), device="cuda")
    bad = tmp_path / "bad.pth"
    torch.save({"model_state_dict": {"nope": torch.zeros(1)}}, bad)
    with pytest.raises(RuntimeError):                                   # load_state_dict(strict)
        restore_audio(str(src), str(tmp_path / "o.wav"), denoiser_checkpoint=str(bad), device="cuda")
