"""CPU: the oracle's restatement of the whole-file-exact chunked mode against the whole-file chain and
the reference's golden vector -- this is what pins the halo / alignment / state-carry scheme the GPU path implements."""
import pytest
import torch

import oracle
from oracle import pipeline as opipe
from oracle.models import calibrate_batchnorm, stereo_forward_window
from oracle.weights import make_input


@pytest.mark.parametrize("N,chunk", [(3001, 768), (2400, 512), (777, 400), (5000, 1027)])
def test_restore_exact_equals_restore_whole(state_dicts, N, chunk):
    audio = make_input(1, N, 1235, scale=0.3)[0]
    with torch.no_grad():
        whole = opipe.restore_whole(state_dicts, audio)
        exact = opipe.restore_exact(state_dicts, audio, chunk_size=chunk)
    assert exact.shape == whole.shape
    assert float((whole - exact).abs().max()) <= 1e-6          # same arithmetic, different conv blocking: rounding only


def test_restore_exact_matches_reference_golden(state_dicts, golden):
    audio = make_input(1, 3001, 1235, scale=0.3)[0]
    with torch.no_grad():
        exact = opipe.restore_exact(state_dicts, audio, chunk_size=768)
        nosr = opipe.restore_exact(state_dicts, audio, chunk_size=768, enable_super_resolution=False)
    assert float((torch.from_numpy(golden["chain_whole"]) - exact).abs().max()) <= 1e-6
    assert float((torch.from_numpy(golden["chain_whole_nosr"]) - nosr).abs().max()) <= 1e-6


def test_exact_mode_needs_the_halos(state_dicts):
    """Negative control: the same chunking WITHOUT state carry (LSTM reset per segment) is visibly different from the
    whole-file result -- the test above would catch a scheme that silently dropped the carry."""
    audio = opipe.normalize_audio(make_input(1, 3001, 1235, scale=0.3)[0])
    sd = state_dicts["stereo"]
    x = audio.unsqueeze(0)
    with torch.no_grad():
        full = oracle.stereo_forward(sd, x)
        y, _ = stereo_forward_window(sd, x[:, :, 960:], None, lstm_start=24, state_pos=None)
    # (random-init LSTMs forget within ~30 steps, so look right behind the restart)
    assert float((full[:, :, 960 + 24:960 + 36] - y[:, :, 24:36]).abs().max()) > 1e-4


def test_window_with_carried_state_is_exact(state_dicts):
    sd = state_dicts["stereo"]
    x = make_input(2, 2000, 3)
    with torch.no_grad():
        full = oracle.stereo_forward(sd, x)
        _, st = stereo_forward_window(sd, x[:, :, :1040], None, 0, 1000 - 16)
        y, _ = stereo_forward_window(sd, x[:, :, 960:], st, 24, None)
    assert float((full[:, :, 1000:] - y[:, :, 40:]).abs().max()) <= 1e-6


def test_calibrated_batchnorm_is_a_valid_checkpoint(state_dicts):
    x = make_input(2, 4096, 5)
    sd, stats = calibrate_batchnorm("stereo", state_dicts["stereo"], x)
    assert set(sd) == set(state_dicts["stereo"]) and min(stats.values()) < 0.01      # var << 1: fold gains >> 1
    assert not torch.equal(sd["encoder.0.1.running_var"], state_dicts["stereo"]["encoder.0.1.running_var"])
    with torch.no_grad():
        y = oracle.stereo_forward(sd, x)
    assert torch.isfinite(y).all() and 0.05 < float(y.pow(2).mean().sqrt()) < 5.0
