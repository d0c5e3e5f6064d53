"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the synthetic 78 rpm degradation generator
(reference: src/utils/audio_processing.py:122-226 `simulate_vinyl_artifacts`).

What the reference does, in order (all of it restated here, line numbers of the reference in brackets):
  1. surface noise   `audio + randn_like(audio) * U(surface_noise_level)`                          [152-153]
  2. pops            Poisson count, uniform locations / amplitudes, +-1 polarity (p = .45/.55), per pop an
                     exponential decay `amp*pol*exp(-k/(sr*decay_time*0.3))` plus, when longer than 10 samples,
                     a decaying sine `0.3*sin(2 pi f k/sr)*decay*amp*0.2`; built in float64, rounded to float32,
                     added in float32 to every channel, pop after pop                              [157-188]
  3. crackle         `randn * U(crackle_level)` -> Butterworth high-pass (order 4, 2.5 kHz) `filtfilt`   [191-200]
  4. rumble          `randn * U(0.005, 0.015)`  -> Butterworth low-pass  (order 4, 100 Hz)  `filtfilt`   [203-213]
  5. roll-off        Butterworth low-pass (order 3, U(6, 8) kHz) `filtfilt` of the sum               [216-224]

Third-party arithmetic: `scipy.signal.butter` / `filtfilt` (requirements.txt: `scipy>=1.10.0`, not vendored; installed
here: scipy 1.18.1).  Their published algorithms are restated below in numpy -- `butter` (analog prototype -> frequency
transform -> bilinear -> polynomial), `filtfilt` (odd extension by 3*max(len(a), len(b)) samples IN THE INPUT'S DTYPE,
`lfilter_zi` steady-state initial conditions, forward and backward direct-form-II-transposed passes in float64) -- with
only the serial recurrence itself delegated to `scipy.signal.lfilter` (a pure-Python loop over 44 100 samples x 6 passes
would make the CPU suite take minutes).  `tests/test_oracle_degrade.py` pins `butter` and `filtfilt` against scipy's own
and the whole function against golden vectors produced by the unmodified reference function
(tests/golden/make_golden_degrade.py -> golden_degrade_v1.npz).

Randomness: the reference draws from TWO global generators -- `torch.randn_like` (three noise tensors: surface,
crackle, rumble, in that order) and `np.random` (levels, pop plan, roll-off frequency).  `draw_plan` consumes
`np.random` in exactly the reference's call order, so seeding both generators reproduces the reference bit for bit.
Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this package.
"""
import numpy as np
import torch
from scipy.signal import lfilter


# ----------------------------------------------------------------------------- scipy.signal.butter, restated
def butter(order: int, wn: float, btype: str = "low"):
    """(b, a) float64 -- `scipy.signal.butter(order, wn, btype)` for 'low' / 'high' (digital, fs = 2)."""
    m = np.arange(-order + 1, order, 2)
    p = -np.exp(1j * np.pi * m / (2 * order))                      # buttap
    k = 1.0
    fs = 2.0
    warped = 2 * fs * np.tan(np.pi * wn / fs)                      # pre-warp
    if btype == "low":                                             # lp2lp_zpk
        z = np.zeros(0, dtype=complex)
        p = warped * p
        k = k * warped ** order
    elif btype == "high":                                          # lp2hp_zpk
        k = k * np.real(1.0 / np.prod(-p))
        p = warped / p
        z = np.zeros(order, dtype=complex)
    else:
        raise ValueError(btype)
    fs2 = 2.0 * fs                                                 # bilinear_zpk
    degree = len(p) - len(z)
    zz = np.append((fs2 + z) / (fs2 - z), -np.ones(degree))
    pz = (fs2 + p) / (fs2 - p)
    kz = k * np.real(np.prod(fs2 - z) / np.prod(fs2 - p))
    b = kz * np.real(np.poly(zz))                                  # zpk2tf
    a = np.real(np.poly(pz))
    return b, a


def lfilter_zi(b, a):
    """Steady-state DF2T state for a unit step input (`scipy.signal.lfilter_zi`)."""
    b = np.asarray(b, dtype=np.float64) / a[0]
    a = np.asarray(a, dtype=np.float64) / a[0]
    # scipy 1.18: y_inf = sum(b) / sum(a) is the steady-state output for a unit step, and the transposed-direct-form
    # state obeys zi[k] = zi[k+1] + b[k+1] - y_inf a[k+1]  (reverse cumulative sum, the last term dropped).
    # (The 100 Hz low-pass has sum(a) ~ 1e-7: one ulp in zi shows as ~1e-10 in the filtered signal, far below float32.)
    y_inf = np.sum(b) / np.sum(a)
    return np.cumsum((b - y_inf * a)[::-1])[::-1][1:]


def filtfilt(b, a, x):
    """`scipy.signal.filtfilt(b, a, x)` (method 'pad', padtype 'odd', padlen 3*max(len(a), len(b))) for a 1-D x.

    The extension is formed in x's dtype (float32 in the reference: `crackle_np[i]`, audio_processing.py:198), the two
    filter passes run in float64, the result is float64 (the reference's assignment back into a float32 array rounds it).
    """
    x = np.asarray(x)
    pad = 3 * max(len(a), len(b))
    if x.shape[-1] <= pad:
        raise ValueError(f"The length of the input vector x must be greater than padlen, which is {pad}.")
    left = 2 * x[0] - x[pad:0:-1]
    right = 2 * x[-1] - x[-2:-(pad + 2):-1]
    ext = np.concatenate((left, x, right))
    zi = lfilter_zi(b, a)
    y, _ = lfilter(b, a, ext, zi=zi * ext[0])
    y, _ = lfilter(b, a, y[::-1], zi=zi * y[-1])
    return y[::-1][pad:-pad]


# ----------------------------------------------------------------------------- the np.random plan
def draw_plan(num_samples, sample_rate, impulse_rate=10.0, impulse_amplitude=(0.1, 0.5),
              surface_noise_level=(0.015, 0.03), crackle_level=(0.01, 0.02), add_rumble=True, add_rolloff=True):
    """Consume `np.random` exactly as audio_processing.py:152-219 does and return the drawn parameters."""
    duration = num_samples / sample_rate
    plan = {"surface_level": np.random.uniform(*surface_noise_level), "pops": []}
    num_pops = np.random.poisson(int(duration * impulse_rate))
    if num_pops > 0:
        locs = np.random.randint(0, num_samples, num_pops)
        amps = np.random.uniform(*impulse_amplitude, num_pops)
        pols = np.random.choice([-1, 1], num_pops, p=[0.45, 0.55])
        for loc, amp, pol in zip(locs, amps, pols):
            decay_time = np.random.uniform(0.001, 0.003) * (1 + amp)
            length = min(int(sample_rate * decay_time), num_samples - loc)
            if length <= 0:
                continue
            freq = np.random.uniform(3000, 8000) if length > 10 else None
            plan["pops"].append({"loc": int(loc), "amp": float(amp), "polarity": int(pol), "decay_time": float(decay_time),
                                 "length": int(length), "resonance_freq": None if freq is None else float(freq)})
    plan["crackle_level"] = np.random.uniform(*crackle_level)
    plan["rumble_level"] = np.random.uniform(0.005, 0.015) if add_rumble else None
    plan["rolloff_hz"] = np.random.uniform(6000, 8000) if add_rolloff else None
    return plan


def pop_impulse(pop, sample_rate):
    """float64 impulse of one pop (audio_processing.py:174-186)."""
    k = np.arange(pop["length"])
    decay = np.exp(-k / (sample_rate * pop["decay_time"] * 0.3))
    impulse = pop["amp"] * pop["polarity"] * decay
    if pop["resonance_freq"] is not None:
        t = k / sample_rate
        resonance = 0.3 * np.sin(2 * np.pi * pop["resonance_freq"] * t) * decay
        impulse = impulse + resonance * pop["amp"] * 0.2
    return impulse


def apply_plan(audio: torch.Tensor, sample_rate: int, plan, surface: torch.Tensor, crackle: torch.Tensor,
               rumble: torch.Tensor = None) -> torch.Tensor:
    """The deterministic part: `audio`, the three unit-variance noise tensors and the drawn plan -> degraded audio."""
    out = audio.clone() + surface * plan["surface_level"]
    for pop in plan["pops"]:
        imp = torch.from_numpy(pop_impulse(pop, sample_rate)).float()
        out[..., pop["loc"]:pop["loc"] + pop["length"]] += imp
    nyquist = sample_rate / 2

    def filt(t, b, a):
        arr = t.numpy().copy()
        for i in range(arr.shape[0]):
            arr[i] = filtfilt(b, a, arr[i])                       # float64 -> float32 on assignment
        return torch.from_numpy(arr)

    out = out + filt(crackle * plan["crackle_level"], *butter(4, 2500 / nyquist, "high"))
    if plan["rumble_level"] is not None:
        out = out + filt(rumble * plan["rumble_level"], *butter(4, 100 / nyquist, "low"))
    if plan["rolloff_hz"] is not None:
        out = filt(out, *butter(3, plan["rolloff_hz"] / nyquist, "low"))
    return out


def simulate_vinyl_artifacts(audio: torch.Tensor, sample_rate: int, impulse_rate=10.0, impulse_amplitude=(0.1, 0.5),
                             surface_noise_level=(0.015, 0.03), crackle_level=(0.01, 0.02), add_rumble=True,
                             add_rolloff=True) -> torch.Tensor:
    """Same signature and same use of the global torch / numpy generators as the reference function."""
    surface = torch.randn_like(audio)
    crackle = torch.randn_like(audio)
    rumble = torch.randn_like(audio) if add_rumble else None
    plan = draw_plan(audio.shape[-1], sample_rate, impulse_rate, impulse_amplitude, surface_noise_level, crackle_level,
                     add_rumble, add_rolloff)
    return apply_plan(audio, sample_rate, plan, surface, crackle, rumble)
