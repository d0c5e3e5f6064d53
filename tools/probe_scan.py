import sys, torch
sys.path.insert(0, '.')
from oracle.weights import make_state_dict, make_input
from ml_audio_restoration_b200.models import StereoSeparator
from ml_audio_restoration_b200 import _lib
m = StereoSeparator(); m.load_state_dict(make_state_dict("stereo")); m = m.cuda().eval()
L = _lib.lib()
import ctypes as C
B, T = 2368, 16384
x = make_input(B, T).cuda()
ncat = len(_lib.PROFILE_CATEGORIES)
with torch.no_grad():
    m(x); torch.cuda.synchronize()
    L.ar_profile_enable(1)
    for _ in range(3): m(x)
    torch.cuda.synchronize()
    p_ms, p_fl, p_ln = (C.c_double * ncat)(), (C.c_double * ncat)(), (C.c_longlong * ncat)()
    L.ar_profile_read(p_ms, p_fl, p_ln, ncat)
    L.ar_profile_enable(0)
i = list(_lib.PROFILE_CATEGORIES).index("lstm")
print("lstm ms per launch %.3f  ns/step %.1f" % (p_ms[i] / 3, 1e6 * p_ms[i] / 3 / T))
