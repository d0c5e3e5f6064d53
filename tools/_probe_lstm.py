import sys, torch
sys.path.insert(0, '.')
from oracle.weights import make_state_dict, make_input
from ml_audio_restoration_b200.models import StereoSeparator
m = StereoSeparator(); m.load_state_dict(make_state_dict("stereo")); m = m.cuda().eval()
for B in (1184, 1776):
    x = make_input(B, 8192).cuda()
    with torch.no_grad():
        m(x); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m(x); e1.record(); e1.synchronize()
    print(B, "stereo forward T=8192: %.2f ms" % e0.elapsed_time(e1))
