import sys, torch
sys.path.insert(0, '.')
from oracle.weights import make_state_dict, make_input
from ml_audio_restoration_b200.models import StereoSeparator, AudioSuperResolution
which = sys.argv[1] if len(sys.argv) > 1 else "stereo"
B, T = int(sys.argv[2]) if len(sys.argv) > 2 else 148, int(sys.argv[3]) if len(sys.argv) > 3 else 44100
if which == "stereo":
    m = StereoSeparator(); m.load_state_dict(make_state_dict("stereo"))
else:
    m = AudioSuperResolution(upscale_factor=2); m.load_state_dict(make_state_dict("super_resolution"))
m = m.cuda().eval()
x = make_input(B, T).cuda()
with torch.no_grad():
    for _ in range(2):
        y = m(x)
torch.cuda.synchronize()
print("ok", tuple(y.shape))
