"""Distance of every tapped activation and of the heads from the float64 network, GPU path vs the reference's own fp32 arithmetic
(GPU box): the shipped 256x320 model and its 3-channel variant on random-pixel inputs. Quoted in DESIGN.md 5.2."""
import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import yolo_fastest_b200 as yf
from oracle import yolo_oracle as O
sd = torch.load('/root/repo/tests/golden/weights/yolo_fastest_256x320.pth', map_location='cpu')
def run(sd, cin, tag):
    m = yf.YoloFastest({"num_cls": 3, "input_channel": cin, "num_anchors": 3}); m.load_state_dict(sd); m = m.cuda().eval()
    x = (torch.randint(0, 256, (2, cin, 256, 320), generator=torch.Generator().manual_seed(29)).float() - 128.0) / 255.0
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    taps32, taps64 = {}, {}
    r32 = O.forward(sd, x, taps32); r64 = O.forward(sd64, x.double(), taps64)
    got = m(x.cuda())
    for name in ("conv1_4", "res1_1", "conv2_1", "res2_2", "conv3_4", "res3_6", "res4_4", "res5_5"):
        g = m.tap(name, 2).cpu().double().reshape(taps64[name].shape)
        sc = taps64[name].abs().max().item()
        print(tag, name, "gpu-fp64 %.2e  ref32-fp64 %.2e (of scale)" % ((g - taps64[name]).abs().max().item() / sc, (taps32[name].double() - taps64[name]).abs().max().item() / sc))
    for h in range(2):
        sc = r64[h].abs().max().item()
        print(tag, "head", h, "gpu %.2e ref %.2e scale %.1f" % ((got[h].cpu().double() - r64[h]).abs().max().item() / sc, (r32[h].double() - r64[h]).abs().max().item() / sc, sc))
run(sd, 1, "1ch")
sd3 = dict(sd); w = sd3["conv0.0.weight"]; gen = torch.Generator().manual_seed(5)
sd3["conv0.0.weight"] = torch.cat([w * c for c in (0.5, 0.3, 0.2)], 1) + 0.05 * w.abs().mean() * torch.randn((8, 3, 3, 3), generator=gen)
run(sd3, 3, "3ch")
