"""Scratch: nrrt_conv_layer with the identity affine (+ residual) as the training step uses it for data gradients, vs torch fp32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C, torch, torch.nn.functional as F
from mgr_sampling_cnn_b200 import _lib
from mgr_sampling_cnn_b200.cnn import _Layer, K3S1, K1S1
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0"); lib = _lib.load()
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
torch.manual_seed(0)
for (n, h, w, cin, cout, k, use_res, ident, per_tap) in [(2, 512, 512, 64, 64, 3, True, True, False), (2, 512, 512, 64, 64, 3, True, True, True),
        (2, 512, 512, 64, 64, 3, False, True, False), (2, 512, 512, 64, 64, 3, True, False, False), (2, 128, 128, 64, 64, 3, True, True, False),
        (2, 256, 256, 128, 128, 3, True, True, False), (2, 512, 512, 128, 192, 3, False, True, False), (2, 512, 512, 64, 128, 3, False, True, False),
        (2, 256, 256, 256, 64, 1, True, True, False)]:
    x = torch.randn(n, h, w, cin, device=dev).half()
    wt = (torch.randn(cout, k * k, cin, device=dev) * 0.05).half()
    res = torch.randn(n, h, w, cout, device=dev).half() if use_res else None
    scale = None if ident else torch.ones(cout, device=dev)
    shift = None if ident else torch.zeros(cout, device=dev)
    L = _Layer(2, K3S1 if k == 3 else K1S1, cin, cout, wt, scale, shift, False)
    L.force_per_tap = per_tap
    out = torch.zeros(n, h, w, cout, device=dev).half()
    L.run(lib, st, x, 0, (1, h, w), out, 0, res=res, res_coff=0)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float().reshape(cout, k, k, cin).permute(0, 3, 1, 2), padding=k // 2).permute(0, 2, 3, 1)
    if use_res:
        ref = ref + res.float()
    d = (out.float() - ref)
    print("n%d %dx%d cin%d cout%d k%d res=%d ident=%d per_tap=%d: rel L2 err %.3e  max abs %.3e (ref max %.2f) bad-pixel frac %.4f" % (
        n, h, w, cin, cout, k, use_res, ident, per_tap, float(d.norm() / ref.norm()), float(d.abs().max()), float(ref.abs().max()),
        float((d.abs().amax(-1) > 0.05).float().mean())))
