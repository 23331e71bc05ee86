"""Scratch timing of the weight-gradient kernels on the training step's layer shapes (batch 3 at 512 x 512; CUDA events; not the
bench contract): nrrt_conv_wgrad_nhwc (MN-major operands from channels-last buffers) and nrrt_conv_wgrad (channel-major copies)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C, torch
from mgr_sampling_cnn_b200 import _lib
from mgr_sampling_cnn_b200._lib import check
dev = torch.device("cuda:0"); lib = _lib.load()
P = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
B = 3
tot = {"nhwc": [0.0, 0.0], "cm": [0.0, 0.0]}
def timed(fn):
    for _ in range(2): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 5
for name, co, ci, s, k in [("c8.0", 128, 192, 512, 3), ("c8.2", 64, 128, 512, 3), ("layer1", 64, 64, 512, 3), ("c7.0", 256, 384, 256, 3), ("c7.2", 128, 256, 256, 3),
                           ("layer2", 128, 128, 256, 3), ("c6.0", 512, 1024, 128, 3), ("c6.2", 256, 512, 128, 3), ("layer3", 256, 256, 128, 3),
                           ("layer4", 512, 512, 64, 3), ("up8 (1x1)", 256, 128, 256, 1), ("up6 (1x1)", 2048, 1024, 64, 1)]:
    g = torch.randn(B, s, s, co, device=dev).half()
    x = torch.randn(B, s, s, ci, device=dev).half()
    gc = torch.randn(co, B, s, s, device=dev).half()
    xc = torch.randn(k, ci, B, s, s, device=dev).half()
    dw = torch.zeros(co, k * k, ci, device=dev)
    t1 = timed(lambda: check(lib.nrrt_conv_wgrad_nhwc(P(g), co, co, P(x), ci, ci, B, s, s, k, P(dw), ci, st)))
    t2 = timed(lambda: check(lib.nrrt_conv_wgrad(P(gc), co, P(xc), ci, B, s, s, k, P(dw), ci, st)))
    fl = 2.0 * B * s * s * co * ci * k * k
    for key, t in (("nhwc", t1), ("cm", t2)):
        tot[key][0] += fl; tot[key][1] += t
    print("%-10s co %4d ci %4d %3dx%-3d k%d: nhwc %.3f ms %5.0f TFLOP/s | channel-major %.3f ms %5.0f TFLOP/s" % (name, co, ci, s, s, k, t1, fl / t1 / 1e9, t2, fl / t2 / 1e9))
print("all: nhwc %.0f TFLOP/s, channel-major %.0f TFLOP/s" % (tot["nhwc"][0] / tot["nhwc"][1] / 1e9, tot["cm"][0] / tot["cm"][1] / 1e9))
