"""Scratch timing of nrrt_conv_wgrad on the training step's layer shapes (batch 3 at 512 x 512; CUDA events; not the bench contract)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C, torch
from mgr_sampling_cnn_b200 import _lib
from mgr_sampling_cnn_b200._lib import check
dev = torch.device("cuda:0"); lib = _lib.load()
P = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
B = 3
tot_f = tot_t = 0.0
for name, co, ci, s, k in [("c8.0", 128, 192, 512, 3), ("c8.2", 64, 128, 512, 3), ("layer1", 64, 64, 512, 3), ("c7.0", 256, 384, 256, 3), ("c7.2", 128, 256, 256, 3),
                           ("layer2", 128, 128, 256, 3), ("c6.0", 512, 1024, 128, 3), ("c6.2", 256, 512, 128, 3), ("layer3", 256, 256, 128, 3),
                           ("layer4", 512, 512, 64, 3), ("up8 (1x1)", 256, 128, 256, 1), ("up6 (1x1)", 2048, 1024, 64, 1)]:
    g = torch.randn(co, B, s, s, device=dev).half()
    x = torch.randn(k, ci, B, s, s, device=dev).half()
    dw = torch.zeros(co, k * k, ci, device=dev)
    for _ in range(2):
        check(lib.nrrt_conv_wgrad(P(g), co, P(x), ci, B, s, s, k, P(dw), ci, st))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        check(lib.nrrt_conv_wgrad(P(g), co, P(x), ci, B, s, s, k, P(dw), ci, st))
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    fl = 2.0 * B * s * s * co * ci * k * k
    tot_f += fl; tot_t += ms
    print("%-10s co %4d ci %4d %3dx%-3d k%d: %.3f ms  %.0f TFLOP/s" % (name, co, ci, s, s, k, ms, fl / ms / 1e9))
print("all: %.0f TFLOP/s" % (tot_f / tot_t / 1e9))
