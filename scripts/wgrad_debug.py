"""Scratch: bisect the wgrad kernel with the NRRT_WG_DBG switches (separate debug .so, one subprocess per setting)."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "gpurun_out", "libwgdbg.so")
if len(sys.argv) == 1:
    cs = os.path.join(ROOT, "mgr_sampling_cnn_b200", "csrc")
    subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
                           "-DNRRT_DEBUG_SWITCHES", "-I", os.path.join(ROOT, "include"), "-I", cs, "-shared", "-o", SO,
                           os.path.join(cs, "wgrad_tc.cu"), os.path.join(cs, "abi.cu"), "-lcudart", "-lcuda"])
    for dbg in (0,):
        r = subprocess.run([sys.executable, __file__, str(dbg)], env=dict(os.environ, NRRT_WG_DBG=str(dbg)), capture_output=True, text=True, timeout=120)
        print("dbg", dbg, "rc", r.returncode, (r.stdout + r.stderr).strip().splitlines()[-1:])
    sys.exit(0)
import torch
lib = C.CDLL(SO)
dev = torch.device("cuda:0")
n, h, w, cout, cin, k = 2, 16, 64, 128, 64, 3
gt = torch.randn(cout, n, h, w, device=dev).half()
x1 = torch.randn(cin, n, h, w, device=dev).half()
xt = torch.zeros(3, cin, n, h, w, device=dev).half(); xt[1] = x1; xt[0][..., 1:] = x1[..., :-1]; xt[2][..., :-1] = x1[..., 1:]
dw = torch.zeros(cout, k * k, cin, device=dev)
P = lambda t: C.c_void_p(t.data_ptr())
rc = lib.nrrt_conv_wgrad(P(gt), cout, P(xt), cin, n, h, w, k, P(dw), cin, C.c_void_p(torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
ref = torch.nn.grad.conv2d_weight(x1.float().permute(1, 0, 2, 3), (cout, cin, k, k), gt.float().permute(1, 0, 2, 3), padding=1)
got = dw.reshape(cout, k, k, cin).permute(0, 3, 1, 2)
print("rc", rc, "ok; max err", float((got - ref).abs().max()), "ref max", float(ref.abs().max()))
