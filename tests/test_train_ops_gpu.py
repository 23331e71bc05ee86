"""Training-step kernels (csrc/wgrad_tc.cu, csrc/train_ops.cu) one by one against plain PyTorch fp32 on the CPU."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from mgr_sampling_cnn_b200 import _lib
from mgr_sampling_cnn_b200._lib import check

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _st():
    return C.c_void_p(torch.cuda.current_stream(DEV).cuda_stream)


def channel_major(x_nhwc, mode=0, mask=None, copies=1):
    lib = _lib.load()
    n, h, w, c = x_nhwc.shape
    shape = {0: (copies, c, n, h, w), 1: (c, n, 2 * h, 2 * w), 2: (4, c, n, h // 2, w // 2)}[mode]
    out = torch.empty(shape, dtype=torch.float16, device=DEV)
    check(lib.nrrt_to_channel_major(_p(x_nhwc), c, _p(mask), c if mask is not None else 0, n, h, w, c, mode, copies, _p(out), _st()))
    return out[0] if (mode == 0 and copies == 1) else out


def test_channel_major_layouts():
    torch.manual_seed(0)
    x = torch.randn(2, 6, 64, 72, device=DEV).half()
    m = torch.randn(2, 6, 64, 72, device=DEV).half()
    assert torch.equal(channel_major(x), x.permute(3, 0, 1, 2).contiguous())
    assert torch.equal(channel_major(x, 0, m), (x * (m > 0)).permute(3, 0, 1, 2).contiguous())
    sh = channel_major(x, 0, None, 3)
    xc = x.permute(3, 0, 1, 2)
    assert torch.equal(sh[1], xc) and torch.equal(sh[0][..., 1:], xc[..., :-1]) and torch.equal(sh[2][..., :-1], xc[..., 1:])
    assert not sh[0][..., 0].any() and not sh[2][..., -1].any()
    st = torch.zeros(72, 2, 12, 128, dtype=torch.float16, device=DEV)
    st[:, :, ::2, ::2] = x.permute(3, 0, 1, 2)
    assert torch.equal(channel_major(x, 1), st)
    s2d = x.reshape(2, 3, 2, 32, 2, 72).permute(2, 4, 5, 0, 1, 3).reshape(4, 72, 2, 3, 32).contiguous()
    assert torch.equal(channel_major(x, 2), s2d)


@pytest.mark.parametrize("n,h,w,cout,cin,k", [(2, 16, 64, 128, 64, 3), (1, 8, 128, 64, 64, 3), (2, 12, 64, 256, 192, 3),
                                              (1, 16, 64, 128, 384, 3), (3, 8, 64, 512, 256, 1), (1, 70, 192, 64, 128, 3)])
def test_wgrad_matches_torch(n, h, w, cout, cin, k):
    torch.manual_seed(1)
    lib = _lib.load()
    g = (torch.randn(n, h, w, cout) * 0.5).half()
    x = torch.randn(n, h, w, cin).half()
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2).double(), (cout, cin, k, k), g.float().permute(0, 3, 1, 2).double(),
                                      padding=k // 2)                                        # [cout][cin][k][k]
    gt, xt = channel_major(g.to(DEV)), channel_major(x.to(DEV), copies=k)
    dw = torch.zeros(cout, k * k, cin, dtype=torch.float32, device=DEV)
    check(lib.nrrt_conv_wgrad(_p(gt), cout, _p(xt), cin, n, h, w, k, _p(dw), cin, _st()))
    got = dw.cpu().double().reshape(cout, k, k, cin).permute(0, 3, 1, 2)
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-3, "relative max error %g" % err


def test_bn_train_forward_backward():
    torch.manual_seed(2)
    lib = _lib.load()
    n, h, w, c = 2, 16, 24, 128
    M = n * h * w
    z = (torch.randn(n, h, w, c) * 2 + 0.5).half()
    res = torch.randn(n, h, w, c).half()
    dy = torch.randn(n, h, w, c).half()
    bn = torch.nn.BatchNorm2d(c)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.5, 0.5)
    gamma, beta = bn.weight.detach().clone().to(DEV), bn.bias.detach().clone().to(DEV)
    rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    zz = z.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    rr = res.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    bn.train()
    yref = F.relu(bn(zz) + rr)
    yref.backward(dy.float().permute(0, 3, 1, 2))
    zd, resd, dyd = z.to(DEV), res.to(DEV), dy.to(DEV)
    y = torch.empty_like(zd)
    mean, rstd, ss = torch.empty(c, device=DEV), torch.empty(c, device=DEV), torch.empty(2 * c, device=DEV)
    work = torch.empty(2 * c, dtype=torch.float64, device=DEV)
    check(lib.nrrt_bn_train_forward(_p(zd), c, M, c, _p(gamma), _p(beta), 1e-5, 0.1, _p(rm), _p(rv), _p(resd), c, 1, _p(y), c, _p(mean),
                                    _p(rstd), _p(ss), _p(work), _st()))
    assert (y.float().cpu() - yref.detach().permute(0, 2, 3, 1)).abs().max() < 2e-2
    assert (rm.cpu() - bn.running_mean).abs().max() < 1e-4 and (rv.cpu() - bn.running_var).abs().max() < 1e-3
    dz, gout = torch.empty_like(zd), torch.empty_like(zd)
    dgam, dbet = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
    check(lib.nrrt_bn_train_backward(_p(dyd), c, _p(y), c, _p(zd), c, M, c, _p(gamma), _p(mean), _p(rstd), _p(dz), c, _p(gout), c,
                                     _p(dgam), _p(dbet), _p(work), _st()))
    assert (dz.float().cpu() - zz.grad.permute(0, 2, 3, 1)).abs().max() < 2e-2
    assert (gout.float().cpu() - rr.grad.permute(0, 2, 3, 1)).abs().max() < 1e-2
    assert (dgam.cpu() - bn.weight.grad).abs().max() / bn.weight.grad.abs().max() < 5e-3
    assert (dbet.cpu() - bn.bias.grad).abs().max() / bn.bias.grad.abs().max() < 5e-3


def test_bce_head_adam_and_small_ops():
    torch.manual_seed(3)
    lib = _lib.load()
    # BCE
    M = 5000
    z = (torch.randn(M) * 4).to(DEV)
    masks = torch.randint(0, 256, (2, 2500), dtype=torch.uint8, device=DEV)
    t = torch.empty(2, 2500, device=DEV)
    check(lib.nrrt_mask_targets(_p(masks), 2, 2500, _p(t), _st()))
    mn = masks.cpu().numpy().astype(np.float64)
    tref = torch.from_numpy(np.stack([(m - m.min()) / (m.max() - m.min()) for m in mn])).float()
    assert torch.equal(t.cpu(), tref)
    dl, loss = torch.empty(M, device=DEV), torch.zeros(1, dtype=torch.float64, device=DEV)
    check(lib.nrrt_bce_with_logits(_p(z), None, _p(t), M, 1.0, None, _p(dl), _p(loss), _st()))
    zr = z.cpu().clone().requires_grad_(True)
    lref = F.binary_cross_entropy_with_logits(zr, tref.reshape(-1), reduction="sum")
    lref.backward()
    assert abs(loss.item() - lref.item()) / lref.item() < 1e-5
    assert (dl.cpu() - zr.grad).abs().max() < 1e-5
    # head backward
    P, c = 700, 64
    x = torch.randn(P, c).half()
    w = torch.randn(c)
    d = torch.randn(P)
    g = torch.empty(P, c, dtype=torch.float16, device=DEV)
    dw, db = torch.zeros(c, device=DEV), torch.zeros(1, device=DEV)
    dd, xd, wdev = d.to(DEV), x.to(DEV), w.to(DEV)     # (keep the device tensors alive: the call is asynchronous)
    check(lib.nrrt_head_backward(_p(dd), _p(xd), c, _p(wdev), P, c, _p(g), c, _p(dw), _p(db), _st()))
    assert (g.float().cpu() - (d[:, None] * w[None, :]) * (x > 0)).abs().max() < 1e-2
    assert (dw.cpu() - (d[:, None] * x.float()).sum(0)).abs().max() < 1e-2 and abs(db.item() - d.sum().item()) < 1e-3
    # Adam, two steps
    n = 3000
    p0, g1, g2 = torch.randn(n), torch.randn(n), torch.randn(n)
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=1e-3)
    for gg in (g1, g2):
        pr.grad = gg.clone()
        opt.step()
    pd, m, v = p0.to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for s, gg in enumerate((g1, g2), 1):
        gd = (gg * 4).to(DEV)
        check(lib.nrrt_adam_step(_p(pd), _p(gd), _p(m), _p(v), n, s, 1e-3, 0.9, 0.999, 1e-8, 0.25, None, _st()))
        torch.cuda.synchronize()
    assert (pd.cpu() - pr.detach()).abs().max() < 1e-6
    # the same two steps under the device-resident dynamic loss scale, with an overflowing step in between that must be skipped
    pd2, m2, v2 = p0.to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    state = torch.tensor([8.0, 0.0, 0.0, 0.0], device=DEV)
    bad = (g1 * 8).to(DEV); bad[17] = float("inf")
    for gg in ((g1 * 8).to(DEV), bad, (g2 * 4).to(DEV)):
        check(lib.nrrt_grad_check(_p(gg), n, _p(state), _st()))
        check(lib.nrrt_adam_step(_p(pd2), _p(gg), _p(m2), _p(v2), n, 1, 1e-3, 0.9, 0.999, 1e-8, 1.0, _p(state), _st()))
        check(lib.nrrt_loss_scale_update(_p(state), 1000, 65536.0, 1.0, _st()))
        torch.cuda.synchronize()
    assert state.cpu().tolist() == [4.0, 1.0, 0.0, 2.0]           # halved once, two steps taken
    assert (pd2.cpu() - pr.detach()).abs().max() < 1e-6
    # relu backward + bias sum, remaps, dgrad weight pack
    a, yv = torch.randn(3, 5, 7, 96).half().to(DEV), torch.randn(3, 5, 7, 96).half().to(DEV)
    gq, dbq = torch.empty_like(a), torch.zeros(96, device=DEV)
    check(lib.nrrt_relu_bwd_sum(_p(a), 96, _p(yv), 96, _p(gq), 96, _p(dbq), 105, 96, _st()))
    assert torch.equal(gq, a * (yv > 0)) and (dbq.cpu() - (a * (yv > 0)).float().sum((0, 1, 2)).cpu()).abs().max() < 1e-2
    xin = torch.randn(2, 8, 12, 16).half().to(DEV)
    o0 = torch.empty(2, 4, 6, 64, dtype=torch.float16, device=DEV)
    check(lib.nrrt_remap_channels_last(_p(xin), 16, _p(o0), 64, 2, 4, 6, 16, 0, _st()))
    assert torch.equal(o0, xin.reshape(2, 4, 2, 6, 2, 16).permute(0, 1, 3, 2, 4, 5).reshape(2, 4, 6, 64))
    o1 = torch.empty(2, 16, 24, 16, dtype=torch.float16, device=DEV)
    check(lib.nrrt_remap_channels_last(_p(xin), 16, _p(o1), 16, 2, 8, 12, 16, 1, _st()))
    z1 = torch.zeros_like(o1); z1[:, ::2, ::2] = xin
    assert torch.equal(o1, z1)
    wsrc = torch.randn(40, 9, 24, device=DEV)
    wd = torch.empty(64, 9, 64, dtype=torch.float16, device=DEV)
    check(lib.nrrt_pack_dgrad_weight(_p(wsrc), 40, 9, 24, 64, 64, _p(wd), _st()))
    wr = torch.zeros(64, 9, 64, device=DEV); wr[:24, :, :40] = wsrc.flip(1).permute(2, 1, 0)
    assert torch.equal(wd, wr.half())


def test_stem_wgrad_and_coords_backward():
    torch.manual_seed(4)
    lib = _lib.load()
    n, h, w = 2, 40, 50
    x = (torch.rand(n, 3, h, w) > 0.3).float()
    g = torch.zeros(n, h, w, 64).half(); g[..., :32] = (torch.randn(n, h, w, 32) * 0.3).half()
    ref = torch.nn.grad.conv2d_weight(x.double(), (32, 3, 3, 3), g[..., :32].float().permute(0, 3, 1, 2).double(), padding=1)
    dw, db = torch.zeros(32, 3, 9, device=DEV), torch.zeros(32, device=DEV)
    xd, gd = x.to(DEV), g.to(DEV)
    check(lib.nrrt_stem_wgrad(_p(xd), _p(gd), 64, n, 3, h, w, 32, _p(dw), _p(db), _st()))
    assert (dw.cpu().double().reshape(32, 3, 3, 3) - ref).abs().max() / ref.abs().max() < 1e-4
    assert (db.cpu() - g[..., :32].float().sum((0, 1, 2))).abs().max() < 1e-2
    # coordinate layer: conv3x3 + relu on the 2x2 grid, and the bilinear fill transpose
    B, cin, cout, H, W = 3, 64, 128, 32, 40
    conv = torch.nn.Conv2d(cin, cout, 3, padding=1)
    fin = torch.randn(B, cin, 2, 2, requires_grad=True)
    fout = F.relu(conv(fin))
    up = F.interpolate(fout, size=(H, W), mode="bilinear", align_corners=True)
    dcat = (torch.randn(B, H, W, cout) * 0.1).half()
    up.backward(dcat.float().permute(0, 3, 1, 2))
    dfeat = torch.zeros(B, 4, cout, device=DEV)
    dcd = dcat.to(DEV)
    check(lib.nrrt_coords_fill_backward(_p(dcd), cout, B, H, W, cout, _p(dfeat), _st()))
    f_out = fout.detach().permute(0, 2, 3, 1).reshape(B, 4, cout).contiguous().to(DEV)
    f_in = fin.detach().permute(0, 2, 3, 1).reshape(B, 4, cin).contiguous().to(DEV)
    wt = conv.weight.detach().permute(2, 3, 1, 0).reshape(9, cin, cout).contiguous().to(DEV)
    dW, dB, dfin = torch.zeros(9, cin, cout, device=DEV), torch.zeros(cout, device=DEV), torch.zeros(B, 4, cin, device=DEV)
    check(lib.nrrt_coords_layer_backward(_p(dfeat), _p(f_out), _p(f_in), _p(wt), B, cin, cout, _p(dW), _p(dB), _p(dfin), _st()))
    rW = conv.weight.grad.permute(2, 3, 1, 0).reshape(9, cin, cout)
    assert (dW.cpu() - rW).abs().max() / rW.abs().max() < 1e-3
    assert (dB.cpu() - conv.bias.grad).abs().max() / conv.bias.grad.abs().max() < 1e-3
    rin = fin.grad.permute(0, 2, 3, 1).reshape(B, 4, cin)
    assert (dfin.cpu() - rin).abs().max() / rin.abs().max() < 1e-3


@pytest.mark.parametrize("n,h,w,cout,cin,k", [(2, 16, 64, 128, 64, 3), (1, 8, 128, 64, 64, 3), (2, 12, 64, 256, 192, 3),
                                              (1, 16, 64, 128, 384, 3), (3, 8, 64, 512, 256, 1), (1, 70, 192, 64, 128, 3),
                                              (2, 16, 32, 128, 64, 3), (1, 24, 96, 64, 128, 3), (2, 8, 8, 256, 128, 1)])
def test_wgrad_nhwc_matches_torch(n, h, w, cout, cin, k):
    """nrrt_conv_wgrad_nhwc: MN-major operands straight from channels-last buffers, here channel SLICES of wider buffers."""
    torch.manual_seed(5)
    lib = _lib.load()
    gbuf = (torch.randn(n, h, w, cout + 64) * 0.5).half().to(DEV)
    xbuf = torch.randn(n, h, w, cin + 128).half().to(DEV)
    g, x = gbuf[..., 64:].float().cpu(), xbuf[..., 64:64 + cin].float().cpu()
    ref = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2).double(), (cout, cin, k, k), g.permute(0, 3, 1, 2).double(), padding=k // 2)
    dw = torch.zeros(cout, k * k, cin, dtype=torch.float32, device=DEV)
    check(lib.nrrt_conv_wgrad_nhwc(C.c_void_p(gbuf.data_ptr() + 128), cout + 64, cout, C.c_void_p(xbuf.data_ptr() + 128), cin + 128, cin,
                                   n, h, w, k, _p(dw), cin, _st()))
    got = dw.cpu().double().reshape(cout, k, k, cin).permute(0, 3, 1, 2)
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-3, "relative max error %g" % err


@pytest.mark.parametrize("gh,gw", [(2, 2), (2, 3)])
def test_coords_branch_does_not_depend_on_the_batch_size(gh, gw):
    """nrrt_coords_branch picks a different kernel for small batches (<= 64 tasks): the features of a task must be the same
    bits either way, or a sharded planning run would differ from the single-GPU run."""
    torch.manual_seed(6)
    lib = _lib.load()
    chans = [1, 64, 128, 256, 512]
    ws = [(torch.randn(9, chans[i], chans[i + 1], device=DEV) * (0.5 / np.sqrt(9 * chans[i]))).contiguous() for i in range(4)]
    bs = [(torch.randn(chans[i + 1], device=DEV) * 0.1).contiguous() for i in range(4)]
    wp = (C.c_void_p * 4)(*[t.data_ptr() for t in ws])
    bp = (C.c_void_p * 4)(*[t.data_ptr() for t in bs])
    coords = (torch.rand(100, 1, gh, gw, device=DEV) * 500).contiguous()

    def run(n):
        feats = [torch.zeros(n, gh * gw, c, device=DEV) for c in chans[1:]]
        fp = (C.c_void_p * 4)(*[t.data_ptr() for t in feats])
        check(lib.nrrt_coords_branch(_p(coords), n, gh, gw, wp, bp, fp, _st()))
        torch.cuda.synchronize()
        return feats

    big = run(100)
    for n in (1, 3, 64):
        small = run(n)
        for a, b in zip(small, big):
            assert torch.equal(a, b[:n]), "batch of %d differs from the batch of 100" % n
    assert float(big[3].abs().max()) > 0
