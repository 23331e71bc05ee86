"""GPU: the tcgen05 implicit-GEMM convolution layer (nrrt_conv_layer) against torch fp32 convolutions on the same
fp16-rounded operands, then the whole UNet forward against the fp32 oracle (tolerance <= 1e-2 abs, BASELINE north star)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from mgr_sampling_cnn_b200 import _lib, cnn
from oracle import cnn_oracle
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _cl(x):  # NC(D)HW fp32 -> channels-last fp16 [B, D, H, W, C] (D = 1 in 2D)
    if x.dim() == 4:
        return x.permute(0, 2, 3, 1).unsqueeze(1).contiguous().half()
    return x.permute(0, 2, 3, 4, 1).contiguous().half()


def _from_cl(t, dims):  # [B, D, H, W, C] fp16 -> NC(D)HW fp32
    t = t.float()
    return t[:, 0].permute(0, 3, 1, 2) if dims == 2 else t.permute(0, 4, 1, 2, 3)


def _run_layer(dims, kind, x, w, scale, shift, relu, dev, res=None, out_ctot=None, out_off=0, in_ctot=None, in_off=0,
               post=None):
    cin = x.shape[1]
    if kind == cnn.CONVT:
        cout = w.shape[1]
        wp = cnn._pack_convT_weight(w, dev)
        groups = 2 ** dims
        scale, shift = scale.repeat(groups), shift.repeat(groups)
        up = 2
    else:
        cout = w.shape[0]
        wp = cnn._pack_conv_weight(w, cin, dev)
        up = 1
    layer = cnn._Layer(dims, kind, cin, cout, wp, scale.float().contiguous(), shift.float().contiguous(), relu, post)
    xin = _cl(x)
    if in_ctot:
        big = torch.randn(xin.shape[:-1] + (in_ctot,), device=dev).half()
        big[..., in_off:in_off + cin] = xin
        xin = big
    grid = tuple(xin.shape[1:4])
    stride = 2 if kind in (cnn.K3S2, cnn.K1S2) else 1
    og = tuple((g if (dims == 3 or i > 0) else 1) for i, g in enumerate(grid))
    og = tuple(((g - 1) // stride + 1) * up if (dims == 3 or i > 0) else 1 for i, g in enumerate(og))
    out = torch.full((x.shape[0],) + og + (out_ctot or cout,), 7.0, dtype=torch.float16, device=dev)
    resb = _cl(res) if res is not None else None
    layer.run(_lib.load(), None, xin, in_off, grid, out, out_off, resb, 0)
    torch.cuda.synchronize()
    return out


def _ref(dims, kind, x, w, scale, shift, relu, res=None, post=None):
    xh, wh = x.half().float(), w.half().float()
    conv = F.conv2d if dims == 2 else F.conv3d
    if kind == cnn.CONVT:
        y = (F.conv_transpose2d if dims == 2 else F.conv_transpose3d)(xh, wh, stride=2)
    elif kind == cnn.K3S1:
        y = conv(xh, wh, padding=1)
    elif kind == cnn.K3S2:
        y = conv(xh, wh, stride=2, padding=1)
    elif kind == cnn.K1S2:
        y = conv(xh, wh, stride=2)
    else:
        y = conv(xh, wh)
    shp = (1, -1) + (1,) * dims
    y = y * scale.view(shp) + shift.view(shp)
    if res is not None:
        y = y + res.half().float()
    if relu:
        y = F.relu(y)
    if post is not None:
        y = F.relu(y * post[0].view(shp) + post[1].view(shp))
    return y


def _check(out, ref, dims, cout, off=0):
    got = _from_cl(out[..., off:off + cout], dims)
    err = (got - ref).abs().max().item()
    tol = 2e-3 * max(ref.abs().max().item(), 1.0) + 1e-3
    assert err <= tol, f"max abs err {err} > {tol}"


CASES = [
    # dims, kind, B, cin, cout, spatial
    (2, cnn.K1S1, 2, 64, 64, (16, 16)),      # plain GEMM, one K stage
    (2, cnn.K1S1, 1, 256, 128, (24, 40)),    # 4 K stages, N = 128, ragged tiles
    (2, cnn.K3S1, 2, 64, 64, (16, 16)),      # slab kernel (N <= 128), exact 16x16 tiles
    (2, cnn.K3S1, 2, 64, 64, (24, 40)),      # slab kernel, ragged tiles
    (2, cnn.K3S1, 1, 128, 128, (16, 48)),    # slab kernel, two K chunks, N = 128
    (2, cnn.K3S1, 1, 192, 128, (40, 24)),    # slab kernel, three K chunks
    (2, cnn.K3S1, 1, 128, 256, (40, 24)),
    (2, cnn.K3S1, 1, 192, 512, (16, 32)),    # two N tiles of 256
    (2, cnn.K3S2, 2, 64, 128, (32, 32)),
    (2, cnn.K1S2, 2, 64, 128, (32, 32)),
    (2, cnn.CONVT, 2, 128, 64, (16, 16)),    # 4 groups x 64 = one N tile of 256
    (2, cnn.CONVT, 1, 256, 128, (8, 24)),    # 4 groups x 128 = two N tiles
    (3, cnn.K1S1, 1, 64, 64, (8, 8, 8)),
    (3, cnn.K3S1, 1, 64, 64, (8, 8, 16)),
    (3, cnn.K3S1, 2, 128, 128, (10, 10, 10)),  # 5x5x5 boxes (125 rows)
    (3, cnn.K3S1, 2, 64, 64, (4, 16, 32)),     # 3D slab kernel (16x16 tiles per depth slice), N = 64
    (3, cnn.K3S1, 1, 128, 128, (3, 32, 16)),   # 3D slab kernel, two K chunks, N = 128 (TMA-store epilogue)
    (3, cnn.K3S2, 1, 64, 128, (16, 16, 16)),
    (3, cnn.K1S2, 1, 64, 128, (16, 16, 16)),
    (3, cnn.CONVT, 1, 128, 64, (4, 8, 8)),     # 8 groups x 64 = two N tiles of 256
]


@pytest.mark.parametrize("dims,kind,B,cin,cout,sp", CASES)
def test_conv_layer_matches_torch(dims, kind, B, cin, cout, sp, cuda_device):
    g = torch.Generator(device="cpu").manual_seed(cin * 7 + cout + kind)
    dev = cuda_device
    x = torch.randn((B, cin) + sp, generator=g).to(dev)
    k = 1 if kind in (cnn.K1S1, cnn.K1S2) else (2 if kind == cnn.CONVT else 3)
    wshape = ((cin, cout) if kind == cnn.CONVT else (cout, cin)) + (k,) * dims
    w = (torch.randn(wshape, generator=g) / np.sqrt(cin * k ** dims)).to(dev)
    scale = (torch.rand(cout, generator=g) + 0.5).to(dev)
    shift = torch.randn(cout, generator=g).to(dev)
    out = _run_layer(dims, kind, x, w, scale, shift, True, dev)
    _check(out, _ref(dims, kind, x, w, scale, shift, True), dims, cout)


def test_conv_layer_residual_slices_and_post_affine(cuda_device):
    dev = cuda_device
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randn((2, 64, 16, 16), generator=g).to(dev)
    w = (torch.randn((128, 64, 3, 3), generator=g) / 24).to(dev)
    scale, shift = (torch.rand(128, generator=g) + 0.5).to(dev), torch.randn(128, generator=g).to(dev)
    res = torch.randn((2, 128, 16, 16), generator=g).to(dev)
    post = ((torch.rand(128, generator=g) + 0.5).to(dev), torch.randn(128, generator=g).to(dev))
    # input read from channels [64,128) of a 192-wide buffer; output written to channels [32,160) of a 192-wide buffer
    out = _run_layer(2, cnn.K3S1, x, w, scale, shift, True, dev, res=res, out_ctot=192, out_off=32, in_ctot=192, in_off=64,
                     post=post)
    _check(out, _ref(2, cnn.K3S1, x, w, scale, shift, True, res, post), 2, 128, off=32)
    assert (out[..., :32] == 7.0).all() and (out[..., 160:] == 7.0).all()  # neighbours of the slice untouched


def test_slab_kernel_equals_per_tap_kernel_and_fused_head(cuda_device):
    """The slab (halo-reuse) kernel and the generic per-tap kernel run the same MMAs in a different order of K: equal
    up to fp32 accumulation order; the fused 1x1 head equals a separate dot product over the stored activation."""
    dev = cuda_device
    g = torch.Generator(device="cpu").manual_seed(3)
    x = torch.randn((2, 128, 40, 56), generator=g).to(dev)
    w = (torch.randn((64, 128, 3, 3), generator=g) / 34).to(dev)
    scale, shift = (torch.rand(64, generator=g) + 0.5).to(dev), torch.randn(64, generator=g).to(dev)
    a = _run_layer(2, cnn.K3S1, x, w, scale, shift, True, dev)
    cnn._Layer.force_per_tap = True
    try:
        b = _run_layer(2, cnn.K3S1, x, w, scale, shift, True, dev)
    finally:
        cnn._Layer.force_per_tap = False
    assert (a.float() - b.float()).abs().max().item() <= 4e-3
    _check(a, _ref(2, cnn.K3S1, x, w, scale, shift, True), 2, 64)
    # fused head
    hw = torch.randn(64, generator=g).to(dev)
    layer = cnn._Layer(2, cnn.K3S1, 128, 64, cnn._pack_conv_weight(w, 128, dev), scale.contiguous(), shift.contiguous(), True)
    logits = torch.zeros((2, 40, 56), dtype=torch.float32, device=dev)
    layer.run(_lib.load(), None, _cl(x), 0, (1, 40, 56), None, 0, head=(hw, 0.25, logits))
    torch.cuda.synchronize()
    ref = (_ref(2, cnn.K3S1, x, w, scale, shift, True) * hw.view(1, -1, 1, 1)).sum(1) + 0.25
    assert (logits - ref).abs().max().item() <= 2e-3 * max(ref.abs().max().item(), 1.0)


@pytest.mark.parametrize("cout,sp", [(128, (24, 40)), (256, (16, 24))])  # slab kernel / per-tap kernel
def test_k_split_two_sources_with_sample_index(cout, sp, cuda_device):
    """conv(torch.cat((a, b[index]), 1)) without the concatenated buffer: source 2 is indexed per sample."""
    dev = cuda_device
    g = torch.Generator(device="cpu").manual_seed(11)
    a = torch.randn((5, 64) + sp, generator=g).to(dev)          # per-task part
    b = torch.randn((2, 128) + sp, generator=g).to(dev)         # per-map part
    index = torch.tensor([1, 0, 0, 1, 1], dtype=torch.int32, device=dev)
    w = (torch.randn((cout, 192, 3, 3), generator=g) / 41).to(dev)
    scale, shift = (torch.rand(cout, generator=g) + 0.5).to(dev), torch.randn(cout, generator=g).to(dev)
    layer = cnn._Layer(2, cnn.K3S1, 64, cout, cnn._pack_conv_weight(w, 192, dev), scale.contiguous(), shift.contiguous(), True,
                       cin_real=192)
    layer.cin2 = 128
    out = torch.zeros((5, 1) + sp + (cout,), dtype=torch.float16, device=dev)
    layer.run(_lib.load(), None, _cl(a), 0, (1,) + sp, out, 0, x2=_cl(b), idx2=index)
    torch.cuda.synchronize()
    ref = _ref(2, cnn.K3S1, torch.cat((a, b[index.long()]), 1), w, scale, shift, True)
    _check(out, ref, 2, cout)


def test_indexed_residual_before_relu(cuda_device):
    """out = relu(conv(x) + res[index]): the per-map partial sum of a decoder conv enters the per-task layer as a residual."""
    dev = cuda_device
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.randn((5, 64, 24, 40), generator=g).to(dev)
    res = torch.randn((2, 128, 24, 40), generator=g).to(dev)
    index = torch.tensor([1, 0, 0, 1, 1], dtype=torch.int32, device=dev)
    w = (torch.randn((128, 64, 3, 3), generator=g) / 24).to(dev)
    scale, shift = torch.ones(128, device=dev), torch.randn(128, generator=g).to(dev)
    layer = cnn._Layer(2, cnn.K3S1, 64, 128, cnn._pack_conv_weight(w, 64, dev), scale, shift.contiguous(), True)
    out = torch.zeros((5, 1, 24, 40, 128), dtype=torch.float16, device=dev)
    layer.run(_lib.load(), None, _cl(x), 0, (1, 24, 40), out, 0, _cl(res), 0, res_idx=index)
    torch.cuda.synchronize()
    _check(out, _ref(2, cnn.K3S1, x, w, scale, shift, True, res[index.long()]), 2, 128)


def test_poly16_matches_conv_over_the_coordinate_field(cuda_device):
    """nrrt_poly16 + nrrt_poly_relu against the definition: a 3x3 conv over (ConvTranspose polynomial field | bilinear
    coordinate channels) evaluated densely with torch in fp64, added to an indexed per-map pre-activation, ReLU."""
    import ctypes as C
    dev = cuda_device
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(9)
    B, Cm, Cout, H3, W3 = 3, 32, 128, 6, 10
    H, W = 2 * H3, 2 * W3
    e_up = torch.randn((B, 4, 4, Cm), generator=g).to(dev)            # ConvTranspose polynomial per output offset
    e_conv = torch.randn((B, 9, 4, Cout), generator=g).to(dev)        # the conv's own coordinate polynomial (9 border cases)
    w = (torch.randn((Cout, Cm, 3, 3), generator=g) / 17).to(dev)
    bias = torch.randn(Cout, generator=g).to(dev)
    pre = torch.randn((2, H, W, Cout), generator=g).to(dev).half()
    index = torch.tensor([1, 0, 1], dtype=torch.int32, device=dev)
    w_taps = w.permute(1, 2, 3, 0).reshape(Cm, 9 * Cout).contiguous()
    work = torch.empty(B * 16 * 9 * Cout, dtype=torch.float32, device=dev)
    P16 = torch.empty((B, 16, 4, Cout), dtype=torch.float32, device=dev)
    out = torch.empty((B, H, W, Cout), dtype=torch.float16, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    _lib.check(lib.nrrt_poly16(p(e_up), p(w_taps), p(e_conv), p(bias), B, Cm, Cout, H3, W3, p(work), p(P16), None))
    _lib.check(lib.nrrt_poly_relu(p(pre), p(index), p(P16), p(out), B, H, W, Cout, None))
    torch.cuda.synchronize()
    # dense definition in fp64
    Y, X = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    s, t = (Y >> 1).double() / (H3 - 1), (X >> 1).double() / (W3 - 1)
    grp = ((Y & 1) * 2 + (X & 1)).long()
    eu = e_up.double()[:, grp]                                         # [B, H, W, 4, Cm]
    U = eu[..., 0, :] + s[..., None] * eu[..., 1, :] + t[..., None] * eu[..., 2, :] + (s * t)[..., None] * eu[..., 3, :]
    q = F.conv2d(U.permute(0, 3, 1, 2), w.double(), padding=1).permute(0, 2, 3, 1)
    ty, tx = Y.double() / (H - 1), X.double() / (W - 1)
    rc = torch.where(Y == 0, 0, torch.where(Y == H - 1, 2, 1))
    cc = torch.where(X == 0, 0, torch.where(X == W - 1, 2, 1))
    ec = e_conv.double()[:, (rc * 3 + cc).long()]                      # [B, H, W, 4, Cout]
    poly = ec[..., 0, :] + ty[..., None] * ec[..., 1, :] + tx[..., None] * ec[..., 2, :] + (ty * tx)[..., None] * ec[..., 3, :]
    ref = torch.relu(pre.double()[index.long()] + q + poly + bias.double())
    err = (out.double() - ref).abs().max().item()
    assert err <= 2e-3 * max(ref.abs().max().item(), 1.0), f"max abs err {err}"


@pytest.mark.parametrize("cout,sp", [(128, (40, 56)), (128, (16, 16)), (256, (40, 56)), (256, (24, 32))])
def test_polynomial_term_through_the_tensor_core(cout, sp, cuda_device):
    """Residual-carrying 3x3 layers add the interior coordinate polynomial with one extra MMA per tile (fp16 hi + lo
    coefficients, exact fp16 basis) and patch the image frame in the epilogue: equal to the all-epilogue evaluation and
    to the dense definition."""
    dev = cuda_device
    g = torch.Generator(device="cpu").manual_seed(21)
    B, cin = 3, 64
    x = torch.randn((B, cin) + sp, generator=g).to(dev)
    res = torch.randn((2, cout) + sp, generator=g).to(dev)
    index = torch.tensor([1, 0, 1], dtype=torch.int32, device=dev)
    w = (torch.randn((cout, cin, 3, 3), generator=g) / 24).to(dev)
    scale, shift = torch.ones(cout, device=dev), torch.randn(cout, generator=g).to(dev)
    poly = (torch.randn((B, 9, 4, cout), generator=g) * torch.tensor([100.0, 60.0, 60.0, 30.0]).view(1, 1, 4, 1)).to(dev).contiguous()
    layer = cnn._Layer(2, cnn.K3S1, cin, cout, cnn._pack_conv_weight(w, cin, dev), scale, shift.contiguous(), True)
    outs = []
    for no_mma in (False, True):
        layer.no_poly_mma = no_mma
        out = torch.zeros((B, 1) + sp + (cout,), dtype=torch.float16, device=dev)
        layer.run(_lib.load(), None, _cl(x), 0, (1,) + sp, out, 0, _cl(res), 0, poly=poly, res_idx=index)
        torch.cuda.synchronize()
        outs.append(out.float())
    H, W = sp
    Y, X = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    ty, tx = Y.float() / (H - 1), X.float() / (W - 1)
    rc = torch.where(Y == 0, 0, torch.where(Y == H - 1, 2, 1))
    cc = torch.where(X == 0, 0, torch.where(X == W - 1, 2, 1))
    e = poly[:, (rc * 3 + cc).long()]                                   # [B, H, W, 4, cout]
    pv = e[..., 0, :] + ty[..., None] * e[..., 1, :] + tx[..., None] * e[..., 2, :] + (ty * tx)[..., None] * e[..., 3, :]
    ref = F.relu(_ref(2, cnn.K3S1, x, w, scale, shift, False) + res[index.long()].half().float() + pv.permute(0, 3, 1, 2))
    tol = 2e-3 * max(ref.abs().max().item(), 1.0) + 1e-3
    for out in outs:
        assert (_from_cl(out, 2) - ref).abs().max().item() <= tol
    assert (outs[0] - outs[1]).abs().max().item() <= tol


@pytest.mark.parametrize("cin,cout,sp", [(128, 128, (32, 48)), (64, 128, (16, 16)), (256, 256, (40, 24)), (128, 256, (17, 33))])
def test_folded_upconv_conv_layer(cin, cout, sp, cuda_device):
    """NRRT_CONV_UPFOLD: ConvTranspose2d(k=2, s=2) + bias composed with the 3x3 conv that consumes it (train.py:196-201), with
    the indexed per-map residual, the per-parity polynomial (nrrt_poly_fold) and the ReLU, against the two torch layers.
    (16, 16): the polynomial stays in the epilogue; ragged grids: partial tiles; cout 256: two N halves."""
    import ctypes as C
    dev = cuda_device
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(cin + cout + sp[0])
    B, cup = 3, 64
    H, W = sp
    x = torch.randn((B, cin) + sp, generator=g).to(dev)
    wt = (torch.randn((cin, cup, 2, 2), generator=g) / np.sqrt(cin)).to(dev)
    bu = torch.randn(cup, generator=g).to(dev)
    wc = (torch.randn((cout, cup, 3, 3), generator=g) / 24).to(dev)
    shift = torch.randn(cout, generator=g).to(dev).contiguous()
    res = torch.randn((2, cout, 2 * H, 2 * W), generator=g).to(dev)
    index = torch.tensor([1, 0, 1], dtype=torch.int32, device=dev)
    poly = (torch.randn((B, 9, 4, cout), generator=g) * torch.tensor([100.0, 60.0, 60.0, 30.0]).view(1, 1, 4, 1)).to(dev).contiguous()
    wf, bc = cnn._fold_weights(wt, bu, wc, dev)
    Ef = torch.empty((B, 4, 9, 4, cout), dtype=torch.float32, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    _lib.check(lib.nrrt_poly_fold(p(poly), p(bc), B, cout, H, W, p(Ef), None))
    layer = cnn._Layer(2, cnn.UPFOLD, cin, cout, wf, torch.ones(cout, device=dev), shift, True)
    outs = []
    # group 2 of 3 samples: tile-major schedule, ragged last group; res_mma: residual accumulated by the tensor core
    for no_mma, group, res_mma in ((False, 0, False), (True, 0, False), (False, 2, False), (False, 0, True), (True, 2, True)):
        layer.no_poly_mma, layer.res_mma = no_mma, res_mma
        out = torch.full((B, 1, 2 * H, 2 * W, cout), 7.0, dtype=torch.float16, device=dev)
        layer.run(lib, None, _cl(x), 0, (1, H, W), out, 0, _cl(res), 0, poly=Ef, res_idx=index, res_group=group)
        torch.cuda.synchronize()
        outs.append(_from_cl(out, 2))
    assert torch.equal(outs[0], outs[2])  # the schedule does not change results
    Ho, Wo = 2 * H, 2 * W
    Y, X = torch.meshgrid(torch.arange(Ho, device=dev), torch.arange(Wo, device=dev), indexing="ij")
    ty, tx = Y.float() / (Ho - 1), X.float() / (Wo - 1)
    rc = torch.where(Y == 0, 0, torch.where(Y == Ho - 1, 2, 1))
    cc = torch.where(X == 0, 0, torch.where(X == Wo - 1, 2, 1))
    e = poly[:, (rc * 3 + cc).long()]
    pv = e[..., 0, :] + ty[..., None] * e[..., 1, :] + tx[..., None] * e[..., 2, :] + (ty * tx)[..., None] * e[..., 3, :]
    up = F.conv_transpose2d(x.half().float(), wt, bu, stride=2)
    ref = F.relu(F.conv2d(up, wc, shift, padding=1) + res[index.long()].half().float() + pv.permute(0, 3, 1, 2))
    tol = 3e-3 * max(ref.abs().max().item(), 1.0) + 2e-3
    for out in outs:
        assert (out - ref).abs().max().item() <= tol, (out - ref).abs().max().item()


def test_folded_decoder_equals_separate_layers(cuda_device):
    """The 2D forward with upconv7/conv7.0 and upconv8/conv8.0 folded agrees with the same forward through the separate
    ConvTranspose and conv layers (and both with the fp32 oracle within the 1e-2 tolerance of the north star)."""
    from mgr_sampling_cnn_b200 import train, workload
    wl = workload.make_2d(3, size=128)
    torch.manual_seed(0)
    model = train.UNet_cooler().eval().to(cuda_device)
    x = torch.from_numpy((wl.occ_maps[wl.task_to_map] != 0).astype(np.float32))[:, None].repeat(1, 3, 1, 1).to(cuda_device)
    coords = torch.from_numpy(wl.coords).to(cuda_device)
    a = model(x, coords).clone()
    try:
        cnn.UNetPlan.fold = False
        b = model(x, coords).clone()
    finally:
        cnn.UNetPlan.fold = True
    with torch.no_grad():
        ref = cnn_oracle.forward({k: v.float() for k, v in model.state_dict().items()}, x, coords, 2)
    torch.cuda.synchronize()
    assert (a - ref).abs().max().item() <= 1e-2 and (b - ref).abs().max().item() <= 1e-2
    assert (a - b).abs().max().item() <= 5e-3


def test_stem_from_binary_maps_is_a_table_lookup(cuda_device):
    """nrrt_stem_conv_maps in 2D (every input channel = the {0,1} occupancy map, train.py:41-44) runs as a 512-entry table
    lookup; it must equal the generic first conv on the materialised (B, 3, H, W) tensor (ragged pixel count: 3*20*28)."""
    import ctypes as C
    dev = cuda_device
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(5)
    n_maps, H, W, cin, cout, pad = 4, 20, 28, 3, 32, 64
    occ = (torch.rand((n_maps, H, W), generator=g) < 0.6).to(torch.uint8).mul(255).to(dev)
    t2m = torch.tensor([2, 0, 3], dtype=torch.int32, device=dev)
    w = torch.randn((cout, cin * 9), generator=g).to(dev).contiguous()
    b = torch.randn(cout, generator=g).to(dev)
    x = (occ[t2m.long()] != 0).float()[:, None].repeat(1, cin, 1, 1).contiguous()
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    o1 = torch.full((3, 1, H, W, pad), 7.0, dtype=torch.float16, device=dev)
    o2 = torch.full_like(o1, 9.0)
    _lib.check(lib.nrrt_stem_conv_maps(2, p(occ), p(t2m), 3, cin, 1, H, W, p(w), p(b), cout, pad, p(o1), None))
    _lib.check(lib.nrrt_stem_conv(2, p(x), 3, cin, 1, H, W, p(w), p(b), cout, pad, p(o2), None))
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(x, w.view(cout, cin, 3, 3), b, padding=1))
    assert (o1[..., cout:] == 0).all() and (o2[..., cout:] == 0).all()
    assert (_from_cl(o1[..., :cout], 2) - ref).abs().max().item() <= 4e-3 * max(ref.abs().max().item(), 1.0)
    assert (o1.float() - o2.float()).abs().max().item() <= 4e-3 * max(ref.abs().max().item(), 1.0)


def test_3d_field_in_the_epilogue_equals_the_post_pass(cuda_device):
    """conv8.0 of the 3D decoder adds its coordinate field and the ReLU in the epilogue (one depth slice per tile: A + th * B
    per h-class, staged per tile; frame columns from global memory); same logits as the nrrt_poly_relu_3d post-pass and
    within 1e-2 of the fp32 oracle.  32^3: four 16x16 tiles per slice, the h-classes split between them."""
    from mgr_sampling_cnn_b200 import ThreeD_train, workload
    wl = workload.make_3d(3, size=32)
    torch.manual_seed(0)
    model = ThreeD_train.ThreeD_UNet_cooler().eval().to(cuda_device)
    x = torch.from_numpy((wl.occ_maps[wl.task_to_map] != 0).astype(np.float32)).permute(0, 3, 1, 2).contiguous().to(cuda_device)
    coords = torch.from_numpy(wl.coords).to(cuda_device)
    a = model(x, coords).clone()
    try:
        cnn.UNetPlan.fuse_poly3 = False
        b = model(x, coords).clone()
    finally:
        cnn.UNetPlan.fuse_poly3 = True
    with torch.no_grad():
        ref = cnn_oracle.forward({k: v.float() for k, v in model.state_dict().items()}, x, coords, 3)
    torch.cuda.synchronize()
    assert (a - ref).abs().max().item() <= 1e-2 and (b - ref).abs().max().item() <= 1e-2
    assert (a - b).abs().max().item() <= 5e-3


def test_3d_slab_kernel_residual_and_head(cuda_device):
    """3D 3x3x3 layers on the slab kernel: indexed residual (TMA-staged) and the fused 1x1x1 head."""
    dev = cuda_device
    g = torch.Generator(device="cpu").manual_seed(41)
    sp = (5, 16, 16)
    x = torch.randn((3, 64) + sp, generator=g).to(dev)
    res = torch.randn((2, 128) + sp, generator=g).to(dev)
    index = torch.tensor([1, 0, 1], dtype=torch.int32, device=dev)
    w = (torch.randn((128, 64, 3, 3, 3), generator=g) / 41).to(dev)
    scale, shift = torch.ones(128, device=dev), torch.randn(128, generator=g).to(dev)
    layer = cnn._Layer(3, cnn.K3S1, 64, 128, cnn._pack_conv_weight(w, 64, dev), scale, shift.contiguous(), True)
    out = torch.zeros((3,) + sp + (128,), dtype=torch.float16, device=dev)
    layer.run(_lib.load(), None, _cl(x), 0, sp, out, 0, _cl(res), 0, res_idx=index)
    torch.cuda.synchronize()
    _check(out, _ref(3, cnn.K3S1, x, w, scale, shift, True, res[index.long()]), 3, 128)
    w2 = (torch.randn((64, 64, 3, 3, 3), generator=g) / 41).to(dev)
    sc2, sh2 = (torch.rand(64, generator=g) + 0.5).to(dev), torch.randn(64, generator=g).to(dev)
    hw = torch.randn(64, generator=g).to(dev)
    layer2 = cnn._Layer(3, cnn.K3S1, 64, 64, cnn._pack_conv_weight(w2, 64, dev), sc2.contiguous(), sh2.contiguous(), True)
    logits = torch.zeros((3,) + sp, dtype=torch.float32, device=dev)
    layer2.run(_lib.load(), None, _cl(x), 0, sp, None, 0, head=(hw, -0.5, logits))
    torch.cuda.synchronize()
    ref = (_ref(3, cnn.K3S1, x, w2, sc2, sh2, True) * hw.view(1, -1, 1, 1, 1)).sum(1) - 0.5
    assert (logits - ref).abs().max().item() <= 2e-3 * max(ref.abs().max().item(), 1.0)


@pytest.mark.parametrize("cases,n_uv", [(9, 4), (4, 1)])
def test_coords_poly_gemm_equals_direct_kernel(cases, n_uv, cuda_device):
    """The epilogue-polynomial coefficients computed as one fp32 GEMM over the batch (large batches) equal the direct kernel."""
    import ctypes as C
    dev = cuda_device
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(31)
    B, Cc, Cout = 70, 64, 128
    feat = (torch.randn((B, 4, Cc), generator=g) * 50).to(dev)
    mom = torch.randn((cases, n_uv, Cc, Cout), generator=g).to(dev)
    mom_t = mom.permute(2, 0, 1, 3).reshape(Cc, -1).contiguous()
    a, b = 1.0 / 127, 1.0 / 255
    E1 = torch.empty((B, cases, 4, Cout), dtype=torch.float32, device=dev)
    E2 = torch.empty_like(E1)
    work = torch.empty(B * 4 * (Cc + cases * n_uv * Cout), dtype=torch.float32, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    _lib.check(lib.nrrt_coords_poly(p(feat), B, Cc, p(mom), cases, n_uv, Cout, a, b, p(E1), None))
    _lib.check(lib.nrrt_coords_poly_gemm(p(feat), B, Cc, p(mom_t), cases, n_uv, Cout, a, b, p(work), p(E2), None))
    torch.cuda.synchronize()
    assert (E1 - E2).abs().max().item() <= 1e-4 * max(E1.abs().max().item(), 1.0)


@pytest.mark.parametrize("three_d", [False, True])
@pytest.mark.parametrize("bn", [False, True])
def test_unet_forward_matches_reference_golden(three_d, bn, cuda_device):
    """The golden logits were produced by the UNMODIFIED reference model on CPU fp32 (oracle/gen_golden.py)."""
    from mgr_sampling_cnn_b200 import ThreeD_train, train
    g = H.load(("cnn3d_seed0" if three_d else "cnn2d_seed0") + ("_bn" if bn else ""))
    torch.manual_seed(0)
    model = (ThreeD_train.ThreeD_UNet_cooler() if three_d else train.UNet_cooler()).eval()
    if bn:
        cnn_oracle.randomize_bn(model, 0)
    model = model.to(cuda_device)
    y = model(torch.from_numpy(g["x"]).to(cuda_device), torch.from_numpy(g["coords"]).to(cuda_device))
    torch.cuda.synchronize()
    assert y.shape == g["logits"].shape and y.dtype == torch.float32
    err = np.abs(y.cpu().numpy() - g["logits"]).max()
    assert err <= 1e-2, f"max abs err {err}"


def test_unet_forward_full_size_2d_vs_fp32_oracle(cuda_device):
    """512x512 maps from the repo's generator, pixel coordinates as the reference feeds them (un-normalised)."""
    from mgr_sampling_cnn_b200 import train, workload
    wl = workload.make_2d(3)
    torch.manual_seed(0)
    model = train.UNet_cooler().eval().to(cuda_device)
    x = torch.from_numpy((wl.occ_maps[wl.task_to_map] != 0).astype(np.float32))[:, None].repeat(1, 3, 1, 1).to(cuda_device)
    coords = torch.from_numpy(wl.coords).to(cuda_device)
    y = model(x, coords)
    with torch.no_grad():
        ref = cnn_oracle.forward({k: v.float() for k, v in model.state_dict().items()}, x, coords, 2)
    torch.cuda.synchronize()
    err = (y - ref).abs().max().item()
    assert err <= 1e-2, f"max abs err {err} (ref range {ref.min().item():.3f}..{ref.max().item():.3f})"


def test_shared_encoder_equals_per_task_encoder(cuda_device):
    """GuidedPlanner runs the encoder once per distinct map; logits must equal the per-task forward bit-for-bit."""
    from mgr_sampling_cnn_b200 import engine, train, workload
    wl = workload.make_2d(13, size=128)  # 2 maps: 10 + 3 tasks -> micro-batches with 1 and 2 distinct maps
    torch.manual_seed(0)
    model = train.UNet_cooler().eval().to(cuda_device)
    gp = engine.GuidedPlanner(model, 2, wl.shape, micro_batch=6)
    occ = torch.from_numpy(wl.occ_maps).to(cuda_device)
    t2m = torch.from_numpy(wl.task_to_map).to(cuda_device)
    coords = torch.from_numpy(wl.coords).to(cuda_device)
    a = gp.probability_maps(occ, t2m, coords).clone()
    gp.share_encoder = False
    b = gp.probability_maps(occ, t2m, coords).clone()
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    x = torch.from_numpy((wl.occ_maps[wl.task_to_map] != 0).astype(np.float32))[:, None].repeat(1, 3, 1, 1).to(cuda_device)
    with torch.no_grad():
        ref = cnn_oracle.forward({k: v.float() for k, v in model.state_dict().items()}, x, coords, 2)
    assert (a.reshape(ref.shape) - ref).abs().max().item() <= 1e-2
