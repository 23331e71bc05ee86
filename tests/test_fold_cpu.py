"""CPU: host-side algebra of the folded ConvTranspose + conv layer (cnn._fold_weights) and of the per-parity polynomial
(the transform nrrt_poly_fold applies on the device, restated in numpy) against torch's own layers.  No GPU needed."""
import numpy as np
import torch
import torch.nn.functional as F

from mgr_sampling_cnn_b200 import cnn


def _folded_forward(x, wf, bc, cin, cout):
    """Evaluate the folded layer the way the kernel does: per output parity a 2x2-tap conv over the low-resolution input,
    then the ConvTranspose bias through the conv taps inside the image (9 border cases)."""
    B, _, H, W = x.shape
    w = wf[:, :16 * cin].double().reshape(cout, 4, 4, cin)
    out = torch.zeros(B, cout, 2 * H, 2 * W, dtype=torch.float64)
    xp = F.pad(x.double(), (1, 1, 1, 1))
    for a in range(2):
        for b in range(2):
            acc = torch.zeros(B, cout, H, W, dtype=torch.float64)
            for dyi in range(2):
                for dxi in range(2):
                    ty, tx = a - 1 + dyi, b - 1 + dxi
                    acc += torch.einsum("oc,nchw->nohw", w[:, a * 2 + b, dyi * 2 + dxi], xp[:, :, 1 + ty:1 + ty + H, 1 + tx:1 + tx + W])
            out[:, :, a::2, b::2] = acc
    Ho, Wo = 2 * H, 2 * W
    rc = torch.tensor([0 if h == 0 else (2 if h == Ho - 1 else 1) for h in range(Ho)])
    cc = torch.tensor([0 if v == 0 else (2 if v == Wo - 1 else 1) for v in range(Wo)])
    return out + bc.double()[rc[:, None] * 3 + cc[None, :]].permute(2, 0, 1)[None]


def test_fold_weights_equal_convtranspose_then_conv():
    g = torch.Generator().manual_seed(0)
    cin, cup, cout, H, W = 16, 8, 128, 5, 7
    wt = torch.randn(cin, cup, 2, 2, generator=g)
    bu = torch.randn(cup, generator=g)
    wc = torch.randn(cout, cup, 3, 3, generator=g) / 8
    x = torch.randn(2, cin, H, W, generator=g)
    wf, bc = cnn._fold_weights(wt, bu, wc, "cpu")
    assert wf.shape == (cout, 16 * cin + 128) and wf.dtype == torch.float16 and bc.shape == (9, cout)
    # identity columns: row r has its single 1 in column r % 128 (the residual path through the tensor core)
    eye = wf[:, 16 * cin:].float()
    assert torch.equal(eye, (torch.arange(cout)[:, None] % 128 == torch.arange(128)[None, :]).float())
    ref = F.conv2d(F.conv_transpose2d(x.double(), wt.double(), bu.double(), stride=2), wc.double(), padding=1)
    got = _folded_forward(x, wf, bc, cin, cout)
    # the composite weights are rounded once to fp16: 2^-11 relative per weight
    assert (got - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()


def test_per_parity_polynomial_equals_the_output_space_polynomial():
    """nrrt_poly_fold (conv_aux.cu), restated: pixel (i, j) of parity (a, b) on the (H, W) grid = output pixel (2i+a, 2j+b)."""
    rng = np.random.default_rng(1)
    H, W, C = 6, 5, 3
    E = rng.normal(size=(9, 4, C))
    Ho, Wo = 2 * H, 2 * W
    al_h, al_w = 2.0 * (H - 1) / (Ho - 1), 2.0 * (W - 1) / (Wo - 1)
    for a in range(2):
        for b in range(2):
            bh, bw = a / (Ho - 1), b / (Wo - 1)
            for i in range(H):
                for j in range(W):
                    rt = 0 if i == 0 else (2 if i == H - 1 else 1)
                    ct = 0 if j == 0 else (2 if j == W - 1 else 1)
                    ro = 0 if (rt == 0 and a == 0) else (2 if (rt == 2 and a == 1) else 1)
                    co = 0 if (ct == 0 and b == 0) else (2 if (ct == 2 and b == 1) else 1)
                    e0, e1, e2, e3 = E[ro * 3 + co]
                    f = (e0 + bh * e1 + bw * e2 + bh * bw * e3, al_h * (e1 + bw * e3), al_w * (e2 + bh * e3), al_h * al_w * e3)
                    ty_t, tx_t = i / (H - 1), j / (W - 1)
                    folded = f[0] + ty_t * f[1] + tx_t * f[2] + ty_t * tx_t * f[3]
                    h, w = 2 * i + a, 2 * j + b
                    rc = 0 if h == 0 else (2 if h == Ho - 1 else 1)
                    cc = 0 if w == 0 else (2 if w == Wo - 1 else 1)
                    assert (rc, cc) == (ro, co)
                    g0, g1, g2, g3 = E[rc * 3 + cc]
                    ty, tx = h / (Ho - 1), w / (Wo - 1)
                    assert np.allclose(folded, g0 + ty * g1 + tx * g2 + ty * tx * g3, rtol=1e-12, atol=1e-12)
