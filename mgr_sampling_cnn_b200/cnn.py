"""Host side of the CNN forward: packs the reference models' parameters for the tensor-core kernels and runs the layer
graph of UNet_cooler.forward (train.py:171-203) / ThreeD_UNet_cooler.forward (ThreeD_train.py:169-204) through
libnrrt.so.  PyTorch holds the parameters and the device buffers; every FLOP of the forward is a hand-written kernel
(nrrt_conv_layer = tcgen05 implicit GEMM; nrrt_stem_conv / nrrt_coords_* / nrrt_head_1x1 = small CUDA-core kernels).

Activations are fp16 channels-last.  The three `torch.cat` calls of the decoder never copy: producers write straight
into channel slices of the concat buffers and TMA reads slices back out.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check

K3S1, K3S2, K1S2, K1S1, CONVT, UPFOLD = 0, 1, 2, 3, 4, 5


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _prod(v):
    out = 1
    for a in v:
        out *= int(a)
    return out


class ConvProfiler(object):
    """bench.py's roofline leg: a thin view of the library's own launch profiler (nrrt_conv_profile_*, csrc/conv_tc.cu).
    While `active`, EVERY nrrt_conv_layer launch of this thread -- from the Python layer graph or from inside
    nrrt_conv_forward_* -- is counted with its algorithmic FLOPs and bytes, and every `sample_every`-th launch is bracketed
    by CUDA events on its stream: a uniform sample over the launch sequence (period 29 per encoder pass / 6 per decoder
    micro-batch, coprime with the default 5), so sum(flops) / sum(time) over the sample is the flops-weighted rate of the
    whole step.  Read the results after a synchronize."""

    def __init__(self, sample_every: int = 5):
        self.sample_every = max(int(sample_every), 1)
        self._active = False
        self._started = False
        self.seen, self.flops_seen, self.bytes_seen = 0, 0.0, 0.0
        self.records = []  # (tag, ms, flops, bytes) of the timed sample, filled by collect()

    @property
    def active(self):
        return self._active

    @active.setter
    def active(self, on):
        on = bool(on)
        lib = _lib.load()
        if on and not self._active:  # the first activation clears the library's records, later ones resume counting
            check(lib.nrrt_conv_profile_resume() if self._started else lib.nrrt_conv_profile_start(self.sample_every))
            self._started = True
        elif not on and self._active:
            check(lib.nrrt_conv_profile_stop())
        self._active = on

    def reset(self):
        check(_lib.load().nrrt_conv_profile_stop())
        self._started = self._active = False
        self.seen, self.flops_seen, self.bytes_seen, self.records = 0, 0.0, 0.0, []

    def collect(self):
        """Fetch totals and the timed sample from the library (call after torch.cuda.synchronize())."""
        import numpy as np
        lib = _lib.load()
        seen, fl, by, n = C.c_int64(), C.c_double(), C.c_double(), C.c_int32()
        check(lib.nrrt_conv_profile_read(C.byref(seen), C.byref(fl), C.byref(by), 0, None, None, None, None, C.byref(n)))
        m = n.value
        tags, ms = np.zeros(m, np.int32), np.zeros(m, np.float32)
        flops, nbytes = np.zeros(m, np.float64), np.zeros(m, np.float64)
        if m:
            check(lib.nrrt_conv_profile_read(None, None, None, m, tags.ctypes.data_as(C.c_void_p), ms.ctypes.data_as(C.c_void_p),
                                             flops.ctypes.data_as(C.c_void_p), nbytes.ctypes.data_as(C.c_void_p), C.byref(n)))
        self.seen, self.flops_seen, self.bytes_seen = seen.value, fl.value, by.value
        self.records = list(zip(tags.tolist(), ms.tolist(), flops.tolist(), nbytes.tolist()))
        return self

    def by_layer(self):
        """{layer name: (launches, flops, milliseconds)} of the timed sample (after collect())."""
        out = {}
        for tag, ms, f, _ in self.records:
            name = _lib.tag_name(tag)
            n, fl, t = out.get(name, (0, 0.0, 0.0))
            out[name] = (n + 1, fl + f, t + ms)
        return out

    def totals(self):
        """(launches, total flops, total milliseconds) of the timed sample (after collect())."""
        return len(self.records), sum(r[2] for r in self.records), sum(r[1] for r in self.records)


def _cpu(t, dtype=torch.float32):
    """Parameter -> host tensor.  All packing / folding below is host arithmetic (once per model, a few MB): building a plan
    launches no GPU kernels of its own, only uploads -- the launch list of a process shows this library's kernels."""
    return t.detach().to(device="cpu", dtype=dtype)


def _fold_bn(bn, conv_bias, cout, device):
    """eval-mode BatchNorm (+ optional conv bias) -> per-channel (scale, shift) in fp32."""
    if bn is None:
        scale = torch.ones(cout, dtype=torch.float32)
        shift = torch.zeros(cout, dtype=torch.float32) if conv_bias is None else _cpu(conv_bias)
        return scale.contiguous().to(device), shift.contiguous().to(device)
    g, b = _cpu(bn.weight), _cpu(bn.bias)
    mu, var = _cpu(bn.running_mean), _cpu(bn.running_var)
    scale = g / torch.sqrt(var + bn.eps)
    shift = b - mu * scale
    if conv_bias is not None:
        shift = shift + _cpu(conv_bias) * scale
    return scale.contiguous().to(device), shift.contiguous().to(device)


def _pack_conv_weight(w, cin_pad, device):
    """[cout, cin, (kd,) kh, kw] -> fp16 [cout, taps, cin_pad] (taps in (kd, kh, kw) order, zero-padded channels)."""
    w = _cpu(w)
    cout, cin = w.shape[0], w.shape[1]
    taps = w[0, 0].numel()
    w = w.reshape(cout, cin, taps).permute(0, 2, 1)  # [cout, taps, cin]
    out = torch.zeros((cout, taps, cin_pad), dtype=torch.float16)
    out[:, :, :cin] = w.to(torch.float16)
    return out.contiguous().to(device)


def _pack_convT_weight(w, device):
    """ConvTranspose k=2 s=2 weight [cin, cout, (2,) 2, 2] -> fp16 [groups*cout, cin]; group g = offsets in (d,h,w) order."""
    w = _cpu(w)
    cin, cout = w.shape[0], w.shape[1]
    groups = w[0, 0].numel()
    w = w.reshape(cin, cout, groups).permute(2, 1, 0)  # [g, cout, cin]
    return w.reshape(groups * cout, cin).to(torch.float16).contiguous().to(device)


def _fold_weights(wt, bu, wc, device):
    """ConvTranspose2d(k=2, s=2) (weight wt [cin_low, c_up, 2, 2], bias bu [c_up]) composed with a 3x3 conv over its output
    (weight wc [cout, c_up, 3, 3]) -- train.py:196-201, nothing non-linear in between.  Output pixel (2i+a, 2j+b) of the conv
    reads the upsampled pixels (2i+a+dy, 2j+b+dx), i.e. sub-position ((a+dy)%2, (b+dx)%2) of low-resolution pixel
    (i + (a+dy)//2, j + (b+dx)//2): per output parity (a, b) a 2x2-tap conv over the low-resolution input.
    Returns (fp16 [cout, 16 * cin_low + 128] = [cout][parity a*2+b][tap][cin_low] + identity columns, tap = (row offset - (a-1)) * 2 + (column
    offset - (b-1)), composed in fp64;  fp32 [9 border cases][cout] = the ConvTranspose bias seen through the conv taps
    that fall inside the image, which nrrt_poly_fold adds to the epilogue polynomial)."""
    wt = _cpu(wt, torch.float64)
    wc = _cpu(wc, torch.float64)
    cout, cin = wc.shape[0], wt.shape[0]
    w = torch.zeros((cout, 4, 4, cin), dtype=torch.float64)
    for a in range(2):
        for b in range(2):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    tap = ((a + dy) // 2 - (a - 1)) * 2 + ((b + dx) // 2 - (b - 1))
                    w[:, a * 2 + b, tap] += wc[:, :, dy + 1, dx + 1] @ wt[:, :, (a + dy) % 2, (b + dx) % 2].t()
    valid = {0: [1, 2], 1: [0, 1, 2], 2: [0, 1]}
    bu = _cpu(bu, torch.float64)
    bc = torch.stack([wc[:, :, valid[r]][:, :, :, valid[c]].sum(dim=(2, 3)) @ bu for r in range(3) for c in range(3)])
    # + 128 identity columns per N half of 128 output channels: the residual tile is accumulated by the tensor core
    # (conv_halo2_kernel, p.res_mma)
    eye = (torch.arange(cout)[:, None] % 128 == torch.arange(128)[None, :]).to(torch.float16)
    wf = torch.cat([w.reshape(cout, 16 * cin).to(torch.float16), eye], dim=1)
    return wf.contiguous().to(device), bc.float().contiguous().to(device)


class _Layer(object):
    """One nrrt_conv_layer call with its packed parameters (kept alive here)."""

    def __init__(self, dims, kind, cin, cout, weight, scale, shift, relu, post=None, cin_real=None):
        self.dims, self.kind, self.cin, self.cout = dims, kind, cin, cout
        self.weight, self.scale, self.shift, self.relu = weight, scale, shift, relu
        self.post = post
        self.cin2 = 0  # channels read from the second input (K split)
        self.cin_real = cin_real or cin
        self.tag = 0   # NrrtLayerTag: reported back by the library's launch profiler

    def flops(self, batch, grid):
        """Algorithmic FLOPs (2 x MACs with the true Cin) of this layer on input grid (D, H, W)."""
        sp = [g for i, g in enumerate(grid) if self.dims == 3 or i > 0]
        if self.kind == CONVT:
            return 2.0 * batch * _prod(sp) * (2 ** self.dims) * self.cout * self.cin_real
        if self.kind == UPFOLD:  # per output pixel (4 per input pixel): 2x2 taps over the low-resolution channels
            return 2.0 * batch * _prod(sp) * 4 * 4 * self.cout * self.cin_real
        stride = 2 if self.kind in (K3S2, K1S2) else 1
        taps = (3 ** self.dims) if self.kind in (K3S1, K3S2) else 1
        return 2.0 * batch * _prod([(g - 1) // stride + 1 for g in sp]) * self.cout * taps * self.cin_real

    def algorithmic_bytes(self, n, grid, x, x2, res, head):
        """Bytes one launch has to move if every operand crossed DRAM exactly once: the distinct input samples (per-map
        sources count once per map), the packed weights, the distinct residual rows, the stored output (or fp32 logits)."""
        sp_in = _prod([g for i, g in enumerate(grid) if self.dims == 3 or i > 0])
        up = 2 ** self.dims if self.kind in (CONVT, UPFOLD) else 1
        stride = 2 if self.kind in (K3S2, K1S2) else 1
        sp_out = _prod([(g - 1) // stride + 1 for i, g in enumerate(grid) if self.dims == 3 or i > 0]) * up
        b = 2.0 * min(n, x.shape[0]) * sp_in * self.cin + 2.0 * self.weight.numel()
        if x2 is not None:
            b += 2.0 * min(n, x2.shape[0]) * sp_in * self.cin2
        if res is not None:
            b += 2.0 * min(n, res.shape[0]) * sp_out * self.cout
        b += (4.0 * n * sp_out) if head is not None else (2.0 * n * sp_out * self.cout)
        return b

    force_per_tap = False  # tests: run the generic per-tap kernel where the slab kernel would be chosen
    no_poly_mma = False    # tests: evaluate the whole polynomial term in the epilogue
    no_tma_store = False   # tests: direct global stores instead of the staged TMA-store epilogue
    res_mma = False        # folded layers: residual accumulated by the tensor core instead of staged for the epilogue

    def run(self, lib, stream, x, x_coff, grid, out, out_coff, res=None, res_coff=0, poly=None, head=None, x2=None,
            idx=None, idx2=None, n=None, res_idx=None, res_group=0, poly3=None):
        """x/out/res: channels-last fp16 buffers [B, (D,) H, W, Ctot]; grid = input (D, H, W).
        head = (weight fp32 [cout], bias float, logits fp32 tensor): fuse the 1x1 output head, `out` may be None.
        x2: optional second input buffer (K split: channels [x: self.cin | x2: self.cin2]); idx / idx2: optional int32
        [n] sample indirection of x / x2 (encoder outputs shared per map); n = number of output samples;
        res_idx: optional int32 [n] sample indirection of the residual (per-map partial sums); res_group: scheduling hint,
        consecutive samples that share a residual row (include/nrrt.h); poly3: 3D coordinate field E [n, 54, 4, cout] added in
        the epilogue together with the ReLU (instead of nrrt_poly_relu_3d over the stored output)."""
        n = n if n is not None else (idx.shape[0] if idx is not None else x.shape[0])
        self._run(lib, stream, x, x_coff, grid, out, out_coff, res, res_coff, poly, head, x2, idx, idx2, n, res_idx, res_group, poly3)

    def _run(self, lib, stream, x, x_coff, grid, out, out_coff, res, res_coff, poly, head, x2, idx, idx2, n, res_idx=None,
             res_group=0, poly3=None):
        _lib.count_launch()
        L = _lib.NrrtConvLayer()
        L.dims, L.kind, L.tag = self.dims, self.kind, self.tag
        L.n, L.in_d, L.in_h, L.in_w = n, grid[0], grid[1], grid[2]
        L.cin, L.cout = self.cin, self.cout
        L.in_n = x.shape[0]
        if idx is not None:
            L.in_index = _ptr(idx)
        if x2 is not None:
            L.in2, L.in2_cstride, L.cin2, L.in2_n = _ptr(x2), x2.shape[-1], self.cin2, x2.shape[0]
            if idx2 is not None:
                L.in2_index = _ptr(idx2)
        esz = 2
        L.in_ = C.c_void_p(x.data_ptr() + x_coff * esz)
        L.in_cstride = x.shape[-1]
        L.weight, L.scale, L.shift = _ptr(self.weight), _ptr(self.scale), _ptr(self.shift)
        if out is not None:
            L.out, L.out_cstride, L.out_coffset = _ptr(out), out.shape[-1], out_coff
        else:
            L.out_cstride, L.out_coffset = self.cout, 0
        if head is not None:
            L.head_weight, L.head_bias, L.logits = _ptr(head[0]), float(head[1]), _ptr(head[2])
        L.force_per_tap = 1 if self.force_per_tap else 0
        L.no_poly_mma = 1 if self.no_poly_mma else 0
        L.no_tma_store = 1 if self.no_tma_store else 0
        L.res_mma = 1 if self.res_mma else 0
        if res is not None:
            L.res, L.res_cstride, L.res_coffset = _ptr(res), res.shape[-1], res_coff
            L.res_n = res.shape[0]
            if res_idx is not None:
                L.res_index = _ptr(res_idx)
                L.res_group = int(res_group)
        L.relu = 1 if (self.relu or poly3 is not None) else 0  # the fused 3D field carries the layer's ReLU with it
        if poly3 is not None:
            L.poly3 = _ptr(poly3)
        if self.post is not None:
            L.post_scale, L.post_shift = _ptr(self.post[0]), _ptr(self.post[1])
        if poly is not None:  # [n, 9, 4, cout], or per output parity [n, 4, 9, 4, cout] for a folded layer
            L.poly, L.poly_cases = _ptr(poly), poly.shape[-3]
        check(lib.nrrt_conv_layer(C.byref(L), stream))


class UNetPlan(object):
    """Packed parameters + buffer plan of one UNet_cooler / ThreeD_UNet_cooler instance on one device."""

    def __init__(self, model, dims, device):
        self.dims, self.device = dims, device
        self.lib = _lib.load()
        d = dims
        dev = device

        def conv(m, kind, bn=None, relu=True, cin_pad=None, post=None):
            cin_pad = cin_pad or m.in_channels
            scale, shift = _fold_bn(bn, m.bias, m.out_channels, dev)
            return _Layer(d, kind, cin_pad, m.out_channels, _pack_conv_weight(m.weight, cin_pad, dev), scale, shift, relu, post,
                          cin_real=m.in_channels)

        def convT(m):
            groups = 2 ** d
            scale = torch.ones(groups * m.out_channels, dtype=torch.float32).to(dev)
            shift = _cpu(m.bias).repeat(groups).contiguous().to(dev)
            return _Layer(d, CONVT, m.in_channels, m.out_channels, _pack_convT_weight(m.weight, dev), scale, shift, False)

        def first_conv(seq):  # the Conv inside a torchvision Conv3DSimple is the module itself; 2D convs are plain
            return seq

        # ---- stem ----
        if d == 2:
            stem_seq = model.conv1  # Conv2d(3,32) ReLU Conv2d(32,64) ReLU            (train.py:118-123)
            post = None
        else:
            stem_seq = model.conv1[0]  # r3d stem with stem[0] replaced                  (ThreeD_train.py:113-118)
            bn = model.conv1[1]        # + BatchNorm3d(64) + ReLU left over from torchvision's BasicStem
            post = _fold_bn(bn, None, 64, dev)
        c0, c1 = stem_seq[0], stem_seq[2]
        self.stem_w = _cpu(c0.weight).reshape(c0.out_channels, -1).contiguous().to(dev)
        self.stem_b = _cpu(c0.bias).contiguous().to(dev)
        self.stem_cin, self.stem_cout = c0.in_channels, c0.out_channels
        self.stem2 = conv(c1, K3S1, None, True, cin_pad=64, post=post)

        # ---- encoder: torchvision BasicBlocks ----
        def blocks(layer):
            out = []
            for blk in layer:
                if d == 2:
                    cv1, cv2, b1, b2 = blk.conv1, blk.conv2, blk.bn1, blk.bn2
                else:  # VideoResNet BasicBlock: conv1 = Sequential(Conv3DSimple, BN, ReLU), conv2 = Sequential(Conv3DSimple, BN)
                    cv1, b1, cv2, b2 = blk.conv1[0], blk.conv1[1], blk.conv2[0], blk.conv2[1]
                s2 = cv1.stride[0] == 2
                ds = None
                if blk.downsample is not None:
                    ds = conv(blk.downsample[0], K1S2 if s2 else K1S1, blk.downsample[1], relu=False)
                out.append((conv(cv1, K3S2 if s2 else K3S1, b1, True), conv(cv2, K3S1, b2, True), ds))
            return out

        self.enc = [blocks(model.conv2), blocks(model.conv3), blocks(model.conv4), blocks(model.conv5)]

        # ---- coordinate branch (fp32, tiny) ----
        # [cout][cin][3][3] -> [tap][cin][cout]: the kernel's threads run over cout
        self.cw = [_cpu(m[0].weight).permute(2, 3, 1, 0).reshape(9, m[0].in_channels, m[0].out_channels).contiguous().to(dev)
                   for m in (model.conv_coords_64, model.conv_coords_128, model.conv_coords_256, model.conv_coords_512)]
        self.cb = [_cpu(m[0].bias).contiguous().to(dev) for m in
                   (model.conv_coords_64, model.conv_coords_128, model.conv_coords_256, model.conv_coords_512)]
        self.cw_ptrs = (C.c_void_p * 4)(*[t.data_ptr() for t in self.cw])
        self.cb_ptrs = (C.c_void_p * 4)(*[t.data_ptr() for t in self.cb])

        # ---- decoder ----
        # 2D: the interpolated coordinate channels are bilinear in the pixel position, so the four layers that consume
        # them (upconv6, conv6.0, conv7.0, conv8.0) drop those channels from the GEMM K and add their contribution as a
        # per-task polynomial in the epilogue (nrrt_coords_poly).  3D keeps them as materialised channels (the (2,3,1)
        # coordinate grid is only piecewise bilinear).
        self.poly = d == 2

        def conv_split(m, c_up, c_enc, c_coords_in_gemm, relu=True):
            """First 3x3 conv of a decoder stage over torch.cat((up, enc, coords)) (train.py:194-201).  The layer is linear
            in its input channels up to the ReLU, so it is split into (per-task layer over [up | coords (3D only)] with
            the bias and the ReLU, per-map layer over the encoder channels without bias or ReLU): the per-map partial sum
            is computed once per map and enters the per-task layer as an (indexed) residual before the ReLU."""
            w = _cpu(m.weight)
            up, enc, crd = w[:, :c_up], w[:, c_up:c_up + c_enc], w[:, c_up + c_enc:]
            wp = torch.cat([up] + ([crd] if c_coords_in_gemm else []), dim=1)
            scale, shift = _fold_bn(None, m.bias, m.out_channels, dev)
            task = _Layer(d, K3S1, wp.shape[1], m.out_channels, _pack_conv_weight(wp, wp.shape[1], dev), scale, shift, relu)
            one = torch.ones(m.out_channels, dtype=torch.float32).to(dev)
            zero = torch.zeros(m.out_channels, dtype=torch.float32).to(dev)
            per_map = _Layer(d, K3S1, c_enc, m.out_channels, _pack_conv_weight(enc, c_enc, dev), one, zero, False)
            return task, per_map

        # Tap moments of the coordinate slices.  Inside the image a conv over a (piecewise) bilinear field is the field's
        # value and derivatives times the moments sum_tap w, sum_tap w*dy, ... of the taps that fall inside the image; per
        # axis the admissible taps of a border case are a 0/1 mask, so every moment is a separable contraction of the weights.
        off = torch.tensor([-1.0, 0.0, 1.0])
        border = torch.tensor([[0.0, 1.0, 1.0], [1.0, 1.0, 1.0], [1.0, 1.0, 0.0]])   # first / interior / last row or column
        pw = torch.stack([torch.ones(3), off])                                       # [order 0 / 1][tap]

        def conv_moments(m, keep):
            """2D: [9 border cases][uv = 00,10,01,11][Cc][Cout] fp32 (uv = order in dy, order in dx)."""
            wc = _cpu(m.weight)[:, keep:]                                            # [Cout, Cc, 3, 3]
            my = border[:, None, :] * pw[None, :, :]                                 # [row case][order][dy]
            out = torch.einsum("ocyx,rsy,qtx->rqtsco", wc, my, my)                   # [ry][rx][order x][order y][Cc][Cout]
            out = out.reshape(9, 2, 2, wc.shape[1], wc.shape[0])                     # uv index = order_x * 2 + order_y
            return out.reshape(9, 4, wc.shape[1], wc.shape[0]).contiguous().to(dev)

        def conv_moments3d(m, keep):
            """3D tap moments of the coordinate slice, the two h-pieces stacked along the channel axis, as the GEMM operand
            [2*Cc][54 cases * 4 uv * Cout]; case = (d-border * 6 + h-class) * 3 + w-border, uv = 00, 10 (dz), 01 (dy), 11."""
            wc = _cpu(m.weight)[:, keep:]                                            # [Cout, Cc, 3, 3, 3] (dz, dy, dx)
            cout, cc = wc.shape[0], wc.shape[1]
            pieces = {0: ([1, 2], []), 1: ([0, 1, 2], []), 2: ([0, 1], [2]), 3: ([0], [1, 2]), 4: ([], [0, 1, 2]), 5: ([], [0, 1])}
            hmask = torch.zeros(6, 2, 3)                                             # [h-class][piece][dy]
            for hc, pcs in pieces.items():
                for pce in range(2):
                    hmask[hc, pce, pcs[pce]] = 1.0
            mz = border[:, None, :] * pw[None, :, :]                                 # [d case][order z][dz]
            mh = hmask[:, :, None, :] * pw[None, None, :, :]                         # [h class][piece][order y][dy]
            t = torch.einsum("oczyx,wx->oczyw", wc, border)                          # w-border cases
            t = torch.einsum("oczyw,hpty->ocztwhp", t, mh).contiguous()
            out = torch.einsum("ocztwhp,dsz->dhwtspco", t, mz)                       # [d][h][w][order y][order z][piece][Cc][Cout]
            out = out.reshape(54, 4, 2 * cc, cout)                                   # uv = order_y * 2 + order_z: 00, 10 (dz), 01 (dy), 11
            return out.permute(2, 0, 1, 3).reshape(2 * cc, -1).contiguous().to(dev)

        mu = model.upconv6
        groups = 2 ** d
        scale = torch.ones(groups * mu.out_channels, dtype=torch.float32).to(dev)
        shift = _cpu(mu.bias).repeat(groups).contiguous().to(dev)
        if self.poly:
            # upconv6 reads the encoder output x5 only (once per map); its coordinate rows become polynomial moments
            self.up6 = _Layer(d, CONVT, 512, mu.out_channels, _pack_convT_weight(_cpu(mu.weight)[:512], dev), scale, shift, False)
            self.m_up6 = _cpu(mu.weight)[512:].permute(2, 3, 0, 1).reshape(4, 1, 512, mu.out_channels).contiguous().to(dev)
            self.m_c6, self.m_c7, self.m_c8 = conv_moments(model.conv6[0], 768), conv_moments(model.conv7[0], 256), conv_moments(model.conv8[0], 128)
            # the same moments as GEMM operands [Cc][cases * n_uv * Cout] (nrrt_coords_poly_gemm)
            self.mt = {k: m.cpu().permute(2, 0, 1, 3).reshape(m.shape[2], -1).contiguous().to(dev)
                       for k, m in (("e_up6", self.m_up6), ("e_c6", self.m_c6), ("e_c7", self.m_c7), ("e_c8", self.m_c8))}
            # conv6.0 is linear in (upconv6 output | x4 | coords): the whole GEMM runs once per map over (t6m | e4), the
            # task-dependent part is the 16-case polynomial of nrrt_poly16 applied by nrrt_poly_relu (csrc/conv_aux.cu)
            m6 = model.conv6[0]
            w6 = _cpu(m6.weight)[:, :768]
            one = torch.ones(m6.out_channels, dtype=torch.float32).to(dev)
            self.c6full = _Layer(d, K3S1, 512, m6.out_channels, _pack_conv_weight(w6, 768, dev), one,
                                 torch.zeros(m6.out_channels, dtype=torch.float32).to(dev), False,
                                 cin_real=768)
            self.c6full.cin2 = 256
            self.w6taps = _cpu(m6.weight)[:, :512].permute(1, 2, 3, 0).reshape(512, 9 * m6.out_channels).contiguous().to(dev)
            self.b6 = _cpu(m6.bias).contiguous().to(dev)
        else:
            # source 1 = x5 (per map), source 2 = the interpolated coordinate channels (per task)
            self.up6 = _Layer(d, CONVT, 512, mu.out_channels, _pack_convT_weight(mu.weight, dev), scale, shift, False)
            self.up6.cin2 = 512
            self.up6.cin_real = 1024
        self.up7, self.up8 = convT(model.upconv7), convT(model.upconv8)

        def fold(mu, mc, c_up):
            """upconv7 -> conv7.0 / upconv8 -> conv8.0 as ONE layer (_fold_weights); the conv's bias and ReLU stay."""
            w, bc = _fold_weights(mu.weight, mu.bias, _cpu(mc.weight)[:, :c_up], dev)
            # no BatchNorm here: the affine is the identity and the conv's bias joins the ConvTranspose bias in the constant
            # term of the epilogue polynomial (nrrt_poly_fold adds `bc` per border case), so the epilogue has no scale/shift
            bc = (bc.cpu() + _cpu(mc.bias)[None, :]).contiguous().to(dev)
            layer = _Layer(d, UPFOLD, mu.in_channels, mc.out_channels, w, None, None, True)
            layer.res_mma = UNetPlan.res_mma  # class switch for experiments (profiles/conv_fold_r1.md), never the environment
            return layer, bc
        # 3D: the coordinate channels leave the GEMM too; their (piecewise bilinear) field and the ReLU are applied by
        # nrrt_poly_relu_3d after the layer (csrc/conv_aux.cu)
        c60, self.c6e = (None, None) if self.poly else conv_split(model.conv6[0], 512, 256, False, relu=False)
        c70, self.c7e = conv_split(model.conv7[0], 128, 128, False, relu=self.poly)
        c80, self.c8e = conv_split(model.conv8[0], 64, 64, False, relu=self.poly)
        if not self.poly:
            self.mt3 = {"e_c6": conv_moments3d(model.conv6[0], 768), "e_c7": conv_moments3d(model.conv7[0], 256),
                        "e_c8": conv_moments3d(model.conv8[0], 128)}
        # 2D: upconv7 -> conv7.0 and upconv8 -> conv8.0 as one folded layer each (used when the low-resolution grid holds
        # whole 16x16 tiles; the separate layers above remain for small grids and as the test reference)
        self.c7f, self.bc7 = fold(model.upconv7, model.conv7[0], 128) if self.poly else (None, None)
        self.c8f, self.bc8 = fold(model.upconv8, model.conv8[0], 64) if self.poly else (None, None)
        self.c6 = (c60, conv(model.conv6[2], K3S1))
        self.c7 = (c70, conv(model.conv7[2], K3S1))
        self.c8 = (c80, conv(model.conv8[2], K3S1))
        self.head_w = _cpu(model.final_conv.weight).reshape(-1).contiguous().to(dev)
        self.head_b = float(_cpu(model.final_conv.bias).item())
        self._bufs = {}
        self.profiler = None
        self._assign_tags()

    # activation buffers: an encoder set for Be samples (distinct maps) and a decoder set for B samples (tasks)
    def _cached(self, key, make):
        bufs = self._bufs.get(key)
        if bufs is None:
            if len(self._bufs) >= 4:  # full + ragged configuration of each set
                self._bufs.pop(next(iter(self._bufs)))
            bufs = make()
            self._bufs[key] = bufs
        return bufs

    def _buf(self, n, grid, level, ch):
        D, H, W = grid
        s = 2 ** level
        return torch.empty((n, D // s if self.dims == 3 else 1, H // s, W // s, ch), dtype=torch.float16, device=self.device)

    def _enc_buffers(self, Be, grid):
        b = lambda lv, ch: self._buf(Be, grid, lv, ch)  # noqa: E731
        return self._cached(("enc", Be) + tuple(grid), lambda: {
            "t0": b(0, 64), "x1": b(0, 64), "a0": b(0, 64), "b0": b(0, 64), "e2": b(0, 64),
            "a1": b(1, 128), "b1": b(1, 128), "d1": b(1, 128), "e3": b(1, 128),
            "a2": b(2, 256), "b2": b(2, 256), "d2": b(2, 256), "e4": b(2, 256),
            "a3": b(3, 512), "b3": b(3, 512), "d3": b(3, 512), "e5": b(3, 512),
            # per-map partial sums of conv6.0 / conv7.0 / conv8.0 over the encoder channels (2D: all of conv6.0's GEMM,
            # over t6m = upconv6(x5) and x4)
            "r6": b(2, 512), "r7": b(1, 256), "r8": b(0, 128), "t6m": b(2, 512) if self.poly else None})

    def _dec_buffers(self, B, grid):
        b = lambda lv, ch: self._buf(B, grid, lv, ch)  # noqa: E731
        # task-side decoder inputs: the ConvTranspose output, followed (3D) by the interpolated coordinate channels
        t8w, t7w, t6w = 64, 128, 512  # (the coordinate channels are never materialised next to them)

        def make():
            d = {"j": b(0, 128), "g7": b(1, 256), "i": b(1, 128), "g6": b(2, 512), "h": b(2, 256)}
            if not self.use_fold(grid):  # the ConvTranspose outputs exist only where the layers run unfolded
                d["t8"], d["t7"] = b(0, t8w), b(1, t7w)
            if not self.poly:
                d["t6"] = b(2, t6w)
                d["c5"] = b(3, 512)  # interpolated coordinate channels at the bottleneck (second input of upconv6)
            return d
        return self._cached(("dec", B, self.use_fold(grid)) + tuple(grid), make)

    def activation_bytes(self, B, grid):
        b = list(self._enc_buffers(B, grid).values()) + list(self._dec_buffers(B, grid).values())
        return sum(t.numel() * t.element_size() for t in b if t is not None)

    def layers(self):
        out = [self.stem2, self.up6, self.up7, self.up8] + list(self.c6) + list(self.c7) + list(self.c8)
        out += [self.c6e, self.c7e, self.c8e] + ([self.c6full, self.c7f, self.c8f] if self.poly else [])
        out = [l for l in out if l is not None]
        for stg in self.enc:
            for c1, c2, ds in stg:
                out += [c1, c2] + ([ds] if ds is not None else [])
        return out

    def set_profiler(self, prof):
        """bench.py: the ConvProfiler whose `active` flag GuidedPlanner.plan toggles around the CNN stage."""
        self.profiler = prof

    def _assign_tags(self):
        T = _lib
        for name, tag in (("stem2", T.TAG_STEM2), ("up6", T.TAG_UP6), ("up7", T.TAG_UP7), ("up8", T.TAG_UP8),
                          ("c6full", T.TAG_C6FULL), ("c6e", T.TAG_C6E), ("c7e", T.TAG_C7E), ("c8e", T.TAG_C8E),
                          ("c7f", T.TAG_C7F), ("c8f", T.TAG_C8F)):
            if getattr(self, name, None) is not None:
                getattr(self, name).tag = tag
        for name, t0 in (("c6", T.TAG_C60), ("c7", T.TAG_C70), ("c8", T.TAG_C80)):
            for i, l in enumerate(getattr(self, name)):
                if l is not None:
                    l.tag = t0 + i
        for si, stg in enumerate(self.enc):
            for bi, blk in enumerate(stg):
                for li, l in enumerate(blk):
                    if l is not None:
                        l.tag = T.TAG_ENC + (si * 2 + bi) * 3 + li

    def c_weights(self):
        """NrrtCnnWeights (include/nrrt.h) over this plan's packed device tensors: the argument of nrrt_conv_forward_*."""
        cw = getattr(self, "_c_weights", None)
        if cw is not None:
            return cw
        W = _lib.NrrtCnnWeights()

        def fill(dst, l):
            if l is None:
                return
            dst.weight, dst.scale, dst.shift = _ptr(l.weight), _ptr(l.scale), _ptr(l.shift)
            if l.post is not None:
                dst.post_scale, dst.post_shift = _ptr(l.post[0]), _ptr(l.post[1])
            dst.kind, dst.cin, dst.cin2, dst.cout, dst.relu, dst.tag = l.kind, l.cin, l.cin2, l.cout, 1 if l.relu else 0, l.tag

        W.dims, W.stem_cin, W.stem_cout = self.dims, self.stem_cin, self.stem_cout
        W.stem_w, W.stem_b = _ptr(self.stem_w), _ptr(self.stem_b)
        fill(W.stem2, self.stem2)
        for si, stg in enumerate(self.enc):
            for bi, blk in enumerate(stg):
                for li, l in enumerate(blk):
                    fill(W.enc[si][bi][li], l)
        for k in range(4):
            W.coords_w[k], W.coords_b[k] = self.cw[k].data_ptr(), self.cb[k].data_ptr()
        fill(W.up6, self.up6), fill(W.up7, self.up7), fill(W.up8, self.up8)
        fill(W.c6e, self.c6e), fill(W.c7e, self.c7e), fill(W.c8e, self.c8e)
        fill(W.c60, self.c6[0]), fill(W.c61, self.c6[1]), fill(W.c70, self.c7[0]), fill(W.c71, self.c7[1])
        fill(W.c80, self.c8[0]), fill(W.c81, self.c8[1])
        if self.poly:
            fill(W.c6full, self.c6full), fill(W.c7f, self.c7f), fill(W.c8f, self.c8f)
            W.bc7, W.bc8 = _ptr(self.bc7), _ptr(self.bc8)
            W.mt_up6, W.mt_c6, W.mt_c7, W.mt_c8 = (_ptr(self.mt[k]) for k in ("e_up6", "e_c6", "e_c7", "e_c8"))
            W.w6taps, W.b6 = _ptr(self.w6taps), _ptr(self.b6)
        else:
            W.mt3_c6, W.mt3_c7, W.mt3_c8 = (_ptr(self.mt3[k]) for k in ("e_c6", "e_c7", "e_c8"))
        W.head_w, W.head_bias = _ptr(self.head_w), self.head_b
        self._c_weights = W
        return W

    def forward_tasks(self, occ_maps, map_ids, task_to_row, coords, logits, tasks_per_map=0):
        """The whole forward in ONE library call (nrrt_conv_forward_2d / _3d, csrc/forward.cu): logits[t] for task t on map
        occ_maps[map_ids[task_to_row[t]]].  occ_maps uint8 [n_maps, ...] and coords fp32 on the device; map_ids and
        task_to_row are HOST int32 arrays (task_to_row non-decreasing); logits fp32 [n_tasks, cells] on the device."""
        import numpy as np
        lib = self.lib
        map_ids = np.ascontiguousarray(map_ids, dtype=np.int32)
        task_to_row = np.ascontiguousarray(task_to_row, dtype=np.int32)
        n_tasks, n_maps = int(task_to_row.shape[0]), int(map_ids.shape[0])
        shp = tuple(int(v) for v in occ_maps.shape[1:])
        # 3D: the network's (d, h, w) are the map file's (z, x, y) axes
        d, h, w = (1, shp[0], shp[1]) if self.dims == 2 else (shp[2], shp[0], shp[1])
        W = self.c_weights()
        need = lib.nrrt_conv_forward_workspace_bytes(C.byref(W), d, h, w, n_tasks)
        ws = getattr(self, "_ws", None)
        if ws is None or ws.numel() < need or ws.device != self.device:
            self._ws = ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        self._host_keep = (map_ids, task_to_row)  # copied on the stream: keep the host arrays alive until the next call
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        coords = coords.contiguous().float()
        _lib.count_launch(self.launches_per_forward(task_to_row))
        if self.dims == 2:
            check(lib.nrrt_conv_forward_2d(C.byref(W), _ptr(occ_maps), map_ids.ctypes.data_as(C.c_void_p), n_maps,
                                           task_to_row.ctypes.data_as(C.c_void_p), _ptr(coords), n_tasks, h, w, int(tasks_per_map),
                                           _ptr(logits), _ptr(ws), ws.numel(), st))
        else:
            check(lib.nrrt_conv_forward_3d(C.byref(W), _ptr(occ_maps), map_ids.ctypes.data_as(C.c_void_p), n_maps,
                                           task_to_row.ctypes.data_as(C.c_void_p), _ptr(coords), n_tasks, d, h, w, int(tasks_per_map),
                                           _ptr(logits), _ptr(ws), ws.numel(), st))
        return logits

    def launches_per_forward(self, task_to_row):
        """Kernels one nrrt_conv_forward_* call enqueues (bench.py's gpu_launches claim), by replaying its schedule
        (csrc/forward.cu: 8 maps per encoder pass, 40 / 10 tasks per decoder micro-batch, 4096 / 128 tasks per
        coordinate-branch chunk)."""
        mb, pc = (40, 4096) if self.dims == 2 else (10, 128)
        enc = 25 if self.dims == 2 else 24          # stem, stem2, 19 ResNet convs, per-map partial sums
        dec = 6 if self.dims == 2 else 13           # 3D: + coordinate fill, three ConvTranspose, post-passes
        prep = 17 if self.dims == 2 else 10
        n = len(task_to_row)
        total, lo, prep_hi = 0, 0, 0
        while lo < n:
            g = int(task_to_row[lo]) // 8
            hi = lo
            while hi < n and int(task_to_row[hi]) // 8 == g:
                hi += 1
            total += enc
            for i in range(lo, hi, mb):
                j = min(hi, i + mb)
                if j > prep_hi:
                    prep_hi = min(n, i + pc)
                    total += prep
                total += dec
            lo = hi
        return total

    fold = True  # tests: False runs ConvTranspose and the conv after it as separate layers
    res_mma = False  # experiments: folded layers accumulate the residual through the tensor core
    fuse_poly3 = True  # tests: False applies conv8.0's 3D coordinate field with the nrrt_poly_relu_3d post-pass

    def use_fold(self, grid):
        """upconv7/conv7.0 and upconv8/conv8.0 as folded layers: 2D, low-resolution grids of at least one 16x16 tile."""
        return self.poly and self.fold and min(grid[1], grid[2]) >= 64

    def flops_per_sample(self, grid):
        """Algorithmic FLOPs of the tensor-core layers for one sample on input grid (D, H, W) (excludes the 3-/1-channel
        stem conv, the coordinate branch and the 1x1 head, which run on CUDA cores)."""
        g = [tuple(max(v // (2 ** lv), 1) if (self.dims == 3 or i > 0) else 1 for i, v in enumerate(grid)) for lv in range(4)]
        tot = self.stem2.flops(1, g[0])
        gin = [g[0], g[0], g[1], g[2]]
        for s_i, stg in enumerate(self.enc):
            gout = g[s_i]
            (c1, c2, ds), (c3, c4, _) = stg
            tot += c1.flops(1, gin[s_i]) + c2.flops(1, gout) + c3.flops(1, gout) + c4.flops(1, gout)
            if ds is not None:
                tot += ds.flops(1, gin[s_i])
        c6_first = self.c6full if self.poly else self.c6[0]
        tot += self.up6.flops(1, g[3]) + c6_first.flops(1, g[2]) + self.c6[1].flops(1, g[2])
        if self.use_fold(grid):
            tot += self.c7f.flops(1, g[2]) + self.c7[1].flops(1, g[1]) + self.c8f.flops(1, g[1]) + self.c8[1].flops(1, g[0])
        else:
            tot += self.up7.flops(1, g[2]) + self.c7[0].flops(1, g[1]) + self.c7[1].flops(1, g[1])
            tot += self.up8.flops(1, g[1]) + self.c8[0].flops(1, g[0]) + self.c8[1].flops(1, g[0])
        tot += (0 if self.poly else self.c6e.flops(1, g[2])) + self.c7e.flops(1, g[1]) + self.c8e.flops(1, g[0])
        return tot

    def forward(self, x, coords, logits_out=None):
        """x: fp32 (B,3,H,W) / (B,D,H,W) on self.device; coords: fp32 (B,1,2,2) / (B,1,2,3).  Returns fp32 logits
        (B,1,H,W) / (B,1,D,H,W)."""
        x = x.contiguous().float()
        if self.dims == 2:
            grid = (1, x.shape[2], x.shape[3])
            if x.shape[1] != 3:
                raise ValueError("UNet_cooler expects (B, 3, H, W)")
        else:
            grid = (x.shape[1], x.shape[2], x.shape[3])
        return self._forward(x.shape[0], grid, coords, logits_out, x=x)

    def forward_maps(self, occ_maps, task_to_map, coords, logits_out=None, enc_maps=None, enc_index=None, prep=None):
        """Same forward fed from resident uint8 occupancy maps [n_maps, (D,) H, W] + task_to_map [B] (int32): the
        network input of sample b is the {0,1}-normalised map task_to_map[b] replicated over the input channels.

        Encoder reuse: the encoder (conv1..conv5) depends on the map only, so callers that know several tasks share a
        map pass `enc_maps` [Be] (distinct map ids, int32) and `enc_index` [B] (task -> row of enc_maps): the encoder
        then runs on Be samples and its four outputs are gathered into every task's concat buffers."""
        grid = (1,) + tuple(occ_maps.shape[1:]) if self.dims == 2 else tuple(occ_maps.shape[1:])
        B = coords.shape[0]
        if enc_maps is None:
            return self._forward(B, grid, coords, logits_out, occ=occ_maps, t2m=task_to_map, prep=prep)
        return self._forward(B, grid, coords, logits_out, occ=occ_maps, t2m=enc_maps, enc_index=enc_index, prep=prep)

    def coords_prepare(self, coords, grid):
        """Coordinate branch for a whole batch of tasks in a few launches: the four conv_coords_* feature vectors
        (fp32 [B, P, C]) and, in 2D, the epilogue polynomials E of the four consumer layers (fp32 [B, cases, 4, Cout])."""
        coords = coords.contiguous().float()
        B, P = coords.shape[0], coords.shape[2] * coords.shape[3]
        dev = self.device
        feats = [torch.empty((B, P, c), dtype=torch.float32, device=dev) for c in (64, 128, 256, 512)]
        ptrs = (C.c_void_p * 4)(*[t.data_ptr() for t in feats])
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.count_launch()
        check(self.lib.nrrt_coords_branch(_ptr(coords), B, coords.shape[2], coords.shape[3], self.cw_ptrs, self.cb_ptrs, ptrs, st))
        prep = {"feats": feats}
        if self.poly:
            if P != 4:
                raise ValueError("2D coords must be (B, 1, 2, 2)")
            g = [tuple(max(v // (2 ** lv), 1) if i > 0 else 1 for i, v in enumerate(grid)) for lv in range(4)]
            for name, feat, mom, lv in (("e_up6", feats[3], self.m_up6, 3), ("e_c6", feats[2], self.m_c6, 2),
                                        ("e_c7", feats[1], self.m_c7, 1), ("e_c8", feats[0], self.m_c8, 0)):
                cases, n_uv, cc, cout = mom.shape
                E = torch.empty((B, cases, 4, cout), dtype=torch.float32, device=dev)
                a = 1.0 / (g[lv][1] - 1) if g[lv][1] > 1 else 0.0
                b = 1.0 / (g[lv][2] - 1) if g[lv][2] > 1 else 0.0
                # one fp32 GEMM over the batch (chunked so that the scratch stays small); the same kernels at every batch
                # size, so that results do not depend on how a caller batches (nrrt_coords_poly is the direct form of it)
                chunk = 512
                work = torch.empty((min(B, chunk) * 4 * (cc + cases * n_uv * cout),), dtype=torch.float32, device=dev)
                for i in range(0, B, chunk):
                    j = min(B, i + chunk)
                    _lib.count_launch(3)
                    check(self.lib.nrrt_coords_poly_gemm(_ptr(feat[i:j]), j - i, cc, _ptr(self.mt[name]), cases, n_uv, cout, a, b,
                                                         _ptr(work), _ptr(E[i:j]), st))
                prep[name] = E
            # conv6.0's task-dependent term: 16-case polynomial from upconv6's and conv6.0's own coordinate channels
            cout = self.b6.shape[0]
            P16 = torch.empty((B, 16, 4, cout), dtype=torch.float32, device=dev)
            chunk = 512
            work = torch.empty((min(B, chunk) * 16 * 9 * cout,), dtype=torch.float32, device=dev)
            for i in range(0, B, chunk):
                j = min(B, i + chunk)
                _lib.count_launch(2)
                check(self.lib.nrrt_poly16(_ptr(prep["e_up6"][i:j]), _ptr(self.w6taps), _ptr(prep["e_c6"][i:j]), _ptr(self.b6),
                                           j - i, 512, cout, g[3][1], g[3][2], _ptr(work), _ptr(P16[i:j]), st))
            prep["p16"] = P16
            del prep["e_up6"], prep["e_c6"]
            if self.use_fold(grid):  # per-parity polynomials of the folded layers (tile space = the low-resolution grid)
                for name, bc, lv in (("e_c7", self.bc7, 2), ("e_c8", self.bc8, 1)):
                    E = prep[name]
                    Ef = torch.empty((B, 4) + tuple(E.shape[1:]), dtype=torch.float32, device=dev)
                    _lib.count_launch()
                    check(self.lib.nrrt_poly_fold(_ptr(E), _ptr(bc), B, E.shape[-1], g[lv][1], g[lv][2], _ptr(Ef), st))
                    prep[name] = Ef
        else:
            # 3D: post-pass polynomials of conv6.0 / conv7.0 / conv8.0 (54 cases, two h-pieces stacked along the channels)
            if P != 6:
                raise ValueError("3D coords must be (B, 1, 2, 3)")
            g = self._levels(grid)
            for name, feat, lv, cout in (("e_c6", feats[2], 2, 512), ("e_c7", feats[1], 1, 256), ("e_c8", feats[0], 0, 128)):
                cc = feat.shape[-1]
                A = torch.empty((B, 4, 2 * cc), dtype=torch.float32, device=dev)
                _lib.count_launch()
                check(self.lib.nrrt_coords_pieces_3d(_ptr(feat), B, cc, _ptr(A), st))
                E = torch.empty((B, 54, 4, cout), dtype=torch.float32, device=dev)
                n_cols = 54 * 4 * cout
                chunk = max(1, min(B, (64 << 20) // (4 * (2 * cc + n_cols))))  # scratch <= 256 MB
                work = torch.empty((chunk * 4 * (2 * cc + n_cols),), dtype=torch.float32, device=dev)
                a = 1.0 / (g[lv][0] - 1)
                b = 1.0 / (g[lv][1] - 1)
                for i in range(0, B, chunk):
                    j = min(B, i + chunk)
                    _lib.count_launch(2)
                    check(self.lib.nrrt_coords_poly_gemm(_ptr(A[i:j]), j - i, 2 * cc, _ptr(self.mt3[name]), 54, 4, cout, a, b,
                                                         _ptr(work), _ptr(E[i:j]), st))
                prep[name] = E
        return prep

    @staticmethod
    def slice_prep(prep, i, j):
        return {k: ([f[i:j] for f in v] if isinstance(v, list) else v[i:j]) for k, v in prep.items()}

    def _levels(self, grid):
        return [tuple(max(v // (2 ** lv), 1) if (self.dims == 3 or i > 0) else 1 for i, v in enumerate(grid)) for lv in range(4)]

    def encode(self, grid, Be, x=None, occ=None, maps=None):
        """Encoder (conv1 .. conv5, train.py:173-177) for Be samples: either network inputs x or rows `maps` of the
        resident uint8 occupancy maps.  Returns the buffer set holding e2..e5 (x2..x5 of the reference)."""
        lib, dims = self.lib, self.dims
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        if any(v % 8 for v in grid[(0 if dims == 3 else 1):]):
            raise ValueError("spatial sizes must be multiples of 8 (three stride-2 stages)")
        bf = self._enc_buffers(Be, grid)
        g = self._levels(grid)
        _lib.count_launch()
        if x is not None:
            check(lib.nrrt_stem_conv(dims, _ptr(x), Be, self.stem_cin, grid[0], grid[1], grid[2], _ptr(self.stem_w),
                                     _ptr(self.stem_b), self.stem_cout, 64, _ptr(bf["t0"]), st))
        else:
            check(lib.nrrt_stem_conv_maps(dims, _ptr(occ), _ptr(maps), Be, self.stem_cin, grid[0], grid[1], grid[2],
                                          _ptr(self.stem_w), _ptr(self.stem_b), self.stem_cout, 64, _ptr(bf["t0"]), st))
        self.stem2.run(lib, st, bf["t0"], 0, g[0], bf["x1"], 0)

        def stage(blks, xin, gin, gout, a, b, dsb, final):
            (c1, c2, ds), (c3, c4, _) = blks
            res = xin
            if ds is not None:
                ds.run(lib, st, xin, 0, gin, dsb, 0)
                res = dsb
            c1.run(lib, st, xin, 0, gin, a, 0)
            c2.run(lib, st, a, 0, gout, b, 0, res, 0)
            c3.run(lib, st, b, 0, gout, a, 0)
            c4.run(lib, st, a, 0, gout, final, 0, b, 0)

        stage(self.enc[0], bf["x1"], g[0], g[0], bf["a0"], bf["b0"], None, bf["e2"])
        stage(self.enc[1], bf["e2"], g[0], g[1], bf["a1"], bf["b1"], bf["d1"], bf["e3"])
        stage(self.enc[2], bf["e3"], g[1], g[2], bf["a2"], bf["b2"], bf["d2"], bf["e4"])
        stage(self.enc[3], bf["e4"], g[2], g[3], bf["a3"], bf["b3"], bf["d3"], bf["e5"])
        # encoder-side partial sums of the decoder's first convs (linear in torch.cat((up, enc)), train.py:194-201)
        if self.poly:
            self.up6.run(lib, st, bf["e5"], 0, g[3], bf["t6m"], 0)
            self.c6full.run(lib, st, bf["t6m"], 0, g[2], bf["r6"], 0, x2=bf["e4"])
        else:
            self.c6e.run(lib, st, bf["e4"], 0, g[2], bf["r6"], 0)
        self.c7e.run(lib, st, bf["e3"], 0, g[1], bf["r7"], 0)
        self.c8e.run(lib, st, bf["e2"], 0, g[0], bf["r8"], 0)
        return bf

    def _forward(self, B, grid, coords, logits_out, x=None, occ=None, t2m=None, enc_index=None, prep=None):
        Be = t2m.shape[0] if enc_index is not None else B
        enc = self.encode(grid, Be, x=x, occ=occ, maps=t2m)
        return self.decode(enc, B, grid, coords, prep, enc_index, logits_out)

    def decode(self, enc, B, grid, coords, prep=None, enc_index=None, logits_out=None, res_group=0):
        """Coordinate branch + decoder (train.py:179-203) for B tasks whose encoder outputs are rows enc_index[b]
        (None: row b) of the buffer set returned by encode().  res_group: consecutive tasks per map (scheduling hint)."""
        lib, dims = self.lib, self.dims
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        coords = coords.contiguous().float()
        bf = dict(self._dec_buffers(B, grid))
        bf.update({k: enc[k] for k in ("e5", "r6", "r7", "r8")})
        g = self._levels(grid)

        # coordinate branch (train.py:179-190)
        if prep is None:
            prep = self.coords_prepare(coords, grid)
        if not self.poly:  # 3D: only the bottleneck coordinate channels (second input of upconv6, 10^3 voxels) are materialised
            gh, gw = coords.shape[2], coords.shape[3]
            t = bf["c5"]
            _lib.count_launch()
            check(lib.nrrt_coords_fill(_ptr(prep["feats"][3]), B, gh, gw, 512, dims, g[3][0], g[3][1], g[3][2], _ptr(t),
                                       t.shape[-1], 0, st))

        def post3d(buf, E, lv):  # 3D: coordinate field + ReLU after the first conv of a decoder stage
            _lib.count_launch()
            check(lib.nrrt_poly_relu_3d(_ptr(buf), _ptr(E), B, g[lv][0], g[lv][1], g[lv][2], buf.shape[-1], st))

        # decoder: every torch.cat of the reference is a K split over (task buffer, encoder output of the task's map)
        ei = enc_index  # None = one encoder sample per task
        ri = ei  # None = one encoder sample per task (the residual of sample b is row b)
        rg = res_group
        if self.poly:
            _lib.count_launch()
            check(lib.nrrt_poly_relu(_ptr(bf["r6"]), _ptr(ri), _ptr(prep["p16"]), _ptr(bf["g6"]), B, g[2][1], g[2][2],
                                     bf["g6"].shape[-1], st))
        else:
            self.up6.run(lib, st, bf["e5"], 0, g[3], bf["t6"], 0, x2=bf["c5"], idx=ei, n=B)
            self.c6[0].run(lib, st, bf["t6"], 0, g[2], bf["g6"], 0, bf["r6"], 0, res_idx=ri, res_group=rg)
            post3d(bf["g6"], prep["e_c6"], 2)
        self.c6[1].run(lib, st, bf["g6"], 0, g[2], bf["h"], 0)
        folded = self.poly and prep["e_c7"].dim() == 5  # coords_prepare decided (use_fold)
        if folded:
            self.c7f.run(lib, st, bf["h"], 0, g[2], bf["g7"], 0, bf["r7"], 0, poly=prep["e_c7"], res_idx=ri, res_group=rg)
        else:
            self.up7.run(lib, st, bf["h"], 0, g[2], bf["t7"], 0)
            self.c7[0].run(lib, st, bf["t7"], 0, g[1], bf["g7"], 0, bf["r7"], 0, poly=prep.get("e_c7") if self.poly else None, res_idx=ri, res_group=rg)
        if not self.poly:
            post3d(bf["g7"], prep["e_c7"], 1)
        self.c7[1].run(lib, st, bf["g7"], 0, g[1], bf["i"], 0)
        fuse3 = False
        if folded:
            self.c8f.run(lib, st, bf["i"], 0, g[1], bf["j"], 0, bf["r8"], 0, poly=prep["e_c8"], res_idx=ri, res_group=rg)
        else:
            self.up8.run(lib, st, bf["i"], 0, g[1], bf["t8"], 0)
            # 3D: conv8.0 adds its coordinate field and the ReLU in the epilogue when the slices hold whole 16x16 tiles
            fuse3 = (not self.poly) and self.fuse_poly3 and g[0][1] % 16 == 0 and g[0][2] % 16 == 0
            self.c8[0].run(lib, st, bf["t8"], 0, g[0], bf["j"], 0, bf["r8"], 0, poly=prep.get("e_c8") if self.poly else None, res_idx=ri,
                           res_group=rg, poly3=prep["e_c8"] if fuse3 else None)
        if not self.poly and not fuse3:
            post3d(bf["j"], prep["e_c8"], 0)

        # conv8.2 + final_conv fused: the 64-channel activation is reduced to the logit in the epilogue
        logits = logits_out
        if logits is None:
            logits = torch.empty((B, 1) + (tuple(grid) if dims == 3 else tuple(grid[1:])), dtype=torch.float32, device=self.device)
        self.c8[1].run(lib, st, bf["j"], 0, g[0], None, 0, head=(self.head_w, self.head_b, logits))
        return logits



class CudaForwardMixin(object):
    """forward() for the drop-in model classes: eval-mode inference through UNetPlan, micro-batched."""
    _nrrt_dims = 2
    micro_batch = 8

    def _plan(self, device):
        plan = getattr(self, "_nrrt_plan", None)
        if plan is None or plan.device != device:
            plan = UNetPlan(self, self._nrrt_dims, device)
            object.__setattr__(self, "_nrrt_plan", plan)
        return plan

    def invalidate_plan(self):
        """Call after changing parameters in place (load_state_dict does it automatically)."""
        object.__setattr__(self, "_nrrt_plan", None)

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self.invalidate_plan()
        return out

    def forward(self, x, coords):
        if self.training:
            raise _lib.NrrtError("mgr_sampling_cnn_b200 implements inference only: call model.eval() first "
                                 "(training is out of scope, DESIGN.md)")
        if not x.is_cuda:
            raise _lib.NrrtError("forward() needs CUDA tensors: there is no CPU fallback (the oracle lives in oracle/)")
        plan = self._plan(x.device)
        with torch.no_grad():
            outs = [plan.forward(x[i:i + self.micro_batch], coords[i:i + self.micro_batch])
                    for i in range(0, x.shape[0], self.micro_batch)]
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)
