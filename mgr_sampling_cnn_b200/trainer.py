"""Training step of UNet_cooler on the B200 kernels: the host graph behind the reference's
`UNet_cooler.training_step` / `configure_optimizers` (train.py:205-234) -- forward in train mode (batch-statistics
BatchNorm2d in the ResNet blocks), nn.BCEWithLogitsLoss, backward, torch.optim.Adam(lr=1e-3), ReduceLROnPlateau(0.5, 2).

Every FLOP runs in libnrrt.so: the forward convolutions and all data gradients are nrrt_conv_layer (tcgen05 implicit GEMM;
a data gradient is the same convolution with the weights of nrrt_pack_dgrad_weight -- stride-2 layers and the
ConvTranspose through zero-stuffed / space-to-depth views of the gradient), the weight gradients are nrrt_conv_wgrad
(tcgen05, K = pixels), BatchNorm / ReLU / loss / Adam are the one-pass kernels of csrc/train_ops.cu.  PyTorch holds the
buffers.  Parameters live in ONE flat fp32 buffer in the kernels' own layouts ([cout][tap][cin] conv weights), with an
fp16 mirror for the tensor cores, a flat gradient buffer and flat Adam moments: the optimizer is one launch.

2D only (UNet_cooler); image sides must be multiples of 64 (512 and up keep every level's rows whole 64-pixel blocks for the
weight-gradient kernel; smaller images run with partly idle blocks at the coarse levels).  There is no CPU fallback.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import check
from .cnn import CONVT, K1S1, K1S2, K3S1, K3S2, _Layer

BN_EPS, BN_MOMENTUM = 1e-5, 0.1


def _pp(t, coff=0):
    return None if t is None else C.c_void_p(t.data_ptr() + coff * t.element_size())


def to_kernel_layout(t, what: str, convT: bool = False, cin_pad: int = None):
    """torch parameter -> the layout it has in the trainer's flat buffers (pure tensor reshuffling, any device):
    'w'  conv weight [cout][cin][k][k] -> [cout][taps][cin_pad] (zero-padded channels); ConvTranspose2d weight [cin][cout][2][2] ->
         [(a, b)][cout][cin], the forward kernel's group-major rows;
    'b'  bias; a ConvTranspose bias is stored once per output offset (4 copies, all receiving the same gradient);
    'cw' conv_coords weight [cout][cin][3][3] -> [tap][cin][cout];  'v' anything else, unchanged."""
    if what == "w":
        if convT:
            cin, cout = t.shape[0], t.shape[1]
            return t.permute(2, 3, 1, 0).reshape(4 * cout, cin)
        cout, cin = t.shape[0], t.shape[1]
        w = t.reshape(cout, cin, -1).permute(0, 2, 1)
        if cin_pad is not None and cin_pad != cin:
            w = torch.nn.functional.pad(w, (0, cin_pad - cin))
        return w
    if what == "b":
        return t.repeat(4) if convT else t
    if what == "cw":
        return t.permute(2, 3, 1, 0)
    return t


def from_kernel_layout(v, what: str, shape, convT: bool = False, cin_pad: int = None):
    """Inverse of to_kernel_layout: flat buffer segment `v` -> a tensor of the torch parameter's `shape`."""
    shape = tuple(int(x) for x in shape)
    if what == "w":
        if convT:
            cin, cout = shape[0], shape[1]
            return v.reshape(2, 2, cout, cin).permute(3, 2, 0, 1).reshape(shape)
        cout, cin = shape[0], shape[1]
        taps = int(np.prod(shape[2:]))
        return v.reshape(cout, taps, cin_pad or cin)[:, :, :cin].permute(0, 2, 1).reshape(shape)
    if what == "b" and convT:
        return v[:shape[0]].reshape(shape)
    if what == "cw":
        return v.reshape(3, 3, shape[1], shape[0]).permute(3, 2, 0, 1).reshape(shape)
    return v.reshape(shape)


class PlateauScheduler(object):
    """torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, factor=0.5, patience=2) as configure_optimizers() sets it up
    (train.py:230-234; mode 'min', relative threshold 1e-4, no cooldown, min_lr 0): host logic, monitored on val_loss."""

    def __init__(self, lr: float, factor: float = 0.5, patience: int = 2, threshold: float = 1e-4, min_lr: float = 0.0, eps: float = 1e-8):
        self.lr, self.factor, self.patience, self.threshold, self.min_lr, self.eps = float(lr), factor, patience, threshold, min_lr, eps
        self.best, self.num_bad_epochs = math.inf, 0

    def step(self, metric: float) -> float:
        if metric < self.best * (1.0 - self.threshold):
            self.best, self.num_bad_epochs = float(metric), 0
        else:
            self.num_bad_epochs += 1
        if self.num_bad_epochs > self.patience:
            new_lr = max(self.lr * self.factor, self.min_lr)
            if self.lr - new_lr > self.eps:
                self.lr = new_lr
            self.num_bad_epochs = 0
        return self.lr

    def state_dict(self):
        return {"lr": self.lr, "best": self.best, "num_bad_epochs": self.num_bad_epochs}

    def load_state_dict(self, d):
        self.lr, self.best, self.num_bad_epochs = float(d["lr"]), float(d["best"]), int(d["num_bad_epochs"])


class _Conv(object):
    """One convolution of the model: where its parameters live in the flat buffers and the layers that run it."""

    def __init__(self, name, kind, cin, cin_pad, cout, taps):
        self.name, self.kind, self.cin, self.cin_pad, self.cout, self.taps = name, kind, cin, cin_pad, cout, taps
        self.rows = 4 * cout if kind == CONVT else cout
        self.w_off = self.b_off = None
        self.bn = None          # (gamma_off, beta_off, running_mean, running_var, name)
        self.fwd = self.dgrad = self.wd = None


class UNetTrainer(object):
    def __init__(self, model, batch: int, height: int, width: int, device="cuda:0", lr: float = 1e-3, betas=(0.9, 0.999),
                 eps: float = 1e-8, loss_scale: float = None, data_parallel: bool = None):
        if not torch.cuda.is_available():
            raise _lib.NrrtError("UNetTrainer needs a CUDA device (sm_100a); there is no CPU fallback")
        if height % 64 or width % 64:
            raise ValueError("image sides must be multiples of 64 (three stride-2 stages, 8-pixel tiles at the coarsest)")
        self.lib = _lib.load()
        self.dev = torch.device(device)
        self.B, self.H, self.W = int(batch), int(height), int(width)
        self.lr, self.betas, self.eps = float(lr), betas, float(eps)
        self.step_count = 0
        self._exported_steps = 0
        self.M = self.B * self.H * self.W
        # Activation gradients are fp16, so the loss is scaled: dlogits = (sigmoid - t) * g.  g lives on the device
        # (self.scale_state[0]) and adapts like torch.cuda.amp.GradScaler without a host round trip: a step whose gradients are
        # not finite is skipped and halves g, 200 good steps double it.  g = 128 keeps the deepest encoder gradients out of the
        # fp16 subnormals at initialisation.  A fixed `loss_scale` (in units of the mean loss: g = loss_scale / M) turns it off.
        self.dynamic_scale = loss_scale is None
        self.loss_scale = float(128.0 * self.M if loss_scale is None else loss_scale)
        self.scale_state = torch.tensor([self.loss_scale / self.M, 0.0, 0.0, 0.0], dtype=torch.float32, device=self.dev)
        # data parallel: one process per GPU (torch.distributed, NCCL), each rank steps on its own `batch` samples; the flat
        # gradient buffer is summed across ranks with ONE all-reduce before Adam (the loss is the mean over the global batch,
        # BatchNorm statistics stay per rank like torch DDP's default)
        import torch.distributed as dist
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.dp = (self.world > 1) if data_parallel is None else bool(data_parallel and self.world > 1)
        self.model = model
        self.training = True        # False: BatchNorm layers use their running statistics (validation_step)
        self.bn_momentum = BN_MOMENTUM
        self._tmp = {}
        self.scheduler = PlateauScheduler(self.lr)
        self._build_params(model)
        self._alloc()

    # ------------------------------------------------------------------------------------------------------------ params
    def _build_params(self, model):
        dev = self.dev
        self.convs, self.segs = {}, []      # segs: (offset, count, kind, torch tensor getter) for import / export
        total = [0]

        def seg(n):
            off = total[0]
            total[0] += (n + 63) // 64 * 64
            return off

        def add_conv(name, mod, kind, bn=None, cin_pad=None):
            cin = mod.in_channels
            cout = mod.out_channels
            taps = 9 if kind in (K3S1, K3S2) else 1
            cv = _Conv(name, kind, cin, cin_pad or cin, cout, taps)
            cv.w_off = seg(cv.rows * taps * cv.cin_pad)
            self.segs.append((cv.w_off, cv, "w", mod.weight))
            if mod.bias is not None:
                cv.b_off = seg(cv.rows)
                self.segs.append((cv.b_off, cv, "b", mod.bias))
            if bn is not None:
                g, b = seg(cout), seg(cout)
                self.segs.append((g, cout, "v", bn.weight))
                self.segs.append((b, cout, "v", bn.bias))
                cv.bn = [g, b, bn.running_mean.detach().to(dev, torch.float32).clone(),
                         bn.running_var.detach().to(dev, torch.float32).clone(), bn]
            self.convs[name] = cv
            return cv

        m = model
        st0 = m.conv1[0]
        self.stem_w = seg(st0.out_channels * st0.in_channels * 9)
        self.stem_b = seg(st0.out_channels)
        self.segs.append((self.stem_w, st0.weight.numel(), "v", st0.weight))
        self.segs.append((self.stem_b, st0.out_channels, "v", st0.bias))
        self.stem_cin, self.stem_cout = st0.in_channels, st0.out_channels
        add_conv("conv1.2", m.conv1[2], K3S1, cin_pad=64)
        self.stages = []
        for si, layer in enumerate((m.conv2, m.conv3, m.conv4, m.conv5)):
            blks = []
            for bi, blk in enumerate(layer):
                s2 = blk.conv1.stride[0] == 2
                pre = "enc%d.%d." % (si, bi)
                c1 = add_conv(pre + "c1", blk.conv1, K3S2 if s2 else K3S1, blk.bn1)
                c2 = add_conv(pre + "c2", blk.conv2, K3S1, blk.bn2)
                ds = add_conv(pre + "ds", blk.downsample[0], K1S2 if s2 else K1S1, blk.downsample[1]) if blk.downsample is not None else None
                blks.append((c1, c2, ds))
            self.stages.append(blks)
        self.coords = []                    # (w_off [9][cin][cout], b_off, cin, cout)
        for mod in (m.conv_coords_64[0], m.conv_coords_128[0], m.conv_coords_256[0], m.conv_coords_512[0]):
            w, b = seg(mod.weight.numel()), seg(mod.out_channels)
            self.segs.append((w, None, "cw", mod.weight))
            self.segs.append((b, mod.out_channels, "v", mod.bias))
            self.coords.append((w, b, mod.in_channels, mod.out_channels))
        for name, mod, kind in (("up6", m.upconv6, CONVT), ("c6.0", m.conv6[0], K3S1), ("c6.2", m.conv6[2], K3S1),
                                ("up7", m.upconv7, CONVT), ("c7.0", m.conv7[0], K3S1), ("c7.2", m.conv7[2], K3S1),
                                ("up8", m.upconv8, CONVT), ("c8.0", m.conv8[0], K3S1), ("c8.2", m.conv8[2], K3S1)):
            add_conv(name, mod, kind)
        self.head_w, self.head_b = seg(64), seg(1)
        self.segs.append((self.head_w, 64, "v", m.final_conv.weight))
        self.segs.append((self.head_b, 1, "v", m.final_conv.bias))
        self.n_params = total[0]
        self.params = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
        self.params_h = torch.zeros(self.n_params, dtype=torch.float16, device=dev)
        self.grads = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
        self.ones = torch.ones(2048, dtype=torch.float32, device=dev)
        self.import_from(model)
        # layers (views into the flat buffers stay valid for the trainer's lifetime)
        for cv in self.convs.values():
            n = cv.rows * cv.taps * cv.cin_pad
            wv = self.params_h[cv.w_off:cv.w_off + n].view(cv.rows, cv.taps, cv.cin_pad)
            if cv.bn is not None:
                scale = shift = None
            else:
                scale, shift = self.ones[:cv.rows], self.params[cv.b_off:cv.b_off + cv.rows]
            cv.fwd = _Layer(2, cv.kind, cv.cin_pad, cv.cout, wv, scale, shift, cv.bn is None and cv.kind != CONVT, cin_real=cv.cin)
        # data-gradient weights: one flat fp16 buffer, repacked by ONE launch per step.  A 192-channel input (conv8[0]) gets 256
        # gradient rows (64 of zeros): its data gradient then runs on the 256-wide pair kernel instead of three N = 64 tiles
        descs, off = [], 0
        for cv in self.convs.values():
            cv.dx_ch = 256 if cv.cin_pad == 192 else cv.cin_pad
            descs.append([cv.w_off, off, cv.rows, cv.taps, cv.cin_pad, cv.dx_ch, cv.rows, 0])
            cv.wd_off = off
            off += (cv.dx_ch * cv.taps * cv.rows + 63) // 64 * 64
        self.wd_all = torch.zeros(off, dtype=torch.float16, device=dev)
        self.wd_descs = torch.tensor(descs, dtype=torch.int64, device=dev)
        for cv in self.convs.values():
            cv.wd = self.wd_all[cv.wd_off:cv.wd_off + cv.dx_ch * cv.taps * cv.rows].view(cv.dx_ch, cv.taps, cv.rows)
            cv.dgrad = _Layer(2, K3S1 if cv.taps == 9 else K1S1, cv.rows, cv.dx_ch, cv.wd, None, None, False)
        self._refresh_weights()

    def _master_view(self, t, cv, what):
        """torch parameter -> the flat buffer's layout (host-side plumbing at import / export only)."""
        t = t.detach().to(self.dev, torch.float32)
        is_conv = isinstance(cv, _Conv)
        return to_kernel_layout(t, what, convT=is_conv and cv.kind == CONVT, cin_pad=cv.cin_pad if is_conv else None)

    def import_from(self, model):
        with torch.no_grad():
            for off, cv, what, p in self.segs:
                v = self._master_view(p, cv, what).reshape(-1)
                self.params[off:off + v.numel()].copy_(v)

    def _torch_layout(self, flat):
        """[(torch parameter, tensor in its layout)] for a flat buffer laid out like self.params (parameters or gradients)."""
        out = []
        for off, cv, what, p in self.segs:
            is_conv = isinstance(cv, _Conv)
            convT = is_conv and cv.kind == CONVT
            n = self._master_view(p, cv, what).numel()
            out.append((p, from_kernel_layout(flat[off:off + n], what, p.shape, convT=convT, cin_pad=cv.cin_pad if is_conv else None)))
        return out

    def named(self, flat, unscale: float = 1.0):
        """{state_dict key: tensor} of a flat buffer (e.g. self.grads with unscale = 1 / loss_scale): for tests and inspection."""
        names = {}
        for k, p in self.model.named_parameters(remove_duplicate=False):
            if not k.startswith("base_model."):
                names[id(p)] = k
        return {names[id(p)]: (v * unscale).contiguous() for p, v in self._torch_layout(flat)}

    def export_to(self, model=None):
        """Write the trained parameters (and BatchNorm running statistics) back into the torch module's state."""
        model = model or self.model
        with torch.no_grad():
            for p, v in self._torch_layout(self.params):
                p.copy_(v.to(p.device, p.dtype))
            for cv in self.convs.values():
                if cv.bn is not None:
                    cv.bn[4].running_mean.copy_(cv.bn[2].to(cv.bn[4].running_mean.device))
                    cv.bn[4].running_var.copy_(cv.bn[3].to(cv.bn[4].running_var.device))
                    cv.bn[4].num_batches_tracked += self.step_count - self._exported_steps
        self._exported_steps = self.step_count
        if hasattr(model, "invalidate_plan"):
            model.invalidate_plan()
        return model

    def _st(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def _refresh_weights(self):
        """fp16 mirror of the whole parameter buffer + the data-gradient weight layouts of every conv: two launches."""
        lib, st = self.lib, self._st()
        _lib.count_launch(2)
        check(lib.nrrt_to_half(_pp(self.params), _pp(self.params_h), self.n_params, st))
        check(lib.nrrt_pack_dgrad_weights(_pp(self.params), _pp(self.wd_all), _pp(self.wd_descs), len(self.convs), st))

    # ----------------------------------------------------------------------------------------------------------- buffers
    def _alloc(self):
        B, dev = self.B, self.dev
        self.lv = [(self.H >> k, self.W >> k) for k in range(4)]

        def act(level, ch):
            return torch.empty((B,) + self.lv[level] + (ch,), dtype=torch.float16, device=dev)

        self.act = act
        self.t0, self.x1 = act(0, 64), act(0, 64)
        self.cat8, self.cat7, self.cat6, self.c5 = act(0, 192), act(1, 384), act(2, 1024), act(3, 1024)
        self.a8, self.b8 = act(0, 128), act(0, 64)
        self.a7, self.b7 = act(1, 256), act(1, 128)
        self.a6, self.b6 = act(2, 512), act(2, 256)
        self.logits = torch.empty((B, self.H, self.W), dtype=torch.float32, device=dev)
        self.dlogits = torch.empty_like(self.logits)
        self.target = torch.empty_like(self.logits)
        self.loss_sum = torch.zeros(1, dtype=torch.float64, device=dev)
        self.feats = [torch.empty((B, 4, c), dtype=torch.float32, device=dev) for c in (64, 128, 256, 512)]
        self.dfeats = [torch.empty((B, 4, c), dtype=torch.float32, device=dev) for c in (64, 128, 256, 512)]
        self.bn_work = torch.empty(2 * 512, dtype=torch.float64, device=dev)
        self.bn_ss = torch.empty(2 * 512, dtype=torch.float32, device=dev)
        # per block: raw conv outputs, activations, saved statistics
        self.blk = []
        stage_out = [(self.cat8, 64), (self.cat7, 128), (self.cat6, 512), (self.c5, 0)]
        for si, blks in enumerate(self.stages):
            ch = 64 << si
            row = []
            for bi, (c1, c2, ds) in enumerate(blks):
                d = {"z1": act(si, ch), "h": act(si, ch), "z2": act(si, ch),
                     "st1": torch.empty((2, ch), dtype=torch.float32, device=dev), "st2": torch.empty((2, ch), dtype=torch.float32, device=dev)}
                if ds is not None:
                    d["zd"], d["idn"] = act(si, ch), act(si, ch)
                    d["std"] = torch.empty((2, ch), dtype=torch.float32, device=dev)
                d["out"] = stage_out[si] if bi == len(blks) - 1 else (act(si, ch), 0)
                row.append(d)
            self.blk.append(row)

    def _buf(self, key, shape, dtype=torch.float16):
        t = self._tmp.get(key)
        n = int(np.prod(shape))
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty(n, dtype=dtype, device=self.dev)
            self._tmp[key] = t
        return t[:n].view(shape)

    # ----------------------------------------------------------------------------------------------------------- forward
    def _bn_fwd(self, cv, z, y, y_coff, stats, res=None, res_coff=0, relu=True):
        c = cv.cout
        pixels = z.shape[0] * z.shape[1] * z.shape[2]
        if not self.training:   # eval mode: running statistics, nothing updated, nothing saved for a backward
            _lib.count_launch(2)
            check(self.lib.nrrt_bn_eval_forward(_pp(z), c, pixels, c, _pp(self.params, cv.bn[0]), _pp(self.params, cv.bn[1]), BN_EPS,
                                                _pp(cv.bn[2]), _pp(cv.bn[3]), _pp(res, res_coff), res.shape[-1] if res is not None else 0,
                                                1 if relu else 0, _pp(y, y_coff), y.shape[-1], _pp(self.bn_ss), self._st()))
            return
        _lib.count_launch(3)
        check(self.lib.nrrt_bn_train_forward(_pp(z), c, pixels, c, _pp(self.params, cv.bn[0]), _pp(self.params, cv.bn[1]), BN_EPS, self.bn_momentum,
                                             _pp(cv.bn[2]), _pp(cv.bn[3]), _pp(res, res_coff), res.shape[-1] if res is not None else 0,
                                             1 if relu else 0, _pp(y, y_coff), y.shape[-1], _pp(stats[0]), _pp(stats[1]), _pp(self.bn_ss),
                                             _pp(self.bn_work), self._st()))

    def forward(self, x, coords):
        """x fp32 [B, 3, H, W] (the dataset's image tensor), coords fp32 [B, 1, 2, 2] -> self.logits WITHOUT final_conv's bias
        (it is added where the logits are consumed: nrrt_bce_with_logits / predict)."""
        lib, st, B = self.lib, self._st(), self.B
        g = [(1,) + s for s in self.lv]
        self.x_in = x
        _lib.count_launch()
        check(lib.nrrt_stem_conv(2, _pp(x), B, self.stem_cin, 1, self.H, self.W, _pp(self.params, self.stem_w), _pp(self.params, self.stem_b),
                                 self.stem_cout, 64, _pp(self.t0), st))
        self.convs["conv1.2"].fwd.run(lib, st, self.t0, 0, g[0], self.x1, 0)
        xin, xin_off, lv_in = self.x1, 0, 0
        for si, blks in enumerate(self.stages):
            for bi, (c1, c2, ds) in enumerate(blks):
                d = self.blk[si][bi]
                d["xin"] = (xin, xin_off, lv_in)
                c1.fwd.run(lib, st, xin, xin_off, g[lv_in], d["z1"], 0)
                self._bn_fwd(c1, d["z1"], d["h"], 0, d["st1"])
                c2.fwd.run(lib, st, d["h"], 0, g[si], d["z2"], 0)
                if ds is not None:
                    ds.fwd.run(lib, st, xin, xin_off, g[lv_in], d["zd"], 0)
                    self._bn_fwd(ds, d["zd"], d["idn"], 0, d["std"], relu=False)
                    res, res_off = d["idn"], 0
                else:
                    res, res_off = xin, xin_off
                out, out_off = d["out"]
                self._bn_fwd(c2, d["z2"], out, out_off, d["st2"], res=res, res_coff=res_off)
                xin, xin_off, lv_in = out, out_off, si
        # coordinate branch (fp32) and its four bilinear fills into the concat buffers (train.py:179-190)
        ptrs = (C.c_void_p * 4)(*[t.data_ptr() for t in self.feats])
        cw = (C.c_void_p * 4)(*[self.params.data_ptr() + 4 * c[0] for c in self.coords])
        cb = (C.c_void_p * 4)(*[self.params.data_ptr() + 4 * c[1] for c in self.coords])
        self.coords_in = coords.contiguous().float()
        _lib.count_launch(5)
        check(lib.nrrt_coords_branch(_pp(self.coords_in), B, 2, 2, cw, cb, ptrs, st))
        for feat, ch, buf, off, lv in ((self.feats[0], 64, self.cat8, 128, 0), (self.feats[1], 128, self.cat7, 256, 1),
                                       (self.feats[2], 256, self.cat6, 768, 2), (self.feats[3], 512, self.c5, 512, 3)):
            check(lib.nrrt_coords_fill(_pp(feat), B, 2, 2, ch, 2, 1, self.lv[lv][0], self.lv[lv][1], _pp(buf), buf.shape[-1], off, st))
        cv = self.convs
        cv["up6"].fwd.run(lib, st, self.c5, 0, g[3], self.cat6, 0)
        cv["c6.0"].fwd.run(lib, st, self.cat6, 0, g[2], self.a6, 0)
        cv["c6.2"].fwd.run(lib, st, self.a6, 0, g[2], self.b6, 0)
        cv["up7"].fwd.run(lib, st, self.b6, 0, g[2], self.cat7, 0)
        cv["c7.0"].fwd.run(lib, st, self.cat7, 0, g[1], self.a7, 0)
        cv["c7.2"].fwd.run(lib, st, self.a7, 0, g[1], self.b7, 0)
        cv["up8"].fwd.run(lib, st, self.b7, 0, g[1], self.cat8, 0)
        cv["c8.0"].fwd.run(lib, st, self.cat8, 0, g[0], self.a8, 0)
        cv["c8.2"].fwd.run(lib, st, self.a8, 0, g[0], self.b8, 0)
        _lib.count_launch()
        check(lib.nrrt_head_1x1(_pp(self.b8), self.M, 64, 64, _pp(self.params, self.head_w), 0.0, _pp(self.logits), st))
        return self.logits

    # ---------------------------------------------------------------------------------------------------------- backward
    def _cm(self, key, src, coff, c, mode, copies=1, mask=None):
        """channels-last slice -> channel-major scratch (nrrt_to_channel_major)."""
        n, h, w = src.shape[0], src.shape[1], src.shape[2]
        shape = {0: (copies, c, n, h, w), 1: (c, n, 2 * h, 2 * w), 2: (4, c, n, h // 2, w // 2)}[mode]
        out = self._buf(key, shape)
        _lib.count_launch()
        check(self.lib.nrrt_to_channel_major(_pp(src, coff), src.shape[-1], _pp(mask), mask.shape[-1] if mask is not None else 0, n, h, w, c,
                                             mode, copies, _pp(out), self._st()))
        return out

    def _wgrad(self, cv, g, g_coff, x, x_coff):
        """grads[cv.w] += weight gradient from channels-last buffers: g = masked output gradient (channels [g_coff, +rows)), x = the
        layer's input (channels [x_coff, +cin_pad)), both on the grid the stride-1 form of the layer runs on (nrrt_conv_wgrad_nhwc)."""
        k = 3 if cv.taps == 9 else 1
        _lib.count_launch()
        check(self.lib.nrrt_conv_wgrad_nhwc(_pp(g, g_coff), g.shape[-1], cv.rows, _pp(x, x_coff), x.shape[-1], cv.cin_pad, self.B,
                                            x.shape[1], x.shape[2], k, _pp(self.grads, cv.w_off), cv.cin_pad, self._st()))

    def _relu_bwd(self, dy, dy_off, c, y, y_off, out, dbias_off):
        pixels = dy.shape[0] * dy.shape[1] * dy.shape[2]
        _lib.count_launch()
        check(self.lib.nrrt_relu_bwd_sum(_pp(dy, dy_off), dy.shape[-1], _pp(y, y_off), y.shape[-1] if y is not None else 0,
                                         _pp(out), out.shape[-1] if out is not None else 0,
                                         _pp(self.grads, dbias_off) if dbias_off is not None else None, pixels, c, self._st()))

    def _conv_relu_bwd(self, cv, dy, y, x, lv, need_dx=True):
        """conv + bias + ReLU layer: dy (dense, gradient of the stored output y) is masked in place; returns dx (dense)."""
        lib, st = self.lib, self._st()
        self._relu_bwd(dy, 0, cv.cout, y, 0, dy, cv.b_off)
        self._wgrad(cv, dy, 0, x, 0)
        if not need_dx:
            return None
        dx = self._buf("dx_" + cv.name, (self.B,) + self.lv[lv] + (cv.dx_ch,))     # channels beyond cin_pad (if any) are zeros
        cv.dgrad.run(lib, st, dy, 0, (1,) + self.lv[lv], dx, 0)
        return dx

    def _convT_bwd(self, cv, dcat, x, lv_in):
        """ConvTranspose k2 s2 (+ bias, no activation): dY = channels [0, cout) of dcat on the fine grid; x on grid lv_in."""
        lib, st = self.lib, self._st()
        for gidx in range(4):   # the bias is stored once per output offset (the forward's layout); all four get the same gradient
            self._relu_bwd(dcat, 0, cv.cout, None, 0, None, cv.b_off + gidx * cv.cout)
        s2d = self._buf("s2d", (self.B,) + self.lv[lv_in] + (4 * cv.cout,))
        _lib.count_launch()
        check(lib.nrrt_remap_channels_last(_pp(dcat), dcat.shape[-1], _pp(s2d), 4 * cv.cout, self.B, self.lv[lv_in][0], self.lv[lv_in][1],
                                           cv.cout, 0, st))
        self._wgrad(cv, s2d, 0, x, 0)      # a 1x1 weight gradient over the 4 * cout space-to-depth channels
        dx = self._buf("dx_" + cv.name, (self.B,) + self.lv[lv_in] + (cv.cin_pad,))
        cv.dgrad.run(lib, st, s2d, 0, (1,) + self.lv[lv_in], dx, 0)
        return dx

    def _bn_bwd(self, cv, dy, dy_off, y_mask, y_off, z, stats, dz, g_out=None):
        c = cv.cout
        pixels = z.shape[0] * z.shape[1] * z.shape[2]
        _lib.count_launch(3)
        check(self.lib.nrrt_bn_train_backward(_pp(dy, dy_off), dy.shape[-1], _pp(y_mask, y_off), y_mask.shape[-1] if y_mask is not None else 0,
                                              _pp(z), c, pixels, c, _pp(self.params, cv.bn[0]), _pp(stats[0]), _pp(stats[1]), _pp(dz), c,
                                              _pp(g_out), c if g_out is not None else 0, _pp(self.grads, cv.bn[0]), _pp(self.grads, cv.bn[1]),
                                              _pp(self.bn_work), self._st()))

    def _block_bwd(self, si, bi, dy, dy_off, skip=None):
        """BasicBlock backward.  dy: gradient of the block output (buffer, channel offset).  skip: (buffer, offset) of another
        gradient of the block INPUT to add (the decoder's skip connection).  Returns the dense gradient of the block input."""
        lib, st = self.lib, self._st()
        c1, c2, ds = self.stages[si][bi]
        d = self.blk[si][bi]
        ch = c2.cout
        xin, xin_off, lv_in = d["xin"]
        out, out_off = d["out"]
        shape = (self.B,) + self.lv[si] + (ch,)
        dz2, gres = self._buf("dz2", shape), self._buf("gres", shape)
        self._bn_bwd(c2, dy, dy_off, out, out_off, d["z2"], d["st2"], dz2, g_out=gres)
        self._wgrad(c2, dz2, 0, d["h"], 0)
        dh = self._buf("dh", shape)
        c2.dgrad.run(lib, st, dz2, 0, (1,) + self.lv[si], dh, 0)
        dz1 = self._buf("dz1", shape)
        self._bn_bwd(c1, dh, 0, d["h"], 0, d["z1"], d["st1"], dz1)
        gin = (1,) + self.lv[lv_in]
        dx = self._buf("dx_blk%d%d" % (si, bi), (self.B,) + self.lv[lv_in] + (c1.cin_pad,))
        if ds is None:
            self._wgrad(c1, dz1, 0, xin, xin_off)
            c1.dgrad.run(lib, st, dz1, 0, gin, dx, 0, res=gres, res_coff=0)
            return dx
        # stride-2 block: both convolutions of the input are seen on the input grid through the zero-stuffed gradient
        stuffed = self._buf("stuffed", (self.B,) + self.lv[lv_in] + (ch,))
        _lib.count_launch(2)
        check(lib.nrrt_remap_channels_last(_pp(dz1), ch, _pp(stuffed), ch, self.B, self.lv[si][0], self.lv[si][1], ch, 1, st))
        self._wgrad(c1, stuffed, 0, xin, xin_off)
        dxa = self._buf("dxa", (self.B,) + self.lv[lv_in] + (c1.cin_pad,))
        if skip is not None:
            c1.dgrad.run(lib, st, stuffed, 0, gin, dxa, 0, res=skip[0], res_coff=skip[1])
        else:
            c1.dgrad.run(lib, st, stuffed, 0, gin, dxa, 0)
        dzd = self._buf("dzd", shape)
        self._bn_bwd(ds, gres, 0, None, 0, d["zd"], d["std"], dzd)
        check(lib.nrrt_remap_channels_last(_pp(dzd), ch, _pp(stuffed), ch, self.B, self.lv[si][0], self.lv[si][1], ch, 1, st))
        self._wgrad(ds, stuffed, 0, xin, xin_off)
        ds.dgrad.run(lib, st, stuffed, 0, gin, dx, 0, res=dxa, res_coff=0)
        return dx

    def backward(self):
        lib, st, B, cv = self.lib, self._st(), self.B, self.convs
        self.grads.zero_()
        # final_conv
        g8 = self._buf("g8", (B,) + self.lv[0] + (64,))
        _lib.count_launch()
        check(lib.nrrt_head_backward(_pp(self.dlogits), _pp(self.b8), 64, _pp(self.params, self.head_w), self.M, 64, _pp(g8), 64,
                                     _pp(self.grads, self.head_w), _pp(self.grads, self.head_b), st))
        # conv8.2: g8 is already masked by its ReLU
        c = cv["c8.2"]
        self._relu_bwd(g8, 0, 64, None, 0, None, c.b_off)
        self._wgrad(c, g8, 0, self.a8, 0)
        d_a8 = self._buf("dx_c8.2", (B,) + self.lv[0] + (128,))
        c.dgrad.run(lib, st, g8, 0, (1,) + self.lv[0], d_a8, 0)
        dcat8 = self._conv_relu_bwd(cv["c8.0"], d_a8, self.a8, self.cat8, 0)
        d_b7 = self._convT_bwd(cv["up8"], dcat8, self.b7, 1)
        d_a7 = self._conv_relu_bwd(cv["c7.2"], d_b7, self.b7, self.a7, 1)
        dcat7 = self._conv_relu_bwd(cv["c7.0"], d_a7, self.a7, self.cat7, 1)
        d_b6 = self._convT_bwd(cv["up7"], dcat7, self.b6, 2)
        d_a6 = self._conv_relu_bwd(cv["c6.2"], d_b6, self.b6, self.a6, 2)
        dcat6 = self._conv_relu_bwd(cv["c6.0"], d_a6, self.a6, self.cat6, 2)
        dc5 = self._convT_bwd(cv["up6"], dcat6, self.c5, 3)
        # coordinate branch
        _lib.count_launch(12)
        for k, (buf, off, ch, lv) in enumerate(((dcat8, 128, 64, 0), (dcat7, 256, 128, 1), (dcat6, 768, 256, 2), (dc5, 512, 512, 3))):
            self.dfeats[k].zero_()
            check(lib.nrrt_coords_fill_backward(_pp(buf, off), buf.shape[-1], B, self.lv[lv][0], self.lv[lv][1], ch, _pp(self.dfeats[k]), st))
        for k in (3, 2, 1, 0):
            w_off, b_off, cin, cout = self.coords[k]
            f_in = self.feats[k - 1] if k > 0 else self.coords_in
            check(lib.nrrt_coords_layer_backward(_pp(self.dfeats[k]), _pp(self.feats[k]), _pp(f_in), _pp(self.params, w_off), B, cin, cout,
                                                 _pp(self.grads, w_off), _pp(self.grads, b_off), _pp(self.dfeats[k - 1]) if k > 0 else None, st))
        # encoder
        skips = {3: (dcat6, 512), 2: (dcat7, 128), 1: (dcat8, 64)}
        dy, dy_off = dc5, 0
        for si in (3, 2, 1, 0):
            dy = self._block_bwd(si, 1, dy, dy_off)
            dy = self._block_bwd(si, 0, dy, 0, skip=skips.get(si))
            dy_off = 0
        # conv1[2] and conv1[0]
        dt0 = self._conv_relu_bwd(cv["conv1.2"], dy, self.x1, self.t0, 0)
        self._relu_bwd(dt0, 0, 64, self.t0, 0, dt0, None)
        _lib.count_launch()
        check(lib.nrrt_stem_wgrad(_pp(self.x_in), _pp(dt0), 64, B, self.stem_cin, self.H, self.W, self.stem_cout, _pp(self.grads, self.stem_w),
                                  _pp(self.grads, self.stem_b), st))

    # -------------------------------------------------------------------------------------------------------------- step
    def set_targets(self, masks_u8=None, target=None):
        """masks_u8: uint8 [B, H, W] mask images (normalised per sample like MapsDataset, train.py:47-50), or target fp32 in [0, 1]."""
        if target is not None:
            self.target.copy_(target.reshape(self.target.shape))
            return
        _lib.count_launch()
        self._masks = masks_u8.to(self.dev).contiguous()     # kept alive: the call is asynchronous
        check(self.lib.nrrt_mask_targets(_pp(self._masks), self.B, self.H * self.W, _pp(self.target), self._st()))

    def loss_and_grad(self, with_grad=True):
        _lib.count_launch()
        check(self.lib.nrrt_bce_with_logits(_pp(self.logits), _pp(self.params, self.head_b), _pp(self.target), self.M,
                                            self.loss_scale / self.M, _pp(self.scale_state) if self.dynamic_scale else None,
                                            _pp(self.dlogits) if with_grad else None, _pp(self.loss_sum), self._st()))

    def optimizer_step(self):
        self.step_count += 1
        if self.dp:
            import torch.distributed as dist
            dist.all_reduce(self.grads)          # a non-finite gradient on any rank shows up on all of them: the skip is collective
        st = self._st()
        if self.dynamic_scale:
            _lib.count_launch(3)
            check(self.lib.nrrt_grad_check(_pp(self.grads), self.n_params, _pp(self.scale_state), st))
            check(self.lib.nrrt_adam_step(_pp(self.params), _pp(self.grads), _pp(self.exp_avg), _pp(self.exp_avg_sq), self.n_params, 1,
                                          self.lr, self.betas[0], self.betas[1], self.eps, 1.0 / (self.M * self.world), _pp(self.scale_state), st))
            check(self.lib.nrrt_loss_scale_update(_pp(self.scale_state), 200, 65536.0, 2.0 ** -14, st))
        else:
            _lib.count_launch()
            check(self.lib.nrrt_adam_step(_pp(self.params), _pp(self.grads), _pp(self.exp_avg), _pp(self.exp_avg_sq), self.n_params,
                                          self.step_count, self.lr, self.betas[0], self.betas[1], self.eps,
                                          1.0 / (self.loss_scale * self.world), None, st))
        self._refresh_weights()

    def training_step(self, batch, batch_idx=0, masks_u8=None):
        """UNet_cooler.training_step (train.py:205-210) + the optimizer step Lightning runs after it.  batch = (x, y, coords) as
        MapsDataset yields them (y fp32 [B, 1, H, W] in [0, 1]), or pass y = None and masks_u8.  Returns the loss as a device
        scalar tensor (float64; no host synchronisation)."""
        x, y, coords = batch
        self.forward(x.to(self.dev, torch.float32).contiguous(), coords.to(self.dev))
        self.set_targets(masks_u8=masks_u8, target=None if y is None else y.to(self.dev, torch.float32))
        self.loss_and_grad(True)
        self.backward()
        self.optimizer_step()
        return self.loss_sum / self.M

    def validation_step(self, batch, batch_idx=0, masks_u8=None):
        """UNet_cooler.validation_step (train.py:212-217) as Lightning runs it: model.eval(), i.e. BatchNorm with its running
        statistics, no gradient, nothing updated.  Returns the loss as a device scalar."""
        x, y, coords = batch
        self.training = False
        try:
            self.forward(x.to(self.dev, torch.float32).contiguous(), coords.to(self.dev))
        finally:
            self.training = True
        self.set_targets(masks_u8=masks_u8, target=None if y is None else y.to(self.dev, torch.float32))
        self.loss_and_grad(False)
        return self.loss_sum / self.M

    def scheduler_step(self, val_loss: float):
        """ReduceLROnPlateau(factor 0.5, patience 2) on val_loss (train.py:230-234): call once per validation epoch."""
        self.scheduler.lr = self.lr
        self.lr = self.scheduler.step(float(val_loss))
        return self.lr

    def recalibrate_batchnorm(self, batches):
        """Replace the BatchNorm running statistics by the plain average of the batch statistics over `batches` (an iterable of
        (x, coords)), weights untouched.  With 3 samples per step the exponential average (momentum 0.1) that training leaves
        behind is a noisy, lagging estimate; the eval-mode model (what the planner runs) is only as good as these numbers."""
        k = 0
        try:
            for x, coords in batches:
                k += 1
                self.bn_momentum = 1.0 / k          # running = running * (1 - 1/k) + batch / k: the cumulative mean
                self.forward(x.to(self.dev, torch.float32).contiguous(), coords.to(self.dev))
        finally:
            self.bn_momentum = BN_MOMENTUM
        return k

    def fit(self, train_loader, val_loader=None, max_epochs: int = 100, patience: int = 10, log=None):
        """The loop `pl.Trainer(max_epochs=100, callbacks=[ModelCheckpoint, EarlyStopping(monitor='val_loss', patience=10)]).fit(...)`
        runs around the model (train.py:289-310): per epoch every training batch through training_step, the validation set through
        validation_step (eval mode), ReduceLROnPlateau on the mean val_loss, the best state kept, early stop after `patience`
        epochs without improvement; then the best state is restored and exported into the module.  Loaders yield
        (x, y, coords) like MapsDataset; batches whose size is not the trainer's are skipped (the buffers are sized once).
        Returns the history [(epoch, train_loss, val_loss, lr), ...]."""
        best, best_state, bad, history = math.inf, None, 0, []
        for epoch in range(max_epochs):
            losses = []
            for batch in train_loader:
                if batch[0].shape[0] != self.B:
                    continue
                losses.append(self.training_step(batch))
            train_loss = float(torch.stack(losses).mean().item()) if losses else float("nan")
            val_loss = train_loss
            if val_loader is not None:
                vals = [self.validation_step(b) for b in val_loader if b[0].shape[0] == self.B]
                if vals:
                    val_loss = float(torch.stack(vals).mean().item())
            lr = self.scheduler_step(val_loss)
            history.append((epoch, train_loss, val_loss, lr))
            if log:
                log("epoch %d: train_loss %.5f val_loss %.5f lr %g" % (epoch, train_loss, val_loss, lr))
            if val_loss < best:
                best, bad = val_loss, 0
                best_state = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in self.state_dict().items()}
                best_state["bn"] = {k: (a.cpu(), b.cpu()) for k, (a, b) in best_state["bn"].items()}
            else:
                bad += 1
                if bad >= patience:
                    break
        if best_state is not None:
            keep_lr, keep_sched = self.lr, self.scheduler.state_dict()
            self.load_state_dict(best_state)
            self.lr = keep_lr
            self.scheduler.load_state_dict(keep_sched)
        self.export_to(self.model)
        return history

    # -------------------------------------------------------------------------------------------------------- checkpoints
    def state_dict(self):
        """Everything a run needs to continue (what Lightning's ModelCheckpoint keeps next to the model's own state_dict): the flat
        parameter buffer, Adam moments, BatchNorm running statistics, step counters, the loss-scale state, the rate and scheduler."""
        return {"layout": [self.B, self.H, self.W, self.n_params], "params": self.params.clone(), "exp_avg": self.exp_avg.clone(),
                "exp_avg_sq": self.exp_avg_sq.clone(), "scale_state": self.scale_state.clone(), "step_count": self.step_count,
                "lr": self.lr, "scheduler": self.scheduler.state_dict(),
                "bn": {name: (cv.bn[2].clone(), cv.bn[3].clone()) for name, cv in self.convs.items() if cv.bn is not None}}

    def load_state_dict(self, d):
        if int(d["layout"][3]) != self.n_params:
            raise ValueError("checkpoint was written for a different parameter layout")
        with torch.no_grad():
            self.params.copy_(d["params"].to(self.dev))
            self.exp_avg.copy_(d["exp_avg"].to(self.dev))
            self.exp_avg_sq.copy_(d["exp_avg_sq"].to(self.dev))
            self.scale_state.copy_(d["scale_state"].to(self.dev))
            for name, (rm, rv) in d["bn"].items():
                self.convs[name].bn[2].copy_(rm.to(self.dev))
                self.convs[name].bn[3].copy_(rv.to(self.dev))
        self.step_count, self.lr = int(d["step_count"]), float(d["lr"])
        self.scheduler.load_state_dict(d["scheduler"])
        self._refresh_weights()

    def flops_per_step(self):
        """Algorithmic tensor-core FLOPs of one training step: forward + data gradient + weight gradient of every convolution
        (3 x the forward's 2 x MACs, the first conv's data gradient excepted)."""
        f = 0.0
        for cv in self.convs.values():
            lv = self._conv_level(cv)
            px = self.B * self.lv[lv][0] * self.lv[lv][1]
            f += 3 * 2.0 * px * cv.rows * cv.taps * cv.cin
        return f

    def _conv_level(self, cv):
        """Level of the grid the layer's MACs are counted on (output grid; ConvTranspose: its input grid)."""
        n = cv.name
        if n.startswith("enc"):
            return int(n[3])
        return {"conv1.2": 0, "up6": 3, "c6.0": 2, "c6.2": 2, "up7": 2, "c7.0": 1, "c7.2": 1, "up8": 1, "c8.0": 0, "c8.2": 0}[n]


def generated_training_set(n_maps: int, first_map: int = 5000, size: int = 512, device="cuda:0", timings: dict = None):
    """The reference's data pipeline for n_maps maps x 10 tasks, entirely on the device: generate_maps.create_map +
    generate_tasks.draw_start_and_finish (csrc/generate.cu, the reference's `random` streams: map m is seed 1000 + m / 2000 + m),
    generate_paths.astar (csrc/gridpath.cu: one shortest path per task) and generate_paths.visualize_path (csrc/masks.cu).
    Returns (DeviceWorkload, masks uint8 [tasks, size, size]); `timings` (optional dict) receives the milliseconds of the three
    stages (CUDA events)."""
    from . import engine, generate_paths, workload
    dev = torch.device(device)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    wl = workload.make_on_device(2, n_maps * 10, first_map=first_map, size=size, device=dev)
    ev[1].record()
    bits = engine.pack_occupancy(wl.occ_maps, 2, dev)
    _, _, paths, path_n = engine.grid_paths(bits, wl.task_to_map, wl.starts, wl.goals, (size, size), path_cap=4 * size)
    ev[2].record()
    masks = generate_paths.path_masks(wl.occ_maps, wl.task_to_map, paths, path_n, dev)
    ev[3].record()
    if timings is not None:
        torch.cuda.synchronize(dev)
        n = n_maps * 10
        timings.update({"maps_and_tasks_ms": ev[0].elapsed_time(ev[1]), "shortest_paths_ms": ev[1].elapsed_time(ev[2]),
                        "masks_ms": ev[2].elapsed_time(ev[3]), "paths_per_s": n / ev[1].elapsed_time(ev[2]) * 1e3,
                        "masks_per_s": n / ev[2].elapsed_time(ev[3]) * 1e3,
                        "masks_algorithmic_GBps": n * size * size * 2 / ev[2].elapsed_time(ev[3]) / 1e6})
    return wl, masks


def fit_on_generated(model, n_maps: int, steps: int, first_map: int = 5000, batch: int = 3, size: int = 512, device="cuda:0", seed: int = 0,
                     log=None, schedule: str = "steps", recalibrate: int = 64, checkpoint_every: int = 1000):
    """Train `model` (UNet_cooler) for `steps` optimisation steps of `batch` random samples of a generated training set; the
    learning rate is halved at 60 % and 80 % of the run (the reference halves it on validation plateaus, train.py:230); every
    `checkpoint_every` steps the eval-mode loss on held-out tasks decides which state is kept (ModelCheckpoint on val_loss,
    train.py:289-294); the kept state's BatchNorm statistics are recalibrated and it is written back into `model`.
    Returns a dict of run statistics."""
    import time
    dev = torch.device(device)
    t0 = time.time()
    stage_ms = {}
    wl, masks = generated_training_set(n_maps, first_map, size, dev, timings=stage_ms)
    torch.cuda.synchronize(dev)
    t_data = time.time() - t0
    tr = UNetTrainer(model, batch, size, size, dev)
    rng = np.random.default_rng(seed)
    n = int(masks.shape[0])
    n_val = max(batch * 16, n // 20)                    # the last 5 % of the generated tasks are held out for validation
    n_train = n - n_val

    def sample(lo, hi):
        idx = torch.from_numpy(lo + rng.choice(hi - lo, batch, replace=False)).to(dev)
        occ = wl.occ_maps[wl.task_to_map[idx].long()]
        return idx, (occ != 0).float()[:, None].expand(-1, 3, -1, -1).contiguous()      # MapsDataset's image tensor (train.py:41-44)

    def val_loss(k=16):
        vals = []
        for j in range(k):
            idx = torch.arange(n_train + j * batch, n_train + (j + 1) * batch, device=dev)
            occ = wl.occ_maps[wl.task_to_map[idx].long()]
            x = (occ != 0).float()[:, None].expand(-1, 3, -1, -1).contiguous()
            vals.append(tr.validation_step((x, None, wl.coords[idx]), masks_u8=masks[idx]))
        return float(torch.stack(vals).mean().item())

    run, curve, vcurve = [], [], []
    best, best_state = math.inf, None
    t0 = time.time()
    for step in range(1, steps + 1):
        if schedule == "cosine":   # 100 warm-up steps, then a cosine from 1e-3 down to 2e-5
            tr.lr = 1e-3 * min(1.0, step / 100.0) * max(0.02, 0.5 * (1.0 + math.cos(math.pi * step / steps)))
        else:
            tr.lr = 1e-3 if step <= 0.6 * steps else (5e-4 if step <= 0.8 * steps else 2.5e-4)
        idx, x = sample(0, n_train)
        run.append(tr.training_step((x, None, wl.coords[idx]), masks_u8=masks[idx]))
        if step % 250 == 0 or step == steps:
            curve.append((step, float(torch.stack(run).mean().item())))
            run = []
            if log:
                log("step %d: mean loss %.4f, lr %g, loss scale %g" % (step, curve[-1][1], tr.lr, float(tr.scale_state[0].item())))
        if step % checkpoint_every == 0 or step == steps:
            # ModelCheckpoint(monitor='val_loss') of the reference's Trainer (train.py:289-294): the eval-mode loss on the held-out
            # tasks decides which state is kept
            v = val_loss()
            vcurve.append((step, v))
            if v < best:
                best, best_state = v, {k: (t.clone() if torch.is_tensor(t) else t) for k, t in tr.state_dict().items()}
    t_train = time.time() - t0
    eval_before = vcurve[-1][1]
    if best_state is not None:
        tr.load_state_dict(best_state)
    if recalibrate > 0:
        tr.recalibrate_batchnorm(((x, wl.coords[idx]) for idx, x in (sample(0, n_train) for _ in range(recalibrate))))
    eval_after = val_loss()
    torch.cuda.synchronize(dev)
    taken = int(tr.scale_state[3].item())
    tr.export_to(model)
    return {"val_loss_last_step": eval_before, "val_loss_best": best, "val_loss_kept_state_recalibrated": eval_after, "val_curve": vcurve,
            "train_tasks": n_train, "val_tasks": n_val, "train_maps": n_maps, "steps": steps, "steps_taken": taken, "batch": batch, "data_seconds": t_data, "data_stages": stage_ms,
            "train_seconds": t_train, "samples_per_s": steps * batch / t_train, "loss_curve": curve}
