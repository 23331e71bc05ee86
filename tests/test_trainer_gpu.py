"""The training step on the B200 kernels (mgr_sampling_cnn_b200/trainer.py) against the reference's own step recorded in
tests/golden/train2d_seed0.npz (loss, gradient norms, selected gradients, BatchNorm running statistics) and against the fp32
torch oracle (oracle/train_oracle.py) for every gradient tensor.

Tolerances.  The forward runs fp16 operands / fp32 accumulate like the inference path (logits <= 1e-2 abs), activation
gradients are fp16, parameter gradients fp32.  Decoder, coordinate-branch and head gradients (99.99 % of the gradient norm)
must agree to 5e-3 relative L2 (measured: <= 1.5e-3).  The random-init ENCODER's gradients are 1e3 - 1e7 times smaller and
ill-conditioned: in a pure fp32 autograd run, rounding the weights to fp16 moves them by 2 - 4 %, and fp16-sized relative
noise (3e-4 .. 1e-3) on the conv outputs moves them by 2 - 15 % (scripts/train_probe3.py, profiles/train_r2.md), so they
are held to 15 % here (measured: 0.1 - 11 %)."""
import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _encoder(key):
    return key.startswith(("conv1.", "conv2.", "conv3.", "conv4.", "conv5."))


def _setup():
    from mgr_sampling_cnn_b200 import train, trainer
    x, y, coords, masks, g = H.load_train_batch()
    torch.manual_seed(0)
    model = train.UNet_cooler()
    tr = trainer.UNetTrainer(model, 2, 512, 512, DEV)
    return model, tr, x, y, coords, masks, g


def test_training_step_matches_reference_and_oracle():
    from oracle import train_oracle
    model, tr, x, y, coords, masks, g = _setup()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    xd, cd = torch.from_numpy(x).to(DEV), torch.from_numpy(coords).to(DEV)
    tr.forward(xd, cd)
    tr.set_targets(masks_u8=torch.from_numpy(masks))
    assert torch.equal(tr.target.cpu(), torch.from_numpy(y[:, 0]))           # MapsDataset's normalisation, bit for bit
    tr.loss_and_grad(True)
    loss = float(tr.loss_sum.item()) / tr.M
    assert abs(loss - float(g["loss"])) <= 2e-3 * float(g["loss"]), (loss, float(g["loss"]))
    tr.backward()
    torch.cuda.synchronize()
    grads = {k: v.cpu() for k, v in tr.named(tr.grads, 1.0 / tr.loss_scale).items()}
    # (1) the reference's own numbers
    names = [str(n) for n in g["names"]]
    assert sorted(names) == sorted(grads.keys())
    worst = 0.0
    for k, gn in zip(names, g["grad_norm"]):
        rel = abs(float(grads[k].double().norm()) - gn) / gn
        worst = max(worst, rel)
        assert rel <= (0.10 if _encoder(k) else 5e-3), (k, rel)
    for key in g.keys():
        if key.startswith("grad/"):
            ref = torch.from_numpy(g[key])
            err = float((grads[key[5:]] - ref).norm() / ref.norm())
            assert err <= (0.15 if _encoder(key[5:]) else 5e-3), (key, err)
    # (2) every gradient tensor against the fp32 oracle
    oloss, ograds, ostats, ologits = train_oracle.forward_backward(sd, torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(coords))
    assert float((tr.logits.cpu() - ologits[:, 0]).abs().max()) <= 1e-2
    errs = {k: float((grads[k] - ograds[k]).norm() / ograds[k].norm()) for k in ograds}
    bad = {k: e for k, e in errs.items() if e > (0.15 if _encoder(k) else 5e-3)}
    assert not bad, bad
    print("gradient relative L2 errors: max %.3g (%s), median %.3g; norm error vs reference max %.3g" %
          (max(errs.values()), max(errs, key=errs.get), float(np.median(list(errs.values()))), worst))
    # (3) BatchNorm running statistics after one step
    for key in g.keys():
        if key.startswith("stat/"):
            name = key[5:]
            mod = name.rsplit(".", 1)[0]
            cv = [c for c in tr.convs.values() if c.bn is not None and c.bn[4] is model.get_submodule(mod)][0]
            got = cv.bn[2 if name.endswith("running_mean") else 3].cpu().numpy()
            assert np.allclose(got, g[key], rtol=2e-3, atol=2e-4), key
    # (4) Adam: the first step moves every weight by lr * sign(g) (up to eps); elements whose gradient is numerically zero may flip
    tr.optimizer_step()
    tr.export_to(model)
    new_sd = model.state_dict()
    for key in g.keys():
        if key.startswith("new/"):
            ref, got = g[key], new_sd[key[4:]].cpu().numpy()
            frac = float((np.abs(got - ref) > 2e-4).mean())
            assert frac <= (0.15 if _encoder(key[4:]) else 0.02), (key, frac)


def test_loss_decreases_and_exported_model_runs():
    model, tr, x, y, coords, masks, g = _setup()
    batch = (torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(coords))
    losses = [float(tr.training_step(batch).item()) for _ in range(8)]
    assert losses[-1] < 0.8 * losses[0], losses
    assert all(np.isfinite(losses))
    tr.export_to(model)
    model = model.to(DEV).eval()
    with torch.no_grad():
        out = model(batch[0].to(DEV), batch[2].to(DEV))          # the inference path (eval-mode BatchNorm, folded kernels)
    assert out.shape == (2, 1, 512, 512) and bool(torch.isfinite(out).all())


def test_three_steps_follow_the_oracle():
    """Three optimisation steps (Adam state, bias corrections from the device-side step counter, BatchNorm running statistics fed
    forward) against the fp32 torch oracle stepping the same batch: the loss trajectory must agree."""
    from oracle import train_oracle
    model, tr, x, y, coords, masks, g = _setup()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    tx, ty, tc = torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(coords)
    state, ref_losses, got_losses = {}, [], []
    for step in (1, 2, 3):
        loss, grads, stats, _ = train_oracle.forward_backward(sd, tx, ty, tc)
        ref_losses.append(loss)
        new, state = train_oracle.adam_step(sd, grads, state=state, step=step)
        sd.update(new)
        sd.update(stats)
        got_losses.append(float(tr.training_step((tx, ty, tc)).item()))
    assert int(tr.scale_state[3].item()) == 3                     # no step was skipped by the loss-scale logic
    for a, b in zip(got_losses, ref_losses):
        assert abs(a - b) <= 0.03 * abs(b), (got_losses, ref_losses)
    tr.export_to(model)
    new_sd = model.state_dict()
    for k in ("conv8.2.weight", "conv7.0.weight", "upconv8.weight", "final_conv.weight"):
        # after three +-lr-sized Adam moves the weights agree far inside one move (the decoder gradients agree to 1e-3)
        assert float((new_sd[k].cpu() - sd[k]).abs().mean()) <= 2e-4, k
    for k in ("conv2.0.bn1.running_mean", "conv5.1.bn2.running_var"):
        # (the encoder's weights have taken up to three +-1e-3 moves whose signs follow ill-conditioned gradients: the statistics
        # agree to the size of those moves, not to rounding)
        dev_rel = float((new_sd[k].cpu() - sd[k]).abs().mean() / sd[k].abs().mean())
        assert dev_rel <= 0.15, (k, dev_rel)


def test_small_images_match_the_oracle():
    """128 x 128 (coarsest level 16 x 16: the weight-gradient kernel's 64-pixel blocks are partly outside the image): loss and
    gradients against the fp32 oracle.  With 16 x fewer pixels per gradient than at 512 x 512 the fp16 rounding noise averages
    out 4 x less: the layers fed by the deepest features (upconv6, conv_coords_512) sit at 1 - 1.5 % here against 0.15 % there."""
    from mgr_sampling_cnn_b200 import train, trainer
    from oracle import train_oracle
    rng = np.random.default_rng(8)
    B, S = 2, 128
    x = np.repeat((rng.random((B, 1, S, S)) > 0.25).astype(np.float32), 3, axis=1)
    y = (rng.random((B, 1, S, S)) > 0.9).astype(np.float32)
    coords = rng.integers(0, S, (B, 1, 2, 2)).astype(np.float32)
    torch.manual_seed(3)
    model = train.UNet_cooler()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    tr = trainer.UNetTrainer(model, B, S, S, DEV)
    tr.forward(torch.from_numpy(x).to(DEV), torch.from_numpy(coords).to(DEV))
    tr.set_targets(target=torch.from_numpy(y).to(DEV))
    tr.loss_and_grad(True)
    tr.backward()
    torch.cuda.synchronize()
    oloss, ograds, _, _ = train_oracle.forward_backward(sd, torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(coords))
    loss = float(tr.loss_sum.item()) / tr.M
    assert abs(loss - oloss) <= 2e-3 * oloss
    grads = {k: v.cpu() for k, v in tr.named(tr.grads, 1.0 / tr.loss_scale).items()}
    errs = {k: float((grads[k] - ref).norm() / ref.norm()) for k, ref in ograds.items()}
    bad = {k: round(e, 4) for k, e in errs.items() if e > (0.25 if _encoder(k) else 2.5e-2)}
    assert not bad, bad


def test_checkpoint_resume():
    """UNetTrainer.state_dict / load_state_dict: a fresh trainer that loads a checkpoint holds the same parameters, optimizer
    moments, running statistics, counters and loss scale, and its next step moves the weights like the original's next step."""
    from mgr_sampling_cnn_b200 import train, trainer
    model, tr, x, y, coords, masks, g = _setup()
    batch = (torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(coords))
    for _ in range(2):
        tr.training_step(batch)
    tr.lr = 7e-4
    ck = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in tr.state_dict().items()}        # as torch.save / load would hand it back
    torch.manual_seed(99)
    tr2 = trainer.UNetTrainer(train.UNet_cooler(), 2, 512, 512, DEV)                               # different random init
    tr2.load_state_dict(ck)
    assert torch.equal(tr2.params, tr.params) and torch.equal(tr2.exp_avg_sq, tr.exp_avg_sq) and torch.equal(tr2.params_h, tr.params_h)
    assert tr2.scale_state.tolist() == tr.scale_state.tolist() and tr2.lr == 7e-4 and tr2.step_count == 2
    cv = "enc2.1.c2"
    assert torch.equal(tr2.convs[cv].bn[2], tr.convs[cv].bn[2]) and torch.equal(tr2.convs[cv].wd, tr.convs[cv].wd)
    before = tr.params.clone()
    l1, l2 = float(tr.training_step(batch).item()), float(tr2.training_step(batch).item())
    assert abs(l1 - l2) <= 1e-3 * abs(l1)                   # same weights, same batch (reductions are atomic: not bit-identical)
    d1, d2 = tr.params - before, tr2.params - before
    assert float((d1 - d2).abs().max()) <= 2.0 * 7e-4 and float((d1 * d2 < 0).float().mean()) < 0.05


def test_validation_step_is_eval_mode():
    """validation_step = the loss under model.eval() (running statistics), like Lightning's validation loop: it must equal the
    fp32 oracle's eval-mode loss on the exported weights and must not move any state."""
    import torch.nn.functional as F
    from oracle import cnn_oracle
    model, tr, x, y, coords, masks, g = _setup()
    batch = (torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(coords))
    for _ in range(3):
        tr.training_step(batch)
    before = (tr.params.clone(), tr.convs["enc1.0.c1"].bn[2].clone(), tr.scale_state.clone())
    val = float(tr.validation_step(batch).item())
    assert torch.equal(tr.params, before[0]) and torch.equal(tr.convs["enc1.0.c1"].bn[2], before[1]) and torch.equal(tr.scale_state, before[2])
    tr.export_to(model)
    with torch.no_grad():
        ref = float(F.binary_cross_entropy_with_logits(cnn_oracle.forward(model.state_dict(), batch[0], batch[2], 2), batch[1]))
    assert abs(val - ref) <= 1e-2 * abs(ref), (val, ref)
    train_mode = float(tr.training_step(batch).item())           # batch statistics give a different number after three steps
    assert abs(train_mode - val) > 1e-4


def test_fit_loop_early_stopping_and_best_state():
    """UNetTrainer.fit: epochs over loaders, eval-mode validation, plateau scheduler, early stopping, best state restored and
    exported (the loop Lightning runs for train.py:289-310)."""
    from mgr_sampling_cnn_b200 import train, trainer
    x, y, coords, masks, g = H.load_train_batch()
    data = [(torch.from_numpy(x[i:i + 1]), torch.from_numpy(y[i:i + 1]), torch.from_numpy(coords[i:i + 1])) for i in (0, 1)]
    torch.manual_seed(0)
    model = train.UNet_cooler()
    tr = trainer.UNetTrainer(model, 1, 512, 512, DEV)
    ragged = data + [(torch.zeros(2, 3, 512, 512), torch.zeros(2, 1, 512, 512), torch.zeros(2, 1, 2, 2))]   # wrong batch size: skipped
    hist = tr.fit(ragged * 3, data, max_epochs=4, patience=2)
    assert 1 <= len(hist) <= 4 and all(np.isfinite(h[1]) and np.isfinite(h[2]) for h in hist)
    assert tr.step_count >= 6                                          # 6 usable training batches per epoch
    best = min(h[2] for h in hist)
    val = float(np.mean([float(tr.validation_step(b).item()) for b in data]))
    assert abs(val - best) <= 2e-2 * best                              # the restored state is the best epoch's
    out = model.to(DEV).eval()(data[0][0].to(DEV), data[0][2].to(DEV))   # ... and it is what export_to left in the module
    assert bool(torch.isfinite(out).all())


def test_fit_on_generated_runs_the_whole_device_pipeline():
    """trainer.fit_on_generated: maps, tasks, shortest paths and masks generated on the device, training steps, the best
    held-out state kept, BatchNorm recalibrated, the module updated (what bench.py's trained_guidance section calls)."""
    from mgr_sampling_cnn_b200 import train, trainer
    torch.manual_seed(0)
    model = train.UNet_cooler()
    before = model.conv8[0].weight.detach().clone()
    st = trainer.fit_on_generated(model, 12, 24, first_map=7000, device=DEV, recalibrate=4, checkpoint_every=8)
    assert st["steps"] == 24 and st["train_tasks"] + st["val_tasks"] == 120 and len(st["val_curve"]) == 3
    assert np.isfinite(st["val_loss_best"]) and st["val_loss_best"] <= st["val_curve"][0][1] + 1e-12
    assert st["data_stages"]["masks_per_s"] > 0 and st["loss_curve"][-1][0] == 24
    assert not torch.equal(model.conv8[0].weight.detach().cpu(), before)
    out = model.to(DEV).eval()(torch.ones(1, 3, 512, 512, device=DEV), torch.tensor([[[[10., 20.], [400., 300.]]]], device=DEV))
    assert bool(torch.isfinite(out).all())
