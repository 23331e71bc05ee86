"""CPU: pins oracle/train_oracle.py (torch fp32 restatement of the reference's training step) against one optimisation step
of the unmodified reference model (tests/golden/train2d_seed0.npz, oracle/gen_golden_train.py)."""
import numpy as np
import torch

from oracle import train_oracle
from tests import helpers as H


def test_train_oracle_matches_reference_step():
    from mgr_sampling_cnn_b200 import train
    x, y, coords, _, g = H.load_train_batch()
    torch.manual_seed(0)
    sd = train.UNet_cooler().state_dict()
    loss, grads, stats, _ = train_oracle.forward_backward(sd, torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(coords))
    assert abs(loss - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    names = [str(n) for n in g["names"]]
    assert sorted(names) == sorted(grads.keys())          # the same parameters receive a gradient (base_model.* stem / fc do not)
    for k, gn, gs in zip(names, g["grad_norm"], g["grad_sum"]):
        assert abs(float(grads[k].double().norm()) - gn) <= 1e-4 * gn + 1e-12, k
    new, _ = train_oracle.adam_step(sd, grads)
    for key in g.keys():
        if key.startswith("grad/"):
            k = key[5:]
            ref = g[key]
            assert np.abs(grads[k].numpy() - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-12, k
        elif key.startswith("stat/"):
            assert np.allclose(stats[key[5:]].numpy(), g[key], rtol=1e-5, atol=1e-7), key
    psum = dict(zip(names, g["param_sum_after"]))
    for k in ("conv8.2.weight", "conv6.0.weight", "conv2.0.conv1.weight"):
        assert abs(float(new[k].double().sum()) - psum[k]) <= 2e-3 * new[k].numel() * 1e-3 + 1e-6, k   # Adam's first step is +-lr


def test_oracle_adam_matches_torch_optim():
    """train_oracle.adam_step (used as the multi-step checker of the trainer) == torch.optim.Adam(lr=1e-3) over three steps."""
    g = torch.Generator().manual_seed(0)
    p0 = {"a": torch.randn(40, generator=g), "b": torch.randn(3, 5, generator=g)}
    params = {k: torch.nn.Parameter(v.clone()) for k, v in p0.items()}
    opt = torch.optim.Adam(params.values(), lr=1e-3)
    sd, state = {k: v.clone() for k, v in p0.items()}, {}
    for step in (1, 2, 3):
        grads = {k: torch.randn(v.shape, generator=g) for k, v in p0.items()}
        for k in params:
            params[k].grad = grads[k].clone()
        opt.step()
        new, state = train_oracle.adam_step(sd, grads, state=state, step=step)
        sd.update(new)
        for k in params:
            assert torch.allclose(sd[k], params[k].detach(), rtol=0, atol=2e-7), (k, step)
