"""TEST INFRASTRUCTURE -- plain PyTorch fp32 restatement of the reference's training step, the checker of
mgr_sampling_cnn_b200/trainer.py (a floating-point path keeps a torch fp32 reference).  Only tests/, smoke() and bench.py's
CPU legs may import this.

Follows UNet_cooler.training_step (train.py:205-210): y_hat = self(x, coords) in train mode (batch-statistics BatchNorm2d
in the torchvision BasicBlocks), loss = nn.BCEWithLogitsLoss()(y_hat, y) (:113), then what Lightning does with the returned
loss: backward and one step of configure_optimizers()'s torch.optim.Adam(self.parameters(), lr=1e-3) (:228-229).
Pinned against the unmodified reference class by tests/test_train_oracle.py on tests/golden/train2d_seed0.npz
(oracle/gen_golden_train.py)."""
import torch
import torch.nn.functional as F

from . import cnn_oracle


def is_param(key):
    return not (key.endswith("running_mean") or key.endswith("running_var") or key.endswith("num_batches_tracked"))


def forward_backward(sd, x, y, coords):
    """sd: UNet_cooler state_dict (fp32 CPU tensors).  Returns (loss float, grads {key: tensor} for the parameters the forward
    uses, new running statistics {key: tensor}).  sd is not modified."""
    work = {}
    for k, v in sd.items():
        if k.startswith("base_model.") or k.endswith("num_batches_tracked"):
            continue
        work[k] = v.detach().clone().float().requires_grad_(is_param(k))
    cnn_oracle._TRAIN[0] = True
    try:
        with cnn_oracle.strict_fp32():
            logits = cnn_oracle._forward(work, x, coords, 2)
            loss = F.binary_cross_entropy_with_logits(logits, y)
            loss.backward()
    finally:
        cnn_oracle._TRAIN[0] = False
    grads = {k: v.grad for k, v in work.items() if v.requires_grad and v.grad is not None}
    stats = {k: v.detach() for k, v in work.items() if not v.requires_grad}
    return float(loss.detach()), grads, stats, logits.detach()


def adam_step(sd, grads, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, state=None, step=1):
    """torch.optim.Adam's update written out (first step from a zero state by default); returns ({key: new value}, state)."""
    state = state if state is not None else {}
    out = {}
    for k, g in grads.items():
        m, v = state.get(k, (torch.zeros_like(g), torch.zeros_like(g)))
        m = betas[0] * m + (1 - betas[0]) * g
        v = betas[1] * v + (1 - betas[1]) * g * g
        state[k] = (m, v)
        bc1, bc2 = 1 - betas[0] ** step, 1 - betas[1] ** step
        out[k] = sd[k].float() - (lr / bc1) * m / (v.sqrt() / (bc2 ** 0.5) + eps)
    return out, state
