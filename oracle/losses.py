"""Oracle: training losses + optimiser on torch-CPU (TEST INFRASTRUCTURE -- see
oracle/__init__.py).  PINNED for focal / smooth-L1 values and gradients on fixtures produced by
executing the reference's utils/tpu.py (tests/golden/make_golden_losses.py, tests/
test_oracle_losses_golden.py); the SGD step is parity unpinned (TF semantics, App. A.9).

Restates /root/reference/utils/tpu.py:26-81 (tpu_smooth_l1, delta = lambda_ = 1),
:84-155 (tpu_focal; call site train_tpu.py:259 uses alpha=.25, gamma=1.5) and the
Keras SGD of train_tpu.py:268-269.  keras.backend.binary_crossentropy of the
pinned TF (1.15/2.0): clip p to [1e-7, 1-1e-7], logits = log(p/(1-p)),
sigmoid-CE(logits) = max(z,0) - z*t + log1p(exp(-|z|))   (SURVEY Appendix A.8).
"""
import torch


def smooth_l1(y_true, y_pred, lambda_=1.0):
    target, state = y_true[:, :, :-1], y_true[:, :, -1]
    diff = torch.abs(y_pred - target)
    loss = torch.where(diff > lambda_, diff - 0.5, 0.5 * diff ** 2)
    fg = (state == 1).to(loss.dtype)
    norm = torch.clamp(fg.sum(), min=1.0)
    return (loss.sum(dim=2) * fg).sum() / norm


def binary_crossentropy(t, p, eps=1e-7):
    p = torch.clamp(p, eps, 1 - eps)
    z = torch.log(p / (1 - p))
    return torch.clamp(z, min=0) - z * t + torch.log1p(torch.exp(-torch.abs(z)))


def focal(y_true, y_pred, alpha=0.25, gamma=1.5):
    state = y_true[:, :, -1]
    t = y_true[:, :, :-1]
    is_fg = t == 1
    alpha_f = torch.where(is_fg, torch.full_like(t, alpha), torch.full_like(t, 1 - alpha))
    fw = torch.where(is_fg, 1 - y_pred, y_pred)
    fw = alpha_f * fw ** gamma
    cls = fw * binary_crossentropy(t, y_pred)
    not_ign = (state != -1).to(cls.dtype)
    total = (cls.sum(dim=2) * not_ign).sum()
    norm = torch.clamp((state == 1).to(cls.dtype).sum(), min=1.0)
    return total / norm


def sgd_momentum_step(w, g, v, lr=0.01, decay=4e-5, momentum=0.9, iteration=0):
    """Keras SGD: lr_t = lr/(1+decay*iter); v = m*v - lr_t*g; w += v  (in place)."""
    lr_t = lr / (1.0 + decay * iteration)
    v.mul_(momentum).sub_(lr_t * g)
    w.add_(v)
