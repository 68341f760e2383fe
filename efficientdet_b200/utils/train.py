"""Checkpoint callback and weight loading helpers of the reference's training scripts
(utils/train.py:10-35 load_weights, :66-92 CheckpointSaver).  Checkpoints are .npz files holding the
weights under their Keras names plus the optimizer state (SGD velocity, iteration count) -- the
tf.train.Checkpoint(optimizer=..., model=...) pair of the reference."""
import glob
import os

import numpy as np


class CheckpointSaver:
    """keras-style callback: writes `checkpoint_prefix.format(epoch=epoch, **logs) + '.npz'` at the end
    of every epoch (utils/train.py:66-92)."""

    def __init__(self, checkpoint_prefix, log=None):
        self.checkpoint_prefix = checkpoint_prefix
        self.log = log
        self.model = None

    def set_model(self, model):
        self.model = model

    def on_epoch_end(self, epoch, logs=None):
        path = self.checkpoint_prefix.format(epoch=epoch, **(logs or {}))
        if self.log:
            self.log.debug("saving to %s", path)
        save_checkpoint(self.model, path)
        if self.log:
            self.log.debug("saved %s", path)


def save_checkpoint(model, path):
    d = {"model/" + k: v for k, v in model.get_weights_dict().items()}
    opt = getattr(model, "optimizer", None)
    if opt is not None:
        model.net.ensure_grad_buffers()
        d["optimizer/iterations"] = np.array(opt.iterations, np.int64)
        d["optimizer/lr"] = np.array(opt.lr, np.float64)
        d["optimizer/velocity"] = model.net.velocity.detach().cpu().numpy()
    os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
    np.savez(path if path.endswith(".npz") else path + ".npz", **d)


def restore_checkpoint(model, path):
    """Inverse of save_checkpoint (train_tpu.py:284-304 restores the newest checkpoint of a directory:
    pass a directory to get that behaviour)."""
    if os.path.isdir(path):
        files = sorted(glob.glob(os.path.join(path, "*.npz")), key=os.path.getmtime)
        if not files:
            raise FileNotFoundError("no checkpoint in %s" % path)
        path = files[-1]
    d = dict(np.load(path))
    model.set_weights_dict({k[len("model/"):]: v for k, v in d.items() if k.startswith("model/")}, strict=True)
    opt = getattr(model, "optimizer", None)
    if opt is not None and "optimizer/iterations" in d:
        import torch
        model.net.ensure_grad_buffers()
        opt.iterations = int(d["optimizer/iterations"])
        opt.lr = float(d["optimizer/lr"])
        model.net.velocity.copy_(torch.from_numpy(d["optimizer/velocity"]).to(model.net.velocity.device))
    return path


def load_weights(model, path, by_name=True):
    """utils/train.py:10-35: weights-only load (.npz keyed by Keras names)."""
    return model.load_weights(path, by_name=by_name)
