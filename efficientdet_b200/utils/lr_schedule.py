"""Per-epoch learning-rate schedule of the reference: linear warm-up, then cosine decay
(utils/lr_schedule.py:5-68; used as a keras LearningRateScheduler callback in train.py:90-99).

The cosine part restates tf.keras.experimental.CosineDecay (TensorFlow is not a dependency here):
    decayed = (1 - alpha) * 0.5 * (1 + cos(pi * min(step, decay_steps) / decay_steps)) + alpha
    lr      = initial_learning_rate * decayed
"""
import math


class LearningRateScheduler:
    """Minimal keras.callbacks.LearningRateScheduler: `schedule(epoch_index, lr) -> lr` is applied to
    `model.optimizer.lr` at the beginning of every epoch by Model.fit(callbacks=[...])."""

    def __init__(self, schedule, verbose=0):
        self.schedule = schedule
        self.verbose = verbose
        self.model = None

    def set_model(self, model):
        self.model = model

    def on_epoch_begin(self, epoch, logs=None):
        opt = self.model.optimizer
        opt.lr = float(self.schedule(epoch, opt.lr))
        if self.verbose:
            print("Epoch %05d: LearningRateScheduler setting learning rate to %s." % (epoch + 1, opt.lr))


def cosine_decay(step, initial_learning_rate, decay_steps, alpha=0.0):
    step = min(float(step), float(decay_steps))
    decayed = (1.0 - alpha) * 0.5 * (1.0 + math.cos(math.pi * step / float(decay_steps))) + alpha
    return initial_learning_rate * decayed


def get_cosine_decay_with_linear_warmup(total_epochs, current_epoch=0, learning_rate_start=0.0,
                                        learning_rate_max=.08, warmup_percent=0.05, alpha=0.001,
                                        verbose=False):
    """Same signature and values as the reference (utils/lr_schedule.py:5-68): the rate grows linearly
    from learning_rate_start to learning_rate_max over the first warmup_percent of the epochs, then
    follows a cosine down to alpha * learning_rate_max.  Epoch numbers are 1-based inside."""
    if learning_rate_start > learning_rate_max:
        raise ValueError("learning_rate_start must be < learning_rate_max")
    switch_epoch = warmup_percent * total_epochs
    linear_slope = (learning_rate_max - learning_rate_start) / switch_epoch

    def scheduler(epoch_index, lr):
        epoch_number = epoch_index + 1
        if epoch_number > switch_epoch:
            lr = cosine_decay(epoch_number - switch_epoch, learning_rate_max, total_epochs - switch_epoch, alpha)
        else:
            lr = learning_rate_start + linear_slope * epoch_number
        if verbose:
            print(" Learning rate set to", lr, "for epoch", epoch_number)
        return lr

    return LearningRateScheduler(scheduler)
