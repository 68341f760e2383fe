"""SGD -- mirror of the tf.keras.optimizers.SGD the reference compiles with
(train_tpu.py:268-269 SGD(lr=0.01, decay=4e-5, momentum=0.9); train.py:346 lr=0.08)."""


class SGD:
    def __init__(self, lr=0.01, momentum=0.0, decay=0.0, nesterov=False, learning_rate=None, **kw):
        if nesterov:
            raise NotImplementedError("nesterov momentum is not used by the reference")
        self.lr = float(learning_rate if learning_rate is not None else lr)
        self.momentum = float(momentum)
        self.decay = float(decay)
        self.iterations = 0

    def current_lr(self):
        """lr_t = lr / (1 + decay * iterations)   (keras `decay` = LR time decay, not weight decay)"""
        return self.lr / (1.0 + self.decay * self.iterations)

    def get_config(self):
        return {"lr": self.lr, "momentum": self.momentum, "decay": self.decay, "nesterov": False}
