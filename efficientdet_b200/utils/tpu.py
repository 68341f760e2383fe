"""Loss factories -- mirror of the reference's working losses utils/tpu.py:26-81
(tpu_smooth_l1) and :84-155 (tpu_focal).  The returned objects carry the hyper-parameters; the
arithmetic runs in effdet_detection_losses (csrc/losses.cu), forward + backward fused."""


class _Loss:
    def __repr__(self):
        return "%s(%s)" % (type(self).__name__, ", ".join("%s=%r" % kv for kv in vars(self).items()))


class SmoothL1(_Loss):
    def __init__(self, lambda_=1):
        self.lambda_ = float(lambda_)


class Focal(_Loss):
    def __init__(self, alpha=0.25, gamma=2.0):
        self.alpha, self.gamma = float(alpha), float(gamma)


def tpu_smooth_l1(lambda_=1):
    return SmoothL1(lambda_)


def tpu_focal(alpha=0.25, gamma=2.0):
    return Focal(alpha, gamma)
