"""TF-1.15 ``tf.compat.v1.train.AdamOptimizer`` restated (created at ``styler_3p.py:320-323``,
``styler_2p.py:251-254`` with default beta1=.9, beta2=.999, epsilon=1e-8).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

TF's ApplyAdam:  lr_t = lr * sqrt(1-b2^t) / (1-b1^t);  m += (g-m)(1-b1);  v += (g^2-v)(1-b2);
var -= lr_t * m / (sqrt(v) + eps)  -- epsilon is OUTSIDE the bias correction, unlike
torch.optim.Adam.  The power accumulators are fp32 variables updated after each apply.
"""
import numpy as np
import torch


class TFAdam:
    def __init__(self, beta1=0.9, beta2=0.999, eps=1e-8):
        self.b1, self.b2, self.eps = beta1, beta2, eps
        self.m = None
        self.v = None
        # beta power accumulators start at beta (TF creates them as beta1/beta2 and the
        # first apply uses them before multiplying).
        self.b1p = np.float32(beta1)
        self.b2p = np.float32(beta2)

    def step(self, var, grad, lr):
        if self.m is None:
            self.m = torch.zeros_like(var)
            self.v = torch.zeros_like(var)
        lr_t = np.float32(lr) * np.sqrt(np.float32(1) - self.b2p) / (np.float32(1) - self.b1p)
        lr_t = float(lr_t)
        # TF's kernel forms T(1) - beta in the variable's dtype (training_ops.cc): in fp32 1 - 0.999f = 0.00099998713,
        # 1.3e-5 off the decimal value -- the CUDA kernel does the same; a Python-float (1 - beta) would not
        if grad.dtype == torch.float32:
            omb1 = float(np.float32(1) - np.float32(self.b1))
            omb2 = float(np.float32(1) - np.float32(self.b2))
        else:
            omb1, omb2 = 1 - self.b1, 1 - self.b2
        self.m = self.m + (grad - self.m) * omb1
        self.v = self.v + (grad * grad - self.v) * omb2
        var = var - lr_t * self.m / (torch.sqrt(self.v) + self.eps)
        self.b1p = np.float32(self.b1p * np.float32(self.b1))
        self.b2p = np.float32(self.b2p * np.float32(self.b2))
        return var
