"""TEST INFRASTRUCTURE (oracle): NumPy restatement of the reference's trajectory optimiser step — Adam on the composite action table
(fluidlab/optimizer/optim.py:22-41) as driven by TrainablePolicy.optimize (fluidlab/optimizer/policies.py:152-164).  Only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import this; the product's step is the CUDA kernel k_adam_step (csrc/fmpm_io.cu).

Pinned: tests/golden/reference_optim.npz holds tables produced by the reference's own Adam / TrainablePolicy classes
(tests/golden/make_reference_optim.py); tests/test_optimizer.py checks this restatement against them bit for bit."""
import numpy as np


class AdamOracle:
    """optim.py:22-41.  The dtypes are spelled out: grads is float32 (agent.get_grad), so `(1 - beta) * grads` and `grads * grads` are float32
    products; the moment buffers and the parameters are float64."""

    def __init__(self, shape, lr, beta_1, beta_2, epsilon):
        self.lr, self.beta_1, self.beta_2, self.epsilon = float(lr), float(beta_1), float(beta_2), float(epsilon)
        self.m = np.zeros(shape, np.float64)       # optim.py:24
        self.v = np.zeros(shape, np.float64)       # optim.py:25
        self.iter = 0

    def step(self, params, grads):
        g = np.asarray(grads, np.float32)
        t1 = (np.float32(1.0 - self.beta_1) * g).astype(np.float64)            # optim.py:32, float32 product
        t2 = (np.float32(1.0 - self.beta_2) * (g * g)).astype(np.float64)      # optim.py:33
        self.m = self.beta_1 * self.m + t1
        self.v = self.beta_2 * self.v + t2
        m_cap = self.m / (1 - self.beta_1 ** (self.iter + 1))                  # optim.py:37
        v_cap = self.v / (1 - self.beta_2 ** (self.iter + 1))                  # optim.py:38
        self.iter += 1
        return np.asarray(params, np.float64) - (self.lr * m_cap) / (np.sqrt(v_cap) + self.epsilon)   # optim.py:41


def policy_optimize(adam, actions_v, actions_p, grads, action_range, trainable=None, fix_dim=None):
    """TrainablePolicy.optimize, policies.py:152-164 -> (actions_v, actions_p)"""
    g = np.array(grads, np.float32)
    if trainable is not None:
        g[np.logical_not(trainable)] = 0           # policies.py:155
    if fix_dim is not None:
        g[:, fix_dim] = 0                          # policies.py:157
    new = adam.step(np.vstack([actions_v, actions_p[None, :]]), g)             # policies.py:147-148, 159
    return new[:-1].clip(*action_range), new[-1]   # policies.py:160-161
