"""
CPU oracle for the optimizer step that follows the convolution hot path  --  TEST INFRASTRUCTURE ONLY.

float64 NumPy restatement of numpy_ml/neural_nets/optimizers/optimizers.py (SGD :133-148, AdaGrad :243-259,
RMSProp :345-361, Adam :450-482) and of the Constant / Exponential / Noam learning-rate schedules
(schedulers/schedulers.py:60-69, 124-141, 192-195).  Parity status: PINNED -- tests/golden/optimizers.npz was produced
by the unmodified reference (tests/golden/make_golden.py) and tests/test_oracle_golden.py checks every function here
against it.  Only tests/ may import this module; the product never does.
"""
import numpy as np


def clip(grad, clip_norm):
    """Rescale to L2 norm `clip_norm` when longer (the same three lines in every reference update)."""
    if clip_norm is None:
        return grad
    n = np.linalg.norm(grad)
    return grad * clip_norm / n if n > clip_norm else grad


def lr_constant(lr, step):
    return lr


def lr_exponential(initial_lr, stage_length, staircase, decay, step):
    stage = step / stage_length
    if staircase:
        stage = np.floor(stage)
    return initial_lr * decay ** stage


def lr_noam(model_dim, scale_factor, warmup_steps, step):
    return scale_factor * model_dim ** (-0.5) * min(step ** (-0.5), step * warmup_steps ** (-1.5))


def sgd_step(param, grad, cache, lr, momentum, clip_norm=None):
    g = clip(grad, clip_norm)
    update = momentum * cache + lr * g
    return param - update, update


def adagrad_step(param, grad, cache, lr, eps, clip_norm=None):
    g = clip(grad, clip_norm)
    cache = cache + g ** 2
    return param - lr * g / (np.sqrt(cache) + eps), cache


def rmsprop_step(param, grad, cache, lr, decay, eps, clip_norm=None):
    g = clip(grad, clip_norm)
    cache = decay * cache + (1 - decay) * g ** 2
    return param - lr * g / (np.sqrt(cache) + eps), cache


def adam_step(param, grad, mean, var, t, lr, decay1, decay2, eps, clip_norm=None):
    """t is the step count BEFORE this update (the reference increments it first)."""
    g = clip(grad, clip_norm)
    t = t + 1
    var = decay2 * var + (1 - decay2) * g ** 2
    mean = decay1 * mean + (1 - decay1) * g
    v_hat = var / (1 - decay2 ** t)
    m_hat = mean / (1 - decay1 ** t)
    return param - lr * m_hat / (np.sqrt(v_hat) + eps), mean, var, t
