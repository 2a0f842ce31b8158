"""Learning-rate schedules with the reference's names, string forms and `hyperparameters`
(numpy_ml/neural_nets/schedulers/schedulers.py).  They are scalar host arithmetic: the value they return is passed to
the optimizer kernels as a plain float.  Constant (:43-69), Exponential (:72-141) and Noam (:144-195) are provided;
KingScheduler (:198-362, a regression over the recent loss history) is not."""
import math
import re
from copy import deepcopy


class SchedulerBase(object):
    def __init__(self):
        self.hyperparameters = {}

    def __call__(self, step=None, cur_loss=None):
        return self.learning_rate(step=step, cur_loss=cur_loss)

    def copy(self):
        return deepcopy(self)

    def set_params(self, hparam_dict):
        if hparam_dict is not None:
            for k, v in hparam_dict.items():
                if k in self.hyperparameters:
                    self.hyperparameters[k] = v
                    if hasattr(self, k):
                        setattr(self, k, v)


class ConstantScheduler(SchedulerBase):
    def __init__(self, lr=0.01, **kwargs):
        super().__init__()
        self.lr = lr
        self.hyperparameters = {"id": "ConstantScheduler", "lr": self.lr}

    def __str__(self):
        return "ConstantScheduler(lr={})".format(self.lr)

    def learning_rate(self, **kwargs):
        return self.lr


class ExponentialScheduler(SchedulerBase):
    """lr = initial_lr * decay ** (step / stage_length), the exponent floored when `staircase`."""

    def __init__(self, initial_lr=0.01, stage_length=500, staircase=False, decay=0.1, **kwargs):
        super().__init__()
        self.decay, self.staircase, self.initial_lr, self.stage_length = decay, staircase, initial_lr, stage_length
        self.hyperparameters = {"id": "StepScheduler", "decay": decay, "staircase": staircase,
                                "initial_lr": initial_lr, "stage_length": stage_length}

    def __str__(self):
        return "ExponentialScheduler(initial_lr={}, stage_length={}, staircase={}, decay={})".format(
            self.initial_lr, self.stage_length, self.staircase, self.decay)

    def learning_rate(self, step, **kwargs):
        stage = step / self.stage_length
        if self.staircase:
            stage = math.floor(stage)
        return self.initial_lr * self.decay ** stage


class NoamScheduler(SchedulerBase):
    """lr = scale_factor * model_dim**-0.5 * min(step**-0.5, step * warmup_steps**-1.5)."""

    def __init__(self, model_dim=512, scale_factor=1, warmup_steps=4000, **kwargs):
        super().__init__()
        self.model_dim, self.scale_factor, self.warmup_steps = model_dim, scale_factor, warmup_steps
        self.hyperparameters = {"id": "NoamScheduler", "model_dim": model_dim, "scale_factor": scale_factor,
                                "warmup_steps": warmup_steps}

    def __str__(self):
        return "NoamScheduler(model_dim={}, scale_factor={}, warmup_steps={})".format(
            self.model_dim, self.scale_factor, self.warmup_steps)

    def learning_rate(self, step, **kwargs):
        return self.scale_factor * self.model_dim ** (-0.5) * min(step ** (-0.5), step * self.warmup_steps ** (-1.5))


def _eval(v):
    s = str(v).strip()
    if s.lower() in ("none",):
        return None
    if s.lower() in ("true", "false"):
        return s.lower() == "true"
    try:
        f = float(s)
        return int(f) if f.is_integer() and re.fullmatch(r"[-+]?\d+", s) else f
    except ValueError:
        return s


class SchedulerInitializer(object):
    """None (constant at `lr`) | SchedulerBase | its str() form | {"hyperparameters": ...}
    (initializers/initializers.py:106-173)."""

    def __init__(self, param=None, lr=None):
        if lr is None and param is None:
            raise ValueError("lr and param cannot both be `None`")
        self.lr, self.param = lr, param

    def __call__(self):
        p = self.param
        if p is None:
            return ConstantScheduler(self.lr)
        if isinstance(p, SchedulerBase):
            return p
        if isinstance(p, str):
            s = p.lower()
            kwargs = {k: _eval(v) for k, v in re.findall(r"([a-zA-Z_]*)=([^,)]*)", s)}
            if "constant" in s:
                return ConstantScheduler(**kwargs)
            if "exponential" in s:
                return ExponentialScheduler(**kwargs)
            if "noam" in s:
                return NoamScheduler(**kwargs)
            raise NotImplementedError("{}".format(s))
        if isinstance(p, dict):
            sc = p.get("hyperparameters")
            if sc is None:
                raise ValueError("Must have `hyperparameters` key: {}".format(p))
            kinds = {"ConstantScheduler": ConstantScheduler, "ExponentialScheduler": ExponentialScheduler,
                     "StepScheduler": ExponentialScheduler, "NoamScheduler": NoamScheduler}
            if sc["id"] not in kinds:
                raise NotImplementedError("{}".format(sc["id"]))
            sched = kinds[sc["id"]]()
            sched.set_params(sc)
            return sched
        raise ValueError("Unknown scheduler: {}".format(p))
