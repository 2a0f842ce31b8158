"""numpy-ml_b200: the numpy-ml convolution hot path (conv2D / im2col / col2im / pad2D, Conv2D,
Deconv2D, BatchNorm2D, Add, SkipConnectionConvModule, and the callers next to them: Conv1D, Pool2D, Flatten,
SkipConnectionIdentityModule) on B200, behind the reference's layer API.

The directory name is not a Python identifier; import it with
    import importlib; npml = importlib.import_module("numpy-ml_b200")
"""
from . import _lib
from . import utils
from . import activations
from . import schedulers
from . import optimizers
from . import layers
from . import modules
from . import dist
from . import pipeline
from ._lib import to_numpy, launch_count
from .build import build
from .layers import Conv2D, Deconv2D, Conv1D, Pool2D, Flatten, BatchNorm2D, Add, Multiply, FullyConnected
from .modules import SkipConnectionConvModule, SkipConnectionIdentityModule, WavenetResidualModule
from .utils import (pad2D, calc_pad_dims_2D, calc_conv_out_dims, im2col, col2im, conv2D, conv2D_naive,
                    deconv2D_naive, dilate, pad1D, calc_pad_dims_1D, conv1D)
