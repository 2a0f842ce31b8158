"""cfg4's conv1 forward alone (for an ncu capture): 256 x 32 x 32 x 64 -> 128, 3x3, one output tensor."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("numpy-ml_b200")
rng = np.random.default_rng(0)
X = torch.from_numpy(rng.standard_normal((256, 32, 32, 64), dtype=np.float32)).cuda().to(torch.bfloat16)
L = pkg.Conv2D(out_ch=128, kernel_shape=(3, 3), pad=1, act_fn=None, mode="bf16")
for i in range(6):
    L.forward(X, retain_derived=False)
torch.cuda.synchronize()
