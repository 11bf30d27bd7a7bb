"""Loader for tests/golden/*.npz (written by tests/golden/make_golden.py from the reference classes)."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names(prefix=""):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    case = {}
    params, grads = {}, {}
    for k in z.files:
        v = z[k]
        if k.startswith("param:"):
            params[k[6:]] = torch.from_numpy(v.copy())
        elif k.startswith("grad:"):
            grads[k[5:]] = torch.from_numpy(v.copy())
        elif v.dtype.kind in "US":
            case[k] = str(v)
        elif v.ndim == 0:
            case[k] = v.item()
        else:
            case[k] = torch.from_numpy(v.copy())
    case["params"], case["grads"] = params, grads
    return case


def split_layers(d):
    """'0.in_weight' -> [ {in_weight:..}, ... ] for the 3-layer rep cases."""
    n = 1 + max(int(k.split(".")[0]) for k in d)
    return [{k.split(".", 1)[1]: v for k, v in d.items() if k.startswith("%d." % i)} for i in range(n)]
