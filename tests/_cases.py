"""Shared helpers: load a golden file and re-derive its inputs from the stored seeds."""
import os

import numpy as np
import torch

from genpose_b200 import synth
from oracle import make_golden

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(make_golden.CASES)


def load(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    case = make_golden.CASES[name]
    for k, v in case.items():                      # the file must describe the same case
        assert np.array_equal(np.asarray(v), g[f"case_{k}"]), (name, k)
    sd, esd, clouds, x0, step_noise = make_golden.case_inputs(case)
    cs = make_golden.input_checksums(sd, esd, clouds, x0, step_noise)
    for k, v in cs.items():                        # and the regenerated inputs must be the ones it was made from
        assert abs(v - float(g[k])) <= 1e-9 * max(1.0, abs(v)), f"{name}: input drift in {k}"
    return case, g, dict(sd=sd, esd=esd, clouds=clouds, x0=x0, step_noise=step_noise)


def t(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a))
