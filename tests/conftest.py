import ast
import os
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"), allow_pickle=False)
    case = {k: g[k] for k in g.files}
    case["decoder_params"] = ast.literal_eval(str(case["decoder_params"]))
    case["solver"] = str(case["solver"])
    case["lengths"] = [int(x) for x in case["lengths"]]
    for k in ("T", "n_steps", "input_seed", "weight_seed", "n_params"):
        case[k] = int(case[k])
    return case


GOLDEN_CASES = ["tiny_euler", "tiny_midpoint", "tiny_rk4_padded", "tiny_heun3", "prod_euler", "default_euler"]


def cfm_params(solver="euler"):
    return types.SimpleNamespace(solver=solver, sigma_min=1e-4, use_mu_prior=True)


def rel_l2(a, b):
    a, b = torch.as_tensor(a).detach().cpu().double(), torch.as_tensor(b).detach().cpu().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
