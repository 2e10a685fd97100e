"""Shared helpers of the parity tests."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import oracle.curvature_oracle as orc  # noqa: E402  (the checker; tests may import it)


def rel_fro(a, b):
    """Relative Frobenius error of a against b (b = the reference)."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)


def conv_zoo():
    """Same architecture as tests/golden/make_golden.py:conv_zoo (weights come from the fixture)."""
    return torch.nn.Sequential(
        torch.nn.Conv2d(3, 5, (3, 2), stride=(2, 1), padding=(1, 0), bias=True),
        torch.nn.Tanh(),
        torch.nn.Conv2d(5, 4, 3, stride=1, padding=1, bias=False),
        torch.nn.Tanh(),
        torch.nn.Conv2d(4, 6, (1, 3), stride=(1, 2), padding=(0, 2), bias=True),
        torch.nn.Tanh(),
        torch.nn.Conv2d(6, 7, 1, stride=2, padding=0, bias=False),
        torch.nn.Flatten(),
        torch.nn.Linear(7 * 3 * 3, 9, bias=False),
        torch.nn.Tanh(),
        torch.nn.Linear(9, 4, bias=True))


MODELS = {"lenet5": orc.lenet5, "convzoo": conv_zoo}


def model_from_golden(name, g, device="cpu"):
    model = MODELS[name]()
    sd = {k[len("param/"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("param/")}
    model.load_state_dict(sd)
    return model.to(device)


def selected_layers(model):
    return [m for m in model.modules() if m.__class__.__name__ in ("Linear", "Conv2d")]


def n_batches(g):
    return len([k for k in g.files if k.startswith("x/")])
