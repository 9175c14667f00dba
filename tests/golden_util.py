"""Load golden fixtures (tests/golden/*.npz) and rebuild their inputs."""
import ast
import os

import numpy as np
import torch

from cross_attention_renderer_b200 import synthetic

GOLDEN_DIR = os.path.join(os.path.dirname(__file__), "golden")
CASES = ("tiny_default", "tiny_peaky", "tiny_mixed", "c1_sparse", "p64_sparse")


def load_case(name):
    rec = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    cfg = ast.literal_eval(str(rec.pop("cfg")))
    inp = synthetic.make_inputs(cfg["b"], cfg["H"], cfg["Ht"], seed=cfg["seed"],
                                mode=cfg["mode"], rays=cfg["rays"])
    z = synthetic.make_features(cfg["b"], cfg["H"], seed=cfg["seed"])
    sd = synthetic.make_state_dict(seed=cfg["seed"], peaky=cfg["peaky"])
    rec = {k: torch.from_numpy(v) for k, v in rec.items()}
    return cfg, inp, z, sd, rec


def ulp_diff(a, b):
    """Distance in units of last place between two fp32 tensors (finite values)."""
    ai = a.contiguous().view(torch.int32).long()
    bi = b.contiguous().view(torch.int32).long()
    ai = torch.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = torch.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return (ai - bi).abs()


def rel_err(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
