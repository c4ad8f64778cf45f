"""Helpers for the -m gpu parity tests: error recording (so a GPU run leaves the measured deviations behind)."""
import json
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LOG = os.path.join(ROOT, "gpurun_out", "parity_errors.jsonl")


def record(case: str, **vals):
    try:
        os.makedirs(os.path.dirname(_LOG), exist_ok=True)
        with open(_LOG, "a") as f:
            f.write(json.dumps({"case": case, **{k: float(v) for k, v in vals.items()}}) + "\n")
    except OSError:
        pass


def maxabs(a, b) -> float:
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max()) if a.numel() else 0.0
