import ast
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["l2_n384_3dmatch", "l12_n512_3dmatch", "l2_n300_kitti"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = ast.literal_eval(str(z["meta"]))
    fx = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}
    return meta, fx


def golden_cfg(meta):
    return dict(num_layers=meta["num_layers"], num_iterations=10, ratio=0.1, inlier_threshold=meta["thr"],
                sigma_d=meta["thr"], k=40, nms_radius=meta["thr"])


def golden_state_dict(meta):
    """Hot-path tensors only, regenerated from the weight seed (same generator as gen_golden.py)."""
    from gmf_b200.synth import synth_state_dict
    from gmf_b200.weights import hot_path_spec

    sd = synth_state_dict(hot_path_spec(meta["num_layers"]), seed=meta["wseed"], plain_init=meta["plain"])
    sd["sigma_spat"] = torch.tensor([meta["thr"]], dtype=torch.float32)
    return sd


@pytest.fixture(scope="session")
def has_cuda():
    return torch.cuda.is_available()


def record(name, **metrics):
    """Append measured parity numbers to gpurun_out/parity_measured.jsonl (merged back from the GPU box; the round's copy is committed
    under profiles/).  Never fails a test."""
    import json
    try:
        d = os.path.join(ROOT, "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_measured.jsonl"), "a") as f:
            f.write(json.dumps({"test": name, **{k: (float(v) if isinstance(v, (int, float)) or hasattr(v, "item") else v) for k, v in metrics.items()}}) + "\n")
    except Exception:
        pass
