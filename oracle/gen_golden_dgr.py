"""TEST INFRASTRUCTURE ONLY — golden vectors of the DGR bottleneck head from the UNMODIFIED reference module.

Run in the build container (needs /root/reference):   python oracle/gen_golden_dgr.py
`model/__init__.py` of the reference imports MinkowskiEngine (absent), so perceiver_io.py is loaded directly by path
(SURVEY.md §8c).  Weights / inputs are seed-generated (gmf_b200.synth, oracle.dgr_head_oracle.synth_latents); the fixture stores
the seeds and the reference's output only.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gmf_b200.dgr_head import dgr_head_shapes       # noqa: E402
from gmf_b200.synth import synth_state_dict, synth_tokens   # noqa: E402
from oracle.dgr_head_oracle import synth_latents    # noqa: E402

REF = "/root/reference/GMF_DeepGlobalRegistration/GMF_DeepGlobalRegistration_fcgf/model/perceiver_io.py"
CASES = {"dgr_head_m200_t300": dict(m=200, t=300, wseed=5, dseed=21, pe=True),
         "dgr_head_m130_t257_nope": dict(m=130, t=257, wseed=6, dseed=22, pe=False)}


def load_reference_class():
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location("ref_perceiver_io", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.PerceiverIO


def main():
    PerceiverIO = load_reference_class()
    for name, c in CASES.items():
        sd = synth_state_dict(dgr_head_shapes(c["pe"]), seed=c["wseed"])
        m = PerceiverIO(dim=128, depth=0, latent_dim=256, cross_heads=1, latent_heads=8, cross_dim_head=128, latent_dim_head=128,
                        pe=c["pe"]).eval()
        res = m.load_state_dict(sd, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        x = synth_latents(c["m"], c["dseed"])
        ctx = synth_tokens(1, c["t"], c["dseed"])[0]
        with torch.no_grad():
            out = m(ctx.unsqueeze(0), queries_encoder=x.unsqueeze(0))[0]
            out64 = m.double()(ctx.double().unsqueeze(0), queries_encoder=x.double().unsqueeze(0))[0]
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, m=c["m"], t=c["t"], wseed=c["wseed"], dseed=c["dseed"], pe=int(c["pe"]), out=out.numpy(),
                            out64_absmax=float(out64.abs().max()), fp32_vs_fp64=float((out.double() - out64).abs().max()))
        print(name, tuple(out.shape), "fp32 vs fp64", float((out.double() - out64).abs().max()), os.path.getsize(path))


if __name__ == "__main__":
    main()
