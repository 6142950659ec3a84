"""TEST INFRASTRUCTURE ONLY — NumPy restatement of the reference's correspondence construction
(GMF_PointDSC/datasets/ThreeDMatch.py:384-391 matcher, :401-402 gather, :411-414 corr_pos for in_dim == 6; identical code in
datasets/KITTI.py:94-102).  Only tests/ and benchmark CPU legs may import it.

Pinned: oracle/gen_golden_matcher.py executes the reference's own source lines (read from /root/reference at generation time,
not copied) on seeded descriptors and commits the outputs as tests/golden/matcher_*.npz.
"""
from __future__ import annotations

import numpy as np


def build_correspondences(src_desc, tgt_desc, src_keypts, tgt_keypts, use_mutual=False):
    src_desc, tgt_desc = np.asarray(src_desc, np.float32), np.asarray(tgt_desc, np.float32)
    distance = np.sqrt(2 - 2 * (src_desc @ tgt_desc.T) + 1e-6)                  # :384
    source_idx = np.argmin(distance, axis=1)                                     # :385
    if use_mutual:                                                               # :386-389
        target_idx = np.argmin(distance, axis=0)
        mutual_nearest = target_idx[source_idx] == np.arange(source_idx.shape[0])
        corr = np.concatenate([np.where(mutual_nearest == 1)[0][:, None], source_idx[mutual_nearest][:, None]], axis=-1)
    else:                                                                        # :391
        corr = np.concatenate([np.arange(source_idx.shape[0])[:, None], source_idx[:, None]], axis=-1)
    input_src_keypts = src_keypts[corr[:, 0]]                                    # :401-402
    input_tgt_keypts = tgt_keypts[corr[:, 1]]
    corr_pos = np.concatenate([input_src_keypts, input_tgt_keypts], axis=-1)     # :411-413
    corr_pos = corr_pos - corr_pos.mean(0)
    return dict(distance=distance, source_idx=source_idx, corr=corr, src_keypts=input_src_keypts, tgt_keypts=input_tgt_keypts,
                corr_pos=corr_pos)


def synth_descriptors(ns, nt, d, seed, overlap=0.5, noise=0.15):
    """Unit-norm descriptors: a fraction `overlap` of the target rows are noisy copies of source rows (true matches)."""
    rng = np.random.default_rng(seed)
    s = rng.standard_normal((ns, d)).astype(np.float32)
    t = rng.standard_normal((nt, d)).astype(np.float32)
    m = int(min(ns, nt) * overlap)
    perm = rng.permutation(nt)[:m]
    t[perm] = s[rng.permutation(ns)[:m]] + noise * rng.standard_normal((m, d)).astype(np.float32)
    s /= np.linalg.norm(s, axis=1, keepdims=True)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    sk = (rng.random((ns, 3)) * 3.0).astype(np.float32)
    tk = (rng.random((nt, 3)) * 3.0).astype(np.float32)
    return s, t, sk, tk
