"""Correspondence construction on the GPU — the NumPy block of the reference's datasets
(GMF_PointDSC/datasets/ThreeDMatch.py:384-391, 401-402, 411-414; datasets/KITTI.py:94-102) behind the C ABI call
`gmf_build_correspondences`.  No CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import _lib


def build_correspondences(engine, src_desc: torch.Tensor, tgt_desc: torch.Tensor, src_keypts: torch.Tensor, tgt_keypts: torch.Tensor,
                          use_mutual: bool = False) -> Dict[str, torch.Tensor]:
    """src_desc [B,Ns,D], tgt_desc [B,Nt,D], src_keypts [B,Ns,3], tgt_keypts [B,Nt,3] (cuda fp32; 2-D inputs are one pair).
    Returns source_idx [B,Ns] int32, corr [B,Ns,2] int32, n_corr [B] int32, src_keypts / tgt_keypts [B,Ns,3] (gathered) and
    corr_pos [B,Ns,6]; rows at and past n_corr[b] are zero-filled (only when use_mutual drops rows)."""
    single = src_desc.dim() == 2
    if single:
        src_desc, tgt_desc, src_keypts, tgt_keypts = (t.unsqueeze(0) for t in (src_desc, tgt_desc, src_keypts, tgt_keypts))
    for t in (src_desc, tgt_desc, src_keypts, tgt_keypts):
        if not (t.is_cuda and t.dtype == torch.float32):
            raise _lib.GmfError("build_correspondences needs CUDA fp32 tensors (gmf_b200 has no CPU fallback)")
    B, Ns, D = src_desc.shape
    Nt = tgt_desc.shape[1]
    dev = src_desc.device
    lib = _lib.load()
    src_desc, tgt_desc, src_keypts, tgt_keypts = (t.contiguous() for t in (src_desc, tgt_desc, src_keypts, tgt_keypts))
    source_idx = torch.empty(B, Ns, dtype=torch.int32, device=dev)
    corr = torch.empty(B, Ns, 2, dtype=torch.int32, device=dev)
    n_corr = torch.empty(B, dtype=torch.int32, device=dev)
    src_sel = torch.empty(B, Ns, 3, device=dev)
    tgt_sel = torch.empty(B, Ns, 3, device=dev)
    corr_pos = torch.empty(B, Ns, 6, device=dev)
    ws_bytes = int(lib.gmf_match_workspace_bytes(B, Ns, Nt, D))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(lib.gmf_build_correspondences(engine.h, src_desc.data_ptr(), tgt_desc.data_ptr(), src_keypts.data_ptr(), tgt_keypts.data_ptr(),
                                             B, Ns, Nt, D, int(use_mutual), source_idx.data_ptr(), corr.data_ptr(), n_corr.data_ptr(),
                                             src_sel.data_ptr(), tgt_sel.data_ptr(), corr_pos.data_ptr(), ws.data_ptr(), ws_bytes, C.c_void_p(st)))
    out = {"source_idx": source_idx, "corr": corr, "n_corr": n_corr, "src_keypts": src_sel, "tgt_keypts": tgt_sel, "corr_pos": corr_pos}
    if single:
        out = {k: v[0] for k, v in out.items()}
    return out
