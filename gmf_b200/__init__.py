"""gmf_b200 — B200-native (sm_100a) GMF-PointDSC correspondence outlier-rejection forward path."""
from .weights import hot_path_spec, pack_state_dict  # noqa: F401


def __getattr__(name):
    if name in ("PointDSC",):
        from .module import PointDSC
        return PointDSC
    if name in ("Engine",):
        from .engine import Engine
        return Engine
    if name in ("PerceiverIO", "DgrHeadEngine"):
        from . import dgr_head
        return getattr(dgr_head, name)
    raise AttributeError(name)
