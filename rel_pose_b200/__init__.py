"""rel_pose_b200 -- B200-native (sm_100a) implementation of the rel_pose hot path:
`ViTEss.forward` (CNN-tokenised ViT + Essential Matrix Module + pose regressor) and the lietorch SE3
group operations its loss needs, behind the reference's own Python interface.

    from rel_pose_b200 import ViTEss, SE3        # same surface as `src.model.ViTEss`, `lietorch.SE3`

All arithmetic lives in librelpose_b200.so (hand-written CUDA, C ABI in include/relpose_b200.h).
"""
from .lietorch import SE3, install_as_lietorch  # noqa: F401


def __getattr__(name):
    if name == "ViTEss":          # lazy: pulls in torchvision
        from .model import ViTEss
        return ViTEss
    raise AttributeError(name)


__all__ = ["ViTEss", "SE3", "install_as_lietorch"]
