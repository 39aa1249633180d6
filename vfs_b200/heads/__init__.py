from .sim_siam_head import SimSiamHead, build_norm1d

__all__ = ['SimSiamHead', 'build_norm1d']
