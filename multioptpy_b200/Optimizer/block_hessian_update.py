"""Module alias so ``from multioptpy_b200.Optimizer.block_hessian_update import
BlockHessianUpdate`` mirrors the reference layout."""
from .hessian_update import BlockHessianUpdate  # noqa: F401
