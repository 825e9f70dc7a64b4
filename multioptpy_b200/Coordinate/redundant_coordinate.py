"""Drop-in for the hot-path part of ``multioptpy.Coordinate.redundant_coordinate``
(Coordinate/redundant_coordinate.py:10-146,150-320,377-439) on the CUDA kernels of csrc/ric.cu.
Same function names and argument meaning; NumPy in / NumPy out for one structure, or CUDA tensors
with a leading batch dimension.  ``cartgrad2RICgrad`` (a solve with the singular B B^T)
is not offered: pass the internal gradient you want K built from."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from .._lib import MopError

_DEV = "cuda"


def _xyz(coord, dev=None):
    if isinstance(coord, torch.Tensor):
        return coord if coord.dim() == 3 else coord.reshape(1, -1, 3), True
    a = np.ascontiguousarray(np.asarray(coord, dtype=np.float64)).reshape(1, -1, 3)
    return torch.from_numpy(a).to(dev or _DEV), False


def _t(a, dev, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(np.asarray(a)), dtype=dtype).to(dev)


def _tables(connectivity, dev):
    tabs = []
    for t, w in zip(connectivity, (2, 3, 4)):
        a = np.asarray(t, dtype=np.int32).reshape(-1, w) if len(t) else np.zeros((0, w), np.int32)
        pad = np.full((max(len(a), 1), w), -1, np.int32); pad[: len(a)] = a
        tabs.append(torch.from_numpy(pad).to(dev))
    counts = torch.tensor([len(t) for t in connectivity], dtype=torch.int32, device=dev)
    return tabs, counts


class RedundantInternalCoordinates:
    def __init__(self, device=_DEV):
        self.device = torch.device(device)

    def B_matrix(self, coord):
        x, tensor = _xyz(coord, self.device)
        Bm = ops.ric_bmatrix(x.contiguous())
        return Bm if tensor else Bm[0].cpu().numpy()

    def G_matrix(self, b_mat):
        return b_mat @ b_mat.transpose(-1, -2) if isinstance(b_mat, torch.Tensor) else np.dot(b_mat, b_mat.T)

    def RICgrad2cartgrad(self, RICgrad, b_mat=None, coord=None):
        """B^T q.  The device kernel works from the geometry: pass ``coord`` (b_mat is accepted for
        signature compatibility and used only when no geometry is given)."""
        if coord is None:
            if b_mat is None:
                raise MopError("RICgrad2cartgrad: coord or b_mat required")
            return b_mat.transpose(-1, -2) @ RICgrad if isinstance(b_mat, torch.Tensor) else np.dot(b_mat.T, RICgrad)
        x, tensor = _xyz(coord, self.device)
        q = RICgrad if tensor else _t(np.asarray(RICgrad, float).reshape(1, -1), x.device)
        g = ops.ric_grad_to_cart(x.contiguous(), q.contiguous())
        return g if tensor else g[0].cpu().numpy()

    def cartgrad2RICgrad(self, cartgrad, b_mat):
        raise MopError("cartgrad2RICgrad solves the singular system (B B^T) q = B g (SURVEY H2): not offered on the device")

    def K_matrix(self, cart_coord, connectivity, RICgrad):
        x, tensor = _xyz(cart_coord, self.device)
        tabs, counts = _tables(connectivity, x.device)
        q = RICgrad if tensor else _t(np.asarray(RICgrad, float).reshape(1, -1), x.device)
        K = ops.ric_kmatrix(x.contiguous(), tabs[0], tabs[1], tabs[2], counts, q.contiguous())
        return K if tensor else K[0].cpu().numpy()

    def RIChess2carthess(self, cart_coord, connectivity, RIChess, b_mat, RICgrad):
        """B^T H B + K (redundant_coordinate.py:63-146); b_mat is implied by the geometry."""
        x, tensor = _xyz(cart_coord, self.device)
        tabs, counts = _tables(connectivity, x.device)
        if tensor:
            H, q = RIChess, RICgrad
        else:
            H = _t(np.asarray(RIChess, float)[None], x.device)
            q = _t(np.asarray(RICgrad, float).reshape(1, -1), x.device)
        K = ops.ric_kmatrix(x.contiguous(), tabs[0], tabs[1], tabs[2], counts, q.contiguous())
        out = ops.ric_hess_to_cart(x.contiguous(), H.contiguous(), K)
        return out if tensor else out[0].cpu().numpy()


def _partial(coord, labels):
    x, tensor = _xyz(coord)
    lab = torch.zeros(1, 4, dtype=torch.int32, device=x.device)
    lab[0, : len(labels)] = torch.tensor(labels, dtype=torch.int32)
    rows = ops.ric_partial_rows(x.contiguous(), lab)
    return rows[:, 0] if tensor else rows[0].cpu().numpy()


def partial_stretch_B_matirx(coord, atom_label_1, atom_label_2):
    return _partial(coord, [atom_label_1, atom_label_2])


def partial_bend_B_matrix(coord, atom_label_1, atom_label_2, atom_label_3):
    return _partial(coord, [atom_label_1, atom_label_2, atom_label_3])


def partial_torsion_B_matrix(coord, atom_label_1, atom_label_2, atom_label_3, atom_label_4):
    return _partial(coord, [atom_label_1, atom_label_2, atom_label_3, atom_label_4])


def calc_int_grad_from_pBmat(cart_grad, pBmat):
    if isinstance(pBmat, torch.Tensor):
        return ops.ric_pb_int_grad(pBmat, cart_grad)
    pB = _t(np.asarray(pBmat, float)[None], _DEV)
    g = _t(np.asarray(cart_grad, float).reshape(1, -1), _DEV)
    return ops.ric_pb_int_grad(pB, g)[0].cpu().numpy().reshape(-1, 1)


def calc_cart_grad_from_pBmat(int_grad, pBmat):
    if isinstance(pBmat, torch.Tensor):
        return ops.ric_pb_cart_grad(pBmat, int_grad)
    pB = _t(np.asarray(pBmat, float)[None], _DEV)
    q = _t(np.asarray(int_grad, float).reshape(1, -1), _DEV)
    return ops.ric_pb_cart_grad(pB, q)[0].cpu().numpy().reshape(-1, 1)
