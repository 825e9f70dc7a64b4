"""Config 1 (aldol_rxn, RFO + Bofill, Fischer model Hessian, AFIR bias): the
calc_move_vector boundary replayed on the golden trace recorded from the reference."""
import os

import numpy as np
import pytest

from oracle import np_oracle as O

RTOL = 1e-10


def rel(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0)


def test_oracle_replays_reference_trace(golden_dir):
    z = np.load(os.path.join(golden_dir, "c1_calc_move_vector.npz"))
    cmv = O.CalcMoveVectorOracle("rsirfo_bofill", 0, model_hess_flag="fischer")
    cmv.opt.set_hessian(z["H0"].copy())
    pre = None
    for k in range(z["geom"].shape[0]):
        x, Bg, g, Be = z["geom"][k].ravel(), z["B_g"][k].ravel(), z["g"][k].ravel(), float(z["B_e"][k])
        cmv.opt.set_bias_hessian(z["Hb"][k])
        xn, mv = cmv.step(x, Bg, g, Be, pre)
        assert rel(mv, z["move"][k].ravel()) < RTOL, k
        assert rel(xn, z["new_geom"][k].ravel()) < RTOL, k
        assert abs(cmv.trust - z["trust"][k]) < 1e-12, k
        assert rel(cmv.opt.hessian, z["H_after"][k]) < RTOL, k
        pre = dict(x=x, g=g, Bg=Bg, Be=Be, move=mv)


@pytest.mark.gpu
def test_gpu_dropin_replays_reference_trace(golden_dir):
    """The drop-in classes with the reference calling convention (NumPy arrays in, NumPy out)."""
    from multioptpy_b200.optimizer import CalculateMoveVector
    z = np.load(os.path.join(golden_dir, "c1_calc_move_vector.npz"))
    elems = [str(e) for e in z["elements"]]
    N = len(elems)
    cmv = CalculateMoveVector(0.5, elems, saddle_order=0, FC_COUNT=-1, model_hess_flag="fischer", device="cuda:0")
    insts = cmv.initialization(["rsirfo_bofill"])
    Model_hess = z["H0"].copy()
    pre = dict(B_g=np.zeros((N, 3)), geom=np.zeros((N, 3)), B_e=0.0, move=np.zeros((N, 3)), g=np.zeros((N, 3)))
    for k in range(z["geom"].shape[0]):
        insts[0].set_hessian(Model_hess)
        insts[0].set_bias_hessian(z["Hb"][k].copy())
        new_geom, move, insts = cmv.calc_move_vector(k, z["geom"][k].copy(), z["B_g"][k].copy(), pre["B_g"], pre["geom"],
                                                     float(z["B_e"][k]), pre["B_e"], pre["move"], z["xyz0"].copy(),
                                                     z["g"][k].copy(), pre["g"], insts, print_flag=False)
        assert new_geom.shape == (N, 3) and move.shape == (N, 3)
        assert rel(move, z["move"][k]) < RTOL, k
        assert rel(new_geom, z["new_geom"][k]) < RTOL, k
        assert abs(cmv.trust_radii - z["trust"][k]) < 1e-12, k
        assert rel(Model_hess, z["H_after"][k]) < RTOL, k
        pre = dict(B_g=z["B_g"][k].copy(), geom=z["geom"][k].copy(), B_e=float(z["B_e"][k]), move=move, g=z["g"][k].copy())


@pytest.mark.gpu
def test_gpu_full_chain_model_hessian_bias_step(golden_dir):
    """Whole config-1 chain on the device: Fischer model Hessian + AFIR bias + RFO/Bofill step."""
    from multioptpy_b200.optimizer import CalculateMoveVector
    from multioptpy_b200.ModelHessian.approx_hessian import ApproxHessian
    from multioptpy_b200.Potential.potential import BiasPotentialCalculation
    z = np.load(os.path.join(golden_dir, "c1_calc_move_vector.npz"))
    elems = [str(e) for e in z["elements"]]
    N = len(elems)
    fd = {"AFIR_gamma": [[95.0], [50.0]], "AFIR_Fragm_1": [[1], [3]], "AFIR_Fragm_2": [[5], [11]]}
    H0 = ApproxHessian(device="cuda:0").main(z["xyz0"], elems, np.zeros((N, 3)), "fischer")
    assert rel(H0, z["H0"]) < RTOL
    cmv = CalculateMoveVector(0.5, elems, saddle_order=0, FC_COUNT=-1, model_hess_flag="fischer", device="cuda:0")
    insts = cmv.initialization(["rsirfo_bofill"])
    insts[0].set_hessian(H0)
    bpc = BiasPotentialCalculation(device="cuda:0")
    bg, Be, Bg, Hb = bpc.main(0.0, z["g"][0], z["geom"][0], elems, fd)
    assert rel(Hb, z["Hb"][0]) < RTOL and rel(Bg, z["B_g"][0]) < RTOL
    insts[0].set_bias_hessian(Hb)
    zero = np.zeros((N, 3))
    new_geom, move, _ = cmv.calc_move_vector(0, z["geom"][0].copy(), Bg, zero, zero, float(z["B_e"][0]), 0.0, zero,
                                             z["xyz0"].copy(), z["g"][0].copy(), zero, insts, print_flag=False)
    assert rel(move, z["move"][0]) < RTOL
    assert rel(new_geom, z["new_geom"][0]) < RTOL
