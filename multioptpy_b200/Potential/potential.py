"""Drop-in for ``multioptpy.Potential.potential.BiasPotentialCalculation`` (Potential/potential.py:53-202)
restricted to the potentials built on the device: sums the bias energy / gradient / Hessian of every AFIR
term, of the keep (distance, fragment distance, angle, dihedral, out-of-plane angle, anharmonic) restraints, of the
well potentials (fragment - fragment, atom - wall, atom - fixed point, atom - centre fragment) and of the LJ repulsive potential (scale / value units) of ``force_data``; every other
potential the reference would activate raises ``MopError`` (``active_keys`` walks the reference's key list).
The reference's side effects (.npy / .log files, :144,191-192) are not reproduced."""
from __future__ import annotations

import numpy as np
import torch

from .._lib import MopError
from .. import ops
from .AFIR_potential import AFIRPotential

# Every activation key of the reference aggregator (potential.py:228-300,434-900) with the reference's OWN activation
# test: "nz" = an entry != 0.0, "all" = an entry (a list) without a 0.0 in it, "len" = a non-empty list.  Keys handled
# on the device are in _HANDLED; any other ACTIVE key raises instead of being silently dropped.
_ACTIVATION = {
    "AFIR_gamma": "all", "keep_pot_spring_const": "nz", "keep_pot_v2_spring_const": "all",
    "keep_angle_spring_const": "nz", "keep_dihedral_angle_spring_const": "nz",
    "anharmonic_keep_pot_spring_const": "nz", "well_pot_wall_energy": "nz", "wall_well_pot_wall_energy": "nz",
    "void_point_well_pot_wall_energy": "nz", "around_well_pot_wall_energy": "nz",
    "keep_angle_v2_spring_const": "all", "keep_out_of_plain_angle_spring_const": "nz",
    "keep_dihedral_angle_v2_spring_const": "all", "keep_dihedral_angle_cos_potential_const": "all",
    "keep_out_of_plain_angle_v2_spring_const": "all", "void_point_pot_spring_const": "nz",
    "linear_mechano_force": "nz", "linear_mechano_force_v2": "nz", "flux_pot_const": "len",
    "value_range_upper_const": "nz", "universal_pot_const": "nz", "repulsive_potential_v2_well_scale": "nz",
    "repulsive_potential_well_scale": "nz", "repulsive_potential_gaussian_LJ_well_depth": "nz",
    "repulsive_potential_gaussian_gau_well_depth": "nz", "cone_potential_well_value": "nz",
    "spacer_model_potential_well_depth": "nz", "gaussian_potential_target": "len", "nano_reactor_potential": "len",
    "asymmetric_ellipsoidal_repulsive_potential_eps": "len", "asymmetric_ellipsoidal_repulsive_potential_v2_eps": "len",
}
_HANDLED = {"AFIR_gamma", "keep_pot_spring_const", "keep_pot_v2_spring_const", "keep_angle_spring_const",
            "keep_dihedral_angle_spring_const", "anharmonic_keep_pot_spring_const", "well_pot_wall_energy",
            "keep_out_of_plain_angle_spring_const", "repulsive_potential_well_scale", "keep_angle_v2_spring_const",
            "keep_dihedral_angle_v2_spring_const", "keep_out_of_plain_angle_v2_spring_const",
            "wall_well_pot_wall_energy", "void_point_well_pot_wall_energy", "around_well_pot_wall_energy"}


def centroid_term(kind, fragments, k, p):
    """One fragment-centroid restraint (kinds 9 - 11): all atoms in `atoms`, the fragment sizes in q."""
    atoms = [a - 1 for f in fragments for a in f]
    return (kind, atoms, [], float(k), float(p), [float(len(f)) for f in fragments])


def lj_pair_terms(element_list, fragm_1, fragm_2, well, dist, unit):
    """LJRepulsivePotentialScale / Value (LJ_repulsive_potential.py:9-114) expanded into one term per atom pair of
    the fragment product (torch.meshgrid(..., indexing='ij') order).  "scale": eps, sigma = sqrt(scale^2 p_i p_j) in the
    reference's FLOAT32 arithmetic (its UFF tables go through torch.tensor(list of Python floats)); "value": kJ/mol and
    Angstrom converted in float64."""
    from ..Parameters import tables
    terms = []
    for a in fragm_1:
        for b in fragm_2:
            if unit == "scale":
                ea, eb = element_list[a - 1], element_list[b - 1]
                w = torch.sqrt(well ** 2 * torch.tensor([tables.UFF_VDW_WELL_DEPTH[ea]]) * torch.tensor([tables.UFF_VDW_WELL_DEPTH[eb]]))
                d = torch.sqrt(dist ** 2 * torch.tensor([tables.UFF_VDW_DISTANCE[ea]]) * torch.tensor([tables.UFF_VDW_DISTANCE[eb]]))
                eps, sig = float(w[0]), float(d[0])
            elif unit == "value":
                eps, sig = well / tables.HARTREE2KJMOL, dist / tables.BOHR2ANG
            else:
                raise MopError("repulsive_potential_unit must be 'scale' or 'value'")
            terms.append((ops.BIAS_LJ_PAIR, [a - 1], [b - 1], eps, sig))
    return terms


def active_keys(force_data):
    """Activation keys of ``force_data`` that the reference aggregator would turn into a potential term."""
    out = []
    for key, rule in _ACTIVATION.items():
        val = force_data.get(key, [])
        if val is None or len(val) == 0:
            continue
        if rule == "len":
            on = True
        elif rule == "nz":
            on = any(v != 0.0 for v in val)
        else:
            on = any(0.0 not in list(np.atleast_1d(np.asarray(v, dtype=object))) for v in val)
        if on:
            out.append(key)
    return out


def gradually_change_param(param_1, param_2, iter):
    """potential.py:218-226: linear ramp over 300 iterations."""
    parameter = param_1 + ((param_2 - param_1) / 300) * int(iter)
    if param_1 < param_2:
        return min(parameter, param_2)
    if param_1 > param_2:
        return max(parameter, param_2)
    return parameter


class BiasPotentialCalculation:
    def __init__(self, FOLDER_DIRECTORY="", device="cuda"):
        self.BPA_FOLDER_DIRECTORY = FOLDER_DIRECTORY
        self.device = device
        self.bias_pot_obj_list = []

    def main(self, e, g, geom_num_list, element_list, force_data, pre_B_g="", iter="", initial_geom_num_list=""):
        """-> (bias_grad (N,3), B_e, B_g (N,3), bias_hessian (3N,3N)), potential.py:202."""
        for key in active_keys(force_data):
            if key not in _HANDLED:
                raise MopError(f"bias potential '{key}' is active but not built on the device "
                               f"(handled: {sorted(_HANDLED)}); refusing to drop it silently")
        geom = np.asarray(geom_num_list, dtype=np.float64)
        N = geom.shape[0]
        bias_grad = np.zeros_like(geom)
        bias_hess = np.zeros((3 * N, 3 * N))
        B_e = 0.0
        for i in range(len(force_data.get("AFIR_gamma", []))):
            gam = force_data["AFIR_gamma"][i]
            if 0.0 in gam:                                   # potential.py:469
                continue
            gval = gradually_change_param(gam[0], gam[1], iter) if (len(gam) == 2 and iter != "") else gam[0]
            pot = AFIRPotential(AFIR_Fragm_1=force_data["AFIR_Fragm_1"][i], AFIR_Fragm_2=force_data["AFIR_Fragm_2"][i],
                                element_list=element_list, device=self.device)
            E, gr, H = pot.calc_energy_grad_hess(geom, [gval])
            B_e += float(E.item())
            bias_grad = bias_grad + gr.cpu().numpy()
            bias_hess = bias_hess + H.cpu().numpy()
        # restraints (potential.py:640-672,742-752): one launch for all terms
        terms = []
        for i, k in enumerate(force_data.get("keep_pot_spring_const", [])):
            if k != 0.0:
                a, b = force_data["keep_pot_atom_pairs"][i]
                terms.append((ops.BIAS_KEEP, [a - 1], [b - 1], float(k), float(force_data["keep_pot_distance"][i])))
        for i, k in enumerate(force_data.get("keep_pot_v2_spring_const", [])):
            if 0.0 not in k:
                terms.append((ops.BIAS_KEEP_V2, [a - 1 for a in force_data["keep_pot_v2_fragm1"][i]],
                              [a - 1 for a in force_data["keep_pot_v2_fragm2"][i]], float(k[0]),
                              float(force_data["keep_pot_v2_distance"][i][0])))
        if N > 2:
            for i, k in enumerate(force_data.get("keep_angle_spring_const", [])):
                if k != 0.0:
                    terms.append((ops.BIAS_KEEP_ANGLE, [a - 1 for a in force_data["keep_angle_atom_pairs"][i]], [],
                                  float(k), float(force_data["keep_angle_angle"][i])))
        if N > 3:                                        # potential.py:779-789 (parameters go in as float64)
            for i, k in enumerate(force_data.get("keep_dihedral_angle_spring_const", [])):
                if k != 0.0:
                    phi0 = float(torch.deg2rad(torch.tensor(float(force_data["keep_dihedral_angle_angle"][i]), dtype=torch.float64)))
                    terms.append((ops.BIAS_KEEP_DIHEDRAL, [a - 1 for a in force_data["keep_dihedral_angle_atom_pairs"][i]],
                                  [], float(k), phi0))
        from ..Parameters import tables
        for i, k in enumerate(force_data.get("anharmonic_keep_pot_spring_const", [])):      # potential.py:673-685
            if k != 0.0:
                depth = float(force_data["anharmonic_keep_pot_potential_well_depth"][i])
                if depth != 0.0:                                   # (zero depth: the reference's energy is the constant 0)
                    a, b = force_data["anharmonic_keep_pot_atom_pairs"][i]
                    terms.append((ops.BIAS_ANHARMONIC_KEEP, [a - 1], [b - 1], float(k),
                                  float(force_data["anharmonic_keep_pot_distance"][i]), [depth]))
        for i, wv in enumerate(force_data.get("well_pot_wall_energy", [])):                  # potential.py:687-698
            if wv != 0.0:
                lim = [float(v) / tables.BOHR2ANG for v in force_data["well_pot_limit_dist"][i]]
                terms.append((ops.BIAS_WELL, [a - 1 for a in force_data["well_pot_fragm_1"][i]],
                              [a - 1 for a in force_data["well_pot_fragm_2"][i]], float(wv) / tables.HARTREE2KJMOL, 0.0, lim))
        for i, wv in enumerate(force_data.get("wall_well_pot_wall_energy", [])):             # potential.py:700-708
            if wv != 0.0:
                lim = [float(v) / tables.BOHR2ANG for v in force_data["wall_well_pot_limit_dist"][i]]
                axis = {"x": 0, "y": 1, "z": 2}[force_data["wall_well_pot_direction"][i]]
                for a in force_data["wall_well_pot_target"][i]:
                    terms.append((ops.BIAS_WELL_WALL, [a - 1], [axis], float(wv) / tables.HARTREE2KJMOL, 0.0, lim))
        for i, wv in enumerate(force_data.get("void_point_well_pot_wall_energy", [])):       # potential.py:713-721
            if wv != 0.0:
                lim = [float(v) / tables.BOHR2ANG for v in force_data["void_point_well_pot_limit_dist"][i]]
                # the reference stores the point as a float32 tensor (switching_potential.py:137)
                pt = [float(np.float32(float(v))) for v in force_data["void_point_well_pot_coordinate"][i]]
                for a in force_data["void_point_well_pot_target"][i]:
                    terms.append((ops.BIAS_WELL_POINT, [a - 1], [], float(wv) / tables.HARTREE2KJMOL, 0.0, lim, pt))
        for i, wv in enumerate(force_data.get("around_well_pot_wall_energy", [])):           # potential.py:727-735
            if wv != 0.0:
                lim = [float(v) / tables.BOHR2ANG for v in force_data["around_well_pot_limit_dist"][i]]
                cen = [a - 1 for a in force_data["around_well_pot_center"][i]]
                for a in force_data["around_well_pot_target"][i]:        # one fragment-well term per target atom
                    terms.append((ops.BIAS_WELL, [a - 1], cen, float(wv) / tables.HARTREE2KJMOL, 0.0, lim))
        for i, wv in enumerate(force_data.get("repulsive_potential_well_scale", [])):        # potential.py:574-604
            if wv != 0.0:
                terms += lj_pair_terms(element_list, force_data["repulsive_potential_Fragm_1"][i],
                                       force_data["repulsive_potential_Fragm_2"][i], float(wv),
                                       float(force_data["repulsive_potential_dist_scale"][i]),
                                       force_data["repulsive_potential_unit"][i])
        if N > 3:                                                                            # potential.py:796-809
            for i, k in enumerate(force_data.get("keep_out_of_plain_angle_spring_const", [])):
                if k != 0.0:
                    phi0 = float(torch.deg2rad(torch.tensor(float(force_data["keep_out_of_plain_angle_angle"][i]), dtype=torch.float64)))
                    terms.append((ops.BIAS_KEEP_OOP, [a - 1 for a in force_data["keep_out_of_plain_angle_atom_pairs"][i]],
                                  [], float(k), phi0))
        rad = lambda deg: float(torch.deg2rad(torch.tensor(float(deg), dtype=torch.float64)))
        for i, k in enumerate(force_data.get("keep_angle_v2_spring_const", [])):             # potential.py:758-772
            if 0.0 not in k:
                fr = [force_data[f"keep_angle_v2_fragm{j}"][i] for j in (1, 2, 3)]
                terms.append(centroid_term(ops.BIAS_KEEP_ANGLE_V2, fr, k[0], force_data["keep_angle_v2_angle"][i][0]))
        for i, k in enumerate(force_data.get("keep_dihedral_angle_v2_spring_const", [])):    # potential.py:812-827
            if 0.0 not in k:
                fr = [force_data[f"keep_dihedral_angle_v2_fragm{j}"][i] for j in (1, 2, 3, 4)]
                terms.append(centroid_term(ops.BIAS_KEEP_DIHEDRAL_V2, fr, k[0], rad(force_data["keep_dihedral_angle_v2_angle"][i][0])))
        for i, k in enumerate(force_data.get("keep_out_of_plain_angle_v2_spring_const", [])):  # potential.py:862-880
            if 0.0 not in k:
                fr = [force_data[f"keep_out_of_plain_angle_v2_fragm{j}"][i] for j in (1, 2, 3, 4)]
                terms.append(centroid_term(ops.BIAS_KEEP_OOP_V2, fr, k[0], rad(force_data["keep_out_of_plain_angle_v2_angle"][i][0])))
        if terms:
            dev = torch.device(self.device)
            xyz = torch.as_tensor(np.ascontiguousarray(geom)).reshape(1, N, 3).to(dev)
            E, gr, H = ops.bias_terms(xyz, ops.pack_bias_terms(terms, dev), len(terms))
            B_e += float(E[0].item())
            bias_grad = bias_grad + gr[0].cpu().numpy().reshape(N, 3)
            bias_hess = bias_hess + H[0].cpu().numpy()
        return bias_grad, B_e + e, g + bias_grad, bias_hess
