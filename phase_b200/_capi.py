"""ctypes binding of libphase_b200.so (the C ABI in include/phase_b200.h).

The library is the product: there is no Python/CPU fallback.  Importing this
module fails loudly when the shared object is missing, and creating a context
fails loudly when no CUDA device is present.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libphase_b200.so")

# every symbol include/phase_b200.h declares: (restype, argtypes)
vp, ci, cd, cll = C.c_void_p, C.c_int, C.c_double, C.c_longlong
pi, pd, pvp = C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_void_p)
cs = C.c_char_p
SIGNATURES = {
    "phb_last_error": (cs, []),
    "phb_version": (ci, []),
    "phb_ctx_create": (ci, [ci, pvp]),
    "phb_ctx_destroy": (ci, [vp]),
    "phb_comm_unique_id": (ci, [vp]),
    "phb_ctx_init_comm": (ci, [vp, ci, ci, vp]),
    "phb_ctx_peer_arena_create": (ci, [vp, cll, ci, vp]),
    "phb_ctx_peer_arena_open": (ci, [vp, vp]),
    "phb_mesh_set_peer_layout": (ci, [vp, pi, pi]),
    "phb_mesh_read_cgns": (ci, [vp, cs, pvp]),
    "phb_mesh_refine": (ci, [vp, vp, ci, pvp]),
    "phb_mesh_patch_name": (ci, [vp, ci, cs, ci]),
    "phb_ctx_rank": (ci, [vp]),
    "phb_ctx_nprocs": (ci, [vp]),
    "phb_ctx_sync": (ci, [vp]),
    "phb_ctx_kernel_launches": (cll, [vp]),
    "phb_ctx_stream": (vp, [vp]),
    "phb_mesh_create": (ci, [vp, ci, pd, ci, pi, pi, pvp]),
    "phb_mesh_create_rectilinear": (ci, [vp, ci, ci, cd, cd, pvp]),
    "phb_mesh_create_triangulated": (ci, [vp, ci, ci, cd, cd, pvp]),
    "phb_mesh_add_patch_by_nodes": (ci, [vp, cs, ci, pi]),
    "phb_mesh_patch_id": (ci, [vp, cs]),
    "phb_mesh_finalize": (ci, [vp]),
    "phb_mesh_destroy": (ci, [vp]),
    "phb_mesh_sizes": (ci, [vp, C.POINTER(cll)]),
    "phb_mesh_get_i32": (cll, [vp, cs, pi, cll]),
    "phb_mesh_get_f64": (cll, [vp, cs, pd, cll]),
    "phb_partition_rcb": (ci, [vp, ci, pi]),
    "phb_partition_metis": (ci, [vp, ci, ci, pi, C.POINTER(C.c_longlong)]),
    "phb_partition_file_build": (ci, [vp, pi, ci, cd, pvp]),
    "phb_partition_file_sizes": (ci, [vp, C.POINTER(C.c_longlong)]),
    "phb_partition_file_get": (ci, [vp, pi, pi, pd, pi, pi]),
    "phb_partition_file_patch": (cll, [vp, ci, C.c_char_p, ci, pi]),
    "phb_partition_file_destroy": (ci, [vp]),
    "phb_mesh_create_local": (ci, [vp, vp, pi, pvp]),
    "phb_mesh_create_rect_strip": (ci, [vp, ci, ci, cd, cd, pvp]),
    "phb_mesh_create_rect_block": (ci, [vp, ci, ci, cd, cd, ci, ci, pvp]),
    "phb_solver_create": (ci, [vp, pvp]),
    "phb_solver_destroy": (ci, [vp]),
    "phb_solver_setup": (ci, [vp, cs, cs]),
    "phb_solver_set_rank": (ci, [vp, ci, ci]),
    "phb_solver_set_csr": (ci, [vp, ci, pi, pi, pd]),
    "phb_solver_set_coo": (ci, [vp, ci, cll, pi, pi, pd]),
    "phb_solver_set_halo": (ci, [vp, vp, ci]),
    "phb_solver_set_rhs": (ci, [vp, pd, ci]),
    "phb_solver_set_guess": (ci, [vp, pd, ci]),
    "phb_solver_solve": (ci, [vp, pi, pd]),
    "phb_solver_get_x": (ci, [vp, pd, ci]),
    "phb_solver_spmv": (ci, [vp, pd, pd, ci]),
    "phb_solver_time_spmv": (ci, [vp, ci, pd]),
    "phb_solver_apply_preconditioner": (ci, [vp, pd, pd, ci]),
    "phb_solver_bytes": (ci, [vp, pd]),
    "phb_solver_amg_info": (ci, [vp, pd]),
    "phb_solver_amg_refresh": (ci, [vp]),
    "phb_solver_amg_refresh_info": (ci, [vp, pd]),
    "phb_solver_amg_values": (cll, [vp, ci, ci, pd, cll]),
    "phb_solver_time_amg": (ci, [vp, ci, pd]),
    "phb_amg_host_build": (ci, [ci, pi, pi, pd, cd, ci, pvp]),
    "phb_amg_host_build_ex": (ci, [ci, pi, pi, pd, cd, cd, cd, ci, pvp]),
    "phb_amg_host_levels": (ci, [vp, pi, pi, pi]),
    "phb_amg_host_level_size": (ci, [vp, ci, ci, pi, pi, C.POINTER(cll), pd]),
    "phb_amg_host_level_csr": (ci, [vp, ci, ci, pi, pi, pd]),
    "phb_amg_host_level_weight": (ci, [vp, ci, pd]),
    "phb_amg_host_coarse_inverse": (ci, [vp, pd]),
    "phb_amg_host_destroy": (ci, [vp]),
    "phb_amg_dist_build": (ci, [ci, ci, pi, pi, pd, pi, cd, ci, cll, pvp]),
    "phb_amg_dist_info": (ci, [vp, pi, pi, pi]),
    "phb_amg_dist_tail": (vp, [vp, ci]),
    "phb_amg_dist_matrix_size": (ci, [vp, ci, ci, ci, pi, C.POINTER(cll)]),
    "phb_amg_dist_matrix": (ci, [vp, ci, ci, ci, pi, pi, pd, pi]),
    "phb_amg_dist_halo": (ci, [vp, ci, ci, pi, pi, pi]),
    "phb_amg_dist_ghost_gids": (ci, [vp, ci, ci, pi]),
    "phb_amg_dist_tail_offsets": (ci, [vp, pi]),
    "phb_amg_dist_destroy": (ci, [vp]),
    "phb_field_create": (ci, [vp, ci, cs, pvp]),
    "phb_field_destroy": (ci, [vp]),
    "phb_field_set_bc": (ci, [vp, cs, ci, cd, cd]),
    "phb_field_set": (ci, [vp, cs, pd, cll]),
    "phb_field_get": (ci, [vp, cs, pd, cll]),
    "phb_field_fill": (ci, [vp, cd, cd]),
    "phb_field_save_previous": (ci, [vp]),
    "phb_field_interpolate_faces": (ci, [vp]),
    "phb_field_set_boundary_faces": (ci, [vp]),
    "phb_field_gradient": (ci, [vp, vp]),
    "phb_field_send_messages": (ci, [vp]),
    "phb_eqn_create": (ci, [vp, ci, pvp]),
    "phb_eqn_destroy": (ci, [vp]),
    "phb_eqn_zero": (ci, [vp]),
    "phb_assemble_ddt": (ci, [vp, vp, cd, vp, cd, cd]),
    "phb_assemble_div": (ci, [vp, vp, vp, cd, cd]),
    "phb_assemble_dive": (ci, [vp, vp, vp, cd, cd]),
    "phb_assemble_laplacian": (ci, [vp, cd, vp, vp, cd, cd]),
    "phb_assemble_src": (ci, [vp, vp, cd]),
    "phb_assemble_src_div": (ci, [vp, vp, cd]),
    "phb_cicsam_weights": (ci, [vp, vp, vp, cd, vp]),
    "phb_assemble_cicsam_div": (ci, [vp, vp, vp, vp, cd, cd]),
    "phb_cicsam_momentum_flux": (ci, [cd, cd, vp, vp, vp, vp]),
    "phb_assemble_ddt_cells": (ci, [vp, vp, cd, cd, ci, pi]),
    "phb_assemble_src_div_cells": (ci, [vp, vp, cd, ci, pi]),
    "phb_assemble_src_laplacian": (ci, [vp, cd, vp, vp, cd]),
    "phb_eqn_scale_rows": (ci, [vp, vp]),
    "phb_eqn_relax": (ci, [vp, vp, cd]),
    "phb_eqn_export_csr": (cll, [vp, ci, pi, pi, pd, pd]),
    "phb_eqn_solve": (ci, [vp, vp, vp, ci, pi, pd]),
    "phb_fs_create": (ci, [vp, cd, cd, pvp]),
    "phb_fs_destroy": (ci, [vp]),
    "phb_fs_field": (vp, [vp, cs]),
    "phb_fs_eqn": (vp, [vp, cs]),
    "phb_fs_solver": (vp, [vp, cs]),
    "phb_fs_setup": (ci, [vp, cs, cd]),
    "phb_fs_initialize": (ci, [vp]),
    "phb_fs_rebuild_faces": (ci, [vp, cd]),
    "phb_fs_assemble_u": (ci, [vp, cd]),
    "phb_fs_assemble_p": (ci, [vp, cd]),
    "phb_fs_step": (ci, [vp, cd, pd]),
    "phb_fs_max_time_step": (ci, [vp, cd, cd, cd, pd]),
    "phb_mp_create": (ci, [vp, cd, cd, cd, cd, cd, cd, cd, cd, pvp]),
    "phb_mp_destroy": (ci, [vp]),
    "phb_mp_field": (vp, [vp, cs]),
    "phb_mp_eqn": (vp, [vp, cs]),
    "phb_mp_solver": (vp, [vp, cs]),
    "phb_mp_setup": (ci, [vp, cs, cd]),
    "phb_mp_initialize": (ci, [vp]),
    "phb_mp_step": (ci, [vp, cd, pd]),
    "phb_piso_create": (ci, [vp, cd, cd, pvp]),
    "phb_piso_destroy": (ci, [vp]),
    "phb_piso_field": (vp, [vp, cs]),
    "phb_piso_solver": (vp, [vp, cs]),
    "phb_piso_setup": (ci, [vp, cs, cd]),
    "phb_piso_initialize": (ci, [vp]),
    "phb_piso_step": (ci, [vp, cd, pd]),
}

_lib = None


class PhaseB200Error(RuntimeError):
    """Mirrors the reference's Exception(class, method, description) (S/Exception.h:7-18)."""

    def __init__(self, code, message):
        super().__init__("phase_b200 [%d]: %s" % (code, message))
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libphase_b200.so is not built (%s). Run `python -m phase_b200.build` or "
                "__graft_entry__.build(); there is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc < 0:
        raise PhaseB200Error(rc, lib().phb_last_error().decode(errors="replace"))
    return rc
