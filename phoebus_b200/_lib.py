"""ctypes binding of libphoebus_b200.so (include/phoebus_b200.h).

The library is the product: there is no CPU or PyTorch fallback. Importing this module loads the
shared object (so that the symbol table can be checked on a machine without a GPU); running any
task without the library or without a CUDA device raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PHOEBUS_B200_LIB") or os.path.join(HERE, "libphoebus_b200.so")  # env override: tuning variants

PB_NGLOC, PB_NGDER = 21, 33
PB_GEOM_WORDS = 4 * PB_NGLOC + PB_NGDER
PB_NBR_NONE, PB_NBR_REMOTE = -1, -2
TASK_NAMES = ("flux", "ct_fix", "source", "update", "stale_fixup_floors", "halo", "bcs", "c2p", "fixup_floors",
              "recon", "riemann")  # the last two are nested inside "flux"


class PbBlock(C.Structure):
    _fields_ = [("xmin", C.c_double * 3), ("dx", C.c_double * 3), ("bc", C.c_int * 6), ("geom_slot", C.c_int),
                ("floor_slot", C.c_int), ("nbr", C.c_int * 27), ("pad_", C.c_int)]


class PbCfg(C.Structure):
    _fields_ = [
        ("ndim", C.c_int), ("nx", C.c_int * 3), ("ng", C.c_int), ("mhd", C.c_int),
        ("recon", C.c_int), ("riemann", C.c_int),
        ("c2p_tol", C.c_double), ("c2p_max_iter", C.c_int), ("c2p_floor_scale_fac", C.c_double),
        ("c2p_fail_on_floors", C.c_int), ("c2p_fail_on_ceilings", C.c_int),
        ("floor_type", C.c_int),
        ("rho0", C.c_double), ("sie0", C.c_double), ("rexp", C.c_double), ("sexp", C.c_double),
        ("rho_min_eos", C.c_double), ("sie_min_eos", C.c_double),
        ("gam_max", C.c_double), ("sie_max", C.c_double),
        ("bsqorho_max", C.c_double), ("bsqou_max", C.c_double),
        ("enable_floors", C.c_int), ("enable_mhd_floors", C.c_int), ("enable_c2p_fixup", C.c_int),
        ("enable_flux_fixup", C.c_int), ("c2p_failure_strategy", C.c_int),
        ("gm1", C.c_double), ("Cv", C.c_double),
        ("gk", C.c_int), ("gj", C.c_int), ("gi", C.c_int),
        ("bc_prim", C.c_int), ("ox1_inflow_check", C.c_int),
    ]


class PbHaloPlan(C.Structure):
    _fields_ = [("send_plan", C.c_void_p), ("nsend", C.c_int), ("recv_plan", C.c_void_p), ("nrecv", C.c_int),
                ("send_buf", C.c_void_p), ("recv_buf", C.c_void_p), ("npeers", C.c_int), ("peer", C.POINTER(C.c_int)),
                ("send_off", C.POINTER(C.c_longlong)), ("send_cnt", C.POINTER(C.c_longlong)),
                ("recv_off", C.POINTER(C.c_longlong)), ("recv_cnt", C.POINTER(C.c_longlong))]


class PbPack(C.Structure):
    _fields_ = [("nblocks", C.c_int), ("blocks", C.c_void_p), ("geom", C.c_void_p), ("floor_tab", C.c_void_p),
                ("prim", C.c_void_p), ("temp", C.c_void_p), ("c2p_mu", C.c_void_p), ("fail", C.c_void_p),
                ("csig", C.c_void_p), ("fsig", C.c_void_p), ("flux", C.c_void_p), ("emf", C.c_void_p),
                ("zflags", C.c_void_p), ("bc_blocks", C.c_void_p), ("n_bc_blocks", C.c_int), ("gsrc", C.c_void_p),
                ("inner_blocks", C.c_void_p), ("n_inner_blocks", C.c_int), ("recon_lr", C.c_void_p),
                ("ghost_mask", C.c_void_p), ("c2p_zones", C.c_void_p), ("n_c2p_zones", C.c_longlong),
                ("twin_zones", C.c_void_p), ("n_twin", C.c_longlong), ("twin_buf", C.c_void_p), ("twin_flags", C.c_void_p)]


# every symbol include/phoebus_b200.h declares: name -> (restype, argtypes)
_P, _CFG, _PACK, _I2 = C.c_void_p, C.POINTER(PbCfg), C.POINTER(PbPack), C.POINTER(C.c_int)
_D3 = C.POINTER(C.c_double)
SYMBOLS = {
    "pb_version": (C.c_char_p, []),
    "pb_last_error": (C.c_char_p, []),
    "pb_nprim": (C.c_int, [_CFG]),
    "pb_ncons": (C.c_int, [_CFG]),
    "pb_cells_per_block": (C.c_longlong, [_CFG]),
    "pb_device_sm_count": (C.c_int, []),
    "pb_build_geometry_host": (C.c_int, [C.c_int, _D3, _CFG, _D3, _D3, _P]),
    "pb_build_floor_table_host": (C.c_int, [_CFG, _D3, _D3, _P]),
    "pb_calculate_fluxes": (C.c_int, [_PACK, _CFG, _P]),
    "pb_flux_ct": (C.c_int, [_PACK, _CFG, _P]),
    "pb_fix_fluxes": (C.c_int, [_PACK, _CFG, _P]),
    "pb_flux_divergence": (C.c_int, [_PACK, _CFG, _P, _P]),
    "pb_source_terms": (C.c_int, [_PACK, _CFG, _P, _P]),
    "pb_weighted_sum": (C.c_int, [_P, _P, C.c_double, C.c_double, _P, C.c_longlong, _P]),
    "pb_update_fused": (C.c_int, [_PACK, _CFG, _P, _P, C.c_double, C.c_double, _P, _P, _P, C.c_int, _P]),
    "pb_update_from_sources": (C.c_int, [_PACK, _CFG, _P, _P, C.c_double, C.c_double, _P, _P, _P, C.c_int, _P]),
    "pb_cons_to_prim": (C.c_int, [_PACK, _CFG, _P, _I2, _I2, _I2, _P]),
    "pb_c2p_fixup": (C.c_int, [_PACK, _CFG, _P, _P]),
    "pb_apply_floors": (C.c_int, [_PACK, _CFG, _P, _P, _P]),
    "pb_scan_zone_flags": (C.c_int, [_PACK, _CFG, _P]),
    "pb_prim_to_cons": (C.c_int, [_PACK, _CFG, _P, _I2, _I2, _I2, _P]),
    "pb_halo_exchange_local": (C.c_int, [_PACK, _CFG, _P, _P]),
    "pb_halo_exchange_local_ex": (C.c_int, [_PACK, _CFG, _P, C.c_int, _P]),
    "pb_ghost_plan_host": (C.c_int, [_CFG, _P, C.c_int, _P, _P, _P, _P, C.POINTER(C.c_longlong), _P, C.POINTER(C.c_longlong)]),
    "pb_twin_entries": (C.c_longlong, [_CFG, C.c_int]),
    "pb_ghost_plan_active": (C.c_int, [_PACK, _CFG]),
    "pb_blocks_gather": (C.c_int, [_P, _P, C.c_int, C.c_longlong, _P]),
    "pb_blocks_scatter": (C.c_int, [_P, _P, C.c_int, C.c_longlong, _P]),
    "pb_halo_nvar": (C.c_int, [_CFG]),
    "pb_halo_pack": (C.c_int, [_PACK, _CFG, _P, _P, C.c_int, _P, _P]),
    "pb_halo_unpack": (C.c_int, [_PACK, _CFG, _P, _P, C.c_int, _P, _P]),
    "pb_comm_unique_id": (C.c_int, [C.c_char_p]),
    "pb_comm_init": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_char_p]),
    "pb_comm_destroy": (C.c_int, [_P]),
    "pb_halo_exchange_begin": (C.c_int, [_PACK, _CFG, _P, C.POINTER(PbHaloPlan), _P, _P]),
    "pb_halo_exchange_end": (C.c_int, [_PACK, _CFG, _P, C.POINTER(PbHaloPlan), _P, _P]),
    "pb_comm_allreduce_min": (C.c_int, [_P, _P, C.c_int, _P]),
    "pb_dt_control": (C.c_int, [_P, C.c_double, _P, _P]),
    "pb_stage_graph_create": (C.c_int, [_PACK, _CFG, _P, _P, _P, C.c_double, _P, _P, C.POINTER(PbHaloPlan), _P, _P, C.POINTER(C.c_void_p)]),
    "pb_stage_graph_launch": (C.c_int, [_P, _P]),
    "pb_stage_graph_destroy": (C.c_int, [_P]),
    "pb_apply_bcs": (C.c_int, [_PACK, _CFG, _P, _P]),
    "pb_apply_bcs_ex": (C.c_int, [_PACK, _CFG, _P, C.c_int, _P]),
    "pb_pgen_torus_table_host": (C.c_int, [_D3, _D3, _CFG, _D3, _D3, _P, _P]),
    "pb_pgen_torus_fill": (C.c_int, [_PACK, _CFG, _P, C.c_double, C.c_ulonglong, _P]),
    "pb_pgen_torus_bfield": (C.c_int, [_PACK, _CFG, _P]),
    "pb_pgen_betas": (C.c_int, [_PACK, _CFG, C.c_double, _P, _D3, _P]),
    "pb_scale_bfield": (C.c_int, [_PACK, _CFG, C.c_double, _P]),
    "pb_calculate_divb": (C.c_int, [_PACK, _CFG, _P, _P, _P]),
    "pb_history_fluid": (C.c_int, [_PACK, _CFG, _P, _P, _D3, _P]),
    "pb_history_torus": (C.c_int, [_PACK, _CFG, _P, C.c_double, C.c_double, _P, _D3, _P]),
    "pb_estimate_dt": (C.c_int, [_PACK, _CFG, C.c_int, _P, _P]),
    "pb_profile_enable": (C.c_int, [C.c_int]),
    "pb_profile_read": (C.c_int, [_P, _P]),
    "pb_selftest_arith": (C.c_int, [C.c_longlong, C.c_ulonglong, C.c_int, _P, _P]),
    "pb_arith_redo_count": (C.c_ulonglong, []),
    "pb_c2p_stats": (C.c_int, [C.c_int, _P]),
    "pb_c2p_unbracketed_count": (C.c_longlong, [C.c_int]),
    "pb_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "pb_get_option": (C.c_int, [C.c_char_p]),
    "pb_selftest_riemann": (C.c_int, [C.c_longlong, C.c_ulonglong, _CFG, C.c_int, _P, _P, _P]),
    "pb_rk_stage": (C.c_int, [_PACK, _CFG, _P, _P, _P, C.c_double, C.c_double, _P, _P]),
    "pb_rk_stage_pre": (C.c_int, [_PACK, _CFG, _P, _P, _P, C.c_double, C.c_double, _P, _P]),
    "pb_rk_stage_post": (C.c_int, [_PACK, _CFG, _P, _P, _P]),
    "pb_rk_stage_post_early": (C.c_int, [_PACK, _CFG, _P, _P]),
    "pb_early_c2p_pays": (C.c_int, [_PACK]),
    "pb_rk_stage_post2": (C.c_int, [_PACK, _CFG, _P, _P, C.c_int, _P]),
}

_lib = None


class PhoebusB200Error(RuntimeError):
    pass


def lib():
    """Load the C-ABI library; raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PhoebusB200Error("%s is missing: build it with `make -C phoebus_b200/csrc` "
                                   "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise PhoebusB200Error("%s failed (%d): %s" % (what, rc, lib().pb_last_error().decode()))
