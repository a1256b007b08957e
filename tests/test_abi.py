"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares."""
import ctypes
import os
import re

import armour_b200 as ab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    syms = set()
    for name in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if not name.endswith(".h"):
            continue
        text = open(os.path.join(ROOT, "include", name)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        syms.update(re.findall(r"\b(armour_[a-z0-9_]+)\s*\(", text))
    return sorted(syms)


def test_header_declares_the_tnlp_callbacks():
    syms = declared_symbols()
    for name in ("armour_get_nlp_info", "armour_get_bounds_info", "armour_get_starting_point", "armour_eval_f", "armour_eval_grad_f",
                 "armour_eval_g", "armour_eval_jac_g", "armour_check_feasible", "armour_build", "armour_get_pz"):
        assert name in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(ab.LIB_PATH), "libarmour_b200.so missing: run __graft_entry__.build()"
    L = ctypes.CDLL(ab.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, missing
    from armour_b200 import controller
    assert sorted(ab.EXPORTS + controller.EXPORTS) == declared_symbols()


def test_default_config_matches_reference_macros():
    cfg = ab.default_config()
    assert cfg.num_time_steps == 128          # KPR/Parameters.h:17
    assert abs(cfg.k_range[0] - 3.141592653589793 / 48) < 1e-18   # :21
    assert cfg.simplify_threshold == 5e-4     # :10
    assert cfg.max_obstacles == 40            # :26
    assert cfg.mass_uncertainty == 0.03 and cfg.inertia_uncertainty == 0.03


def test_argument_validation_needs_no_gpu():
    L = ab.lib()
    assert L.armour_create(None, None) == -1                      # ARMOUR_E_INVALID
    assert b"null" in L.armour_last_error()
    L.armour_get_nlp_info.argtypes = [ctypes.c_void_p] * 5
    assert L.armour_get_nlp_info(None, None, None, None, None) == -1


def test_product_path_never_references_the_oracle():
    """The product must not link, load or call anything under oracle/ (no CPU fallback)."""
    pkg = os.path.join(ROOT, "armour-dev_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".hpp", ".cpp", ".py", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "oracle_build" not in text and "_oracle" not in text, os.path.join(dirpath, f)
