"""The C-ABI library loads and exports every symbol include/*.h declares (no GPU needed)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if not fn.endswith(".h"):
            continue
        src = open(os.path.join(ROOT, "include", fn)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        for m in re.finditer(r"\b(av1r_[a-z0-9_]+)\s*\(", src):
            names.add(m.group(1))
    return names


def test_all_declared_symbols_exported(built):
    lib = C.CDLL(built[0])
    names = _declared()
    assert "av1r_open" in names and "av1r_stage_film_grain" in names
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_abi_version_and_defaults(built):
    import av1recon
    l = av1recon.lib()
    assert l.av1r_abi_version() == 0x00010000
    cfg = av1recon.Config()
    l.av1r_default_config(C.byref(cfg))
    assert cfg.struct_size == C.sizeof(av1recon.Config)
    assert cfg.apply_grain == 1 and cfg.inloop_filters == 7
