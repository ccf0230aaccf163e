"""CPU tests: the C-ABI library builds for sm_100a, loads, and exports every symbol of the header."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _header_symbols():
    text = (ROOT / "include" / "kamino_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kmn_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    from kamino_b200 import build, ffi
    path = build.build()
    assert path.exists()
    lib = ctypes.CDLL(str(path))
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/kamino_b200.h but not exported"
    assert sorted(ffi.ABI_SYMBOLS) == syms
    lib.kmn_abi_version.restype = ctypes.c_uint32
    declared = int(re.search(r"kmn_abi_version\(void\);\s*/\* currently (\d+)", (ROOT / "include" / "kamino_b200.h").read_text()).group(1))
    assert lib.kmn_abi_version() == declared


def test_hit_layout_matches_header():
    from kamino_b200 import ffi
    assert ffi.HIT_DTYPE.itemsize == 24
    assert [ffi.HIT_DTYPE.fields[n][1] for n in ("kmer", "rec", "seg_start", "start", "occ")] == [0, 8, 12, 16, 20]


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device kmn_create must fail loudly (never fall back to a host path)."""
    import torch
    from kamino_b200 import ffi
    if torch.cuda.is_available():
        return
    try:
        ffi.Context(13, "sr6")
    except ffi.KaminoError as e:
        assert "no CUDA device" in str(e) or "CUDA" in str(e)
    else:
        raise AssertionError("kmn_create succeeded without a GPU")


def test_product_never_references_oracle():
    pkg = ROOT / "kamino_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.cpp")):
        txt = p.read_text()
        assert "oracle_ffi" not in txt and "kamino_oracle" not in txt and "orc_" not in txt, p


def test_integration_doc_lists_every_export():
    """INTEGRATION.md's Rust extern "C" block is a complete mirror of the header."""
    doc = (ROOT / "INTEGRATION.md").read_text()
    block = doc[doc.index('#[link(name = "kamino_b200")]'):doc.index("pub fn check(rc: c_int)")]
    declared = set(re.findall(r"pub fn (kmn_[a-z0-9_]+)\(", block))
    assert declared == set(_header_symbols()), sorted(set(_header_symbols()) ^ declared)
