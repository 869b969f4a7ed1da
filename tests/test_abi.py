"""The C-ABI library loads on a CPU-only box and exports every symbol include/asr_b200.h declares
(no compute calls here)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "asr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(asr_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for must in ("asr_create", "asr_features", "asr_encode", "asr_decode_greedy", "asr_decode_beam",
                 "asr_transcribe", "asr_set_lm"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from chinese_asr_b200 import _cabi
    for name in declared_symbols():
        assert hasattr(_cabi.lib, name), f"libasr_b200.so does not export {name}"
        assert name in _cabi.SIGNATURES, f"ctypes binding lacks {name}"
    assert _cabi.lib.asr_version() >= 100


def test_num_frames_matches_oracle():
    from chinese_asr_b200 import _cabi
    from oracle import asr_oracle as O
    for n in (512, 513, 672, 673, 993, 8000, 20011, 32000, 80000, 160000, 320000):
        assert _cabi.lib.asr_num_frames(n) == O.num_frames(n) // 3


def test_no_product_import_of_oracle():
    """The product must never route through the oracle (or the reference tree)."""
    pkg = os.path.join(ROOT, "chinese_asr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "import oracle" not in s and "from oracle" not in s, f
                assert "/root/reference" not in s, f
