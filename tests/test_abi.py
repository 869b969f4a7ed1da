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


def test_single_thread_instructions_are_issued_from_uniform_code():
    """Regression guard for the round-2 finding: tcgen05.mma / tcgen05.commit / TMA copies take
    their operands from uniform registers; when the issue loop sits under a divergent branch ptxas wraps every one
    of them in an ELECT + R2UR.BROADCAST + BRA.U.ANY loop that paces the tensor pipe.  The GEMM kernels must contain
    none of those loops, the recurrence kernels at most the one around their multicast copy."""
    import shutil
    import subprocess
    import pytest
    from chinese_asr_b200 import _cabi
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not installed")
    sass = subprocess.run(["cuobjdump", "-sass", _cabi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    counts, mma, name = {}, {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        if name is None:
            continue
        if "BRA.U.ANY" in line:
            counts[name] = counts.get(name, 0) + 1
        if "UTCHMMA" in line:
            mma[name] = mma.get(name, 0) + 1
    gemm = [n for n in mma if "gemm_split_pair_kernel" in n]
    rec = [n for n in mma if "lstm_rec_tc3_kernel" in n]
    assert gemm and rec, "tcgen05 kernels not found in the library"
    for n in gemm:
        assert counts.get(n, 0) == 0, (n, counts.get(n))
    for n in rec:
        assert counts.get(n, 0) <= 1, (n, counts.get(n))
