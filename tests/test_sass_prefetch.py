"""Static regression guard for the outlier-record prefetch of the fast attention kernels (no GPU needed: reads the SASS of the
built library with cuobjdump).  The prefetch loads of the NEXT tile must not be followed by an instruction that waits on their
scoreboard without consuming them — round 2 lost 37 % on the M=32 kernel to exactly that (tools/sass_sb_check.py,
profiles/r02_dm4_scoreboard_regression.txt)."""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
LIB = os.path.join(ROOT, "million_b200", "libmillion_b200.so")

KERNELS = [
    "_ZN7million20attn_fast_dm4_kernelI6__halfLi1ELi0ELi1EEEvNS_8AttnArgsEPKji",      # M=32, MHA, K records (the regressed one)
    "_ZN7million20attn_fast_dm4_kernelI6__halfLi4ELi0ELi3EEEvNS_8AttnArgsEPKji",      # M=32, GQA-4, K and V records
    "_ZN7million16attn_fast_kernelI6__halfLi4ELi0ELi1ELi0EEEvNS_8AttnArgsEPKji",      # M=64, GQA-4, K records (bench extra)
    "_ZN7million16attn_fast_kernelI6__halfLi2ELi0ELi3ELi0EEEvNS_8AttnArgsEPKji",      # M=64, 2-head groups, K and V records
]


@pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB), reason="needs cuobjdump and the built library")
@pytest.mark.parametrize("fn", KERNELS)
def test_outlier_prefetch_is_not_waited_on_early(fn):
    import sass_sb_check
    # the record loads of the 1/2-records-per-token paths: predicated 16/32-bit read-only loads without an immediate offset
    # (the look-up-table build also uses read-only loads; those are consumed in place by design)
    n, hz = sass_sb_check.hazards(LIB, fn, maxd=256, loads=r'^@!?P\d LDG\.E(\.U16)?\.CONSTANT R\d+, desc\[UR\d+\]\[R\d+\.64\]$')
    assert n > 1000, f"kernel {fn} not found in {LIB}"
    assert not hz, "prefetch loads waited on right after issue:\n" + "\n".join(f"{a:06x} {t} -> +{d} {nt}" for a, t, d, nt in hz)


@pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB), reason="needs cuobjdump and the built library")
def test_tensor_core_encoder_is_tcgen05_and_tma_in_the_shipped_binary():
    """The shipped sm_100a code of the tensor-core encoder really issues tcgen05 MMAs with TMEM accumulators and a TMA bulk copy
    (UTCHMMA / LDTM / UBLKCP in SASS), not a recompiled mma.sync path; the decode kernel streams its code tiles with cp.async
    (LDGSTS) and gathers from shared memory (profiles/r02_sass_mnemonics_library.txt is the whole-library histogram)."""
    import re
    import subprocess

    def mnemonics(fn):
        text = subprocess.run(["cuobjdump", "-sass", "-fun", fn, LIB], capture_output=True, text=True).stdout
        ops = re.findall(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', text, re.M)
        return {o.split('.')[0] for o in ops}, len(ops)

    enc, n = mnemonics("_ZN7million2tc16encode_tc_kernelI6__halfLi4EEEvNS0_7EncArgsE")
    assert n > 500 and {"UTCHMMA", "UTCBAR", "LDTM", "UBLKCP"} <= enc and "HMMA" not in enc
    att, n = mnemonics("_ZN7million16attn_fast_kernelI6__halfLi4ELi0ELi0ELi0EEEvNS_8AttnArgsEPKji")
    assert n > 5000 and {"LDGSTS", "LDS", "FHADD", "HFMA2", "PRMT"} <= att
