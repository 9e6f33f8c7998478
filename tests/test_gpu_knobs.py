"""The schedule knobs of the generic kernel (65..224 points; cgp_small.cu: CGP_DIAG_WARP, CGP_Z_IN_LOOP, CGP_LOOKAHEAD,
CGP_BIG_FWD, CGP_BIG_WARPS, CGP_MID_WARPS) select code paths that the defaults never run.  They are read once per
process, so each variant runs the CTA-per-object parity tests (C ABI against the oracle, rtol 1e-9) in a child process."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = [
    {"CGP_LOOKAHEAD": "1"},                                                   # warp 0 runs ahead of the solve phase
    {"CGP_DIAG_WARP": "0", "CGP_Z_IN_LOOP": "0", "CGP_BIG_FWD": "0"},         # round-1 schedule, explicit L^-1 above 128
    {"CGP_BIG_WARPS": "4"},                                                   # four warps above 128 points
    {"CGP_MID_WARPS": "8", "CGP_LOOKAHEAD": "1"},                             # eight warps from 65 points
]


@pytest.mark.parametrize("env", VARIANTS, ids=lambda e: ",".join("%s=%s" % kv for kv in sorted(e.items())))
def test_generic_kernel_variants(env):
    child = dict(os.environ); child.update(env)
    cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_cabi.py"), "-q", "-x", "-p", "no:cacheprovider",
           "-k", "test_sizes_cta_per_object or test_ragged_mixed_sizes_and_empty or test_not_positive_definite"]
    r = subprocess.run(cmd, cwd=ROOT, env=child, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
