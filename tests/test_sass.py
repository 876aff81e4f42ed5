"""The shipped library really is tcgen05 / TMA code: every hot kernel of the CViT path (and of the "next" engines) carries
UTCHMMA (tcgen05.mma) and UTMALDG (TMA loads) in its SASS and no legacy HMMA (mma.sync).  Runs on CPU: cuobjdump only
disassembles fac_fake_b200/libfacfake.so (profiles/r02_sass_table_final.txt is the same scan over the objects)."""
import collections
import os
import re
import shutil
import subprocess

import pytest

from fac_fake_b200 import _lib

HOT = ["c12_kernel", "ws2conv_kernel", "ws2x_conv_kernel", "ptcw_conv_kernel", "ptc2_conv_kernel", "xf_kernel", "tc_gemm_kernel",
       "rvk_conv2_kernel", "rvk_stem_kernel"]


CUOBJDUMP = shutil.which("cuobjdump") or ("/usr/local/cuda/bin/cuobjdump" if os.path.exists("/usr/local/cuda/bin/cuobjdump") else None)


@pytest.mark.skipif(CUOBJDUMP is None, reason="cuobjdump not found")
def test_hot_kernels_are_tcgen05_and_tma():
    sass = subprocess.run([CUOBJDUMP, "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    count = collections.defaultdict(lambda: collections.Counter())
    fn = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if fn and m:
            op = m.group(1)
            for key in ("UTCHMMA", "UTMALDG", "LDTM", "HMMA"):
                if op.startswith(key):
                    count[fn][key] += 1
    assert count, "no SASS found in the library"
    for kernel in HOT:
        fns = [f for f in count if kernel in f]
        assert fns, f"{kernel} is not in the library"
        for f in fns:
            c = count[f]
            assert c["UTCHMMA"] > 0 and c["UTMALDG"] > 0 and c["LDTM"] > 0, (f, dict(c))
    assert all(c["HMMA"] == 0 for c in count.values()), "legacy mma.sync in the library"
    # layers 7-9: the filter tile is the M operand, 256 pixels the N operand -> four N = 256 MMAs per filter tile
    assert all(count[f]["UTCHMMA"] == 4 for f in count if "ptcw_conv_kernel" in f)
