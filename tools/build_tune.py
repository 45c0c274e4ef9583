"""Tuning build of K1 (-DSR_TUNING): tools/libct_tune.so with the experimental configurations of
csrc/ct_tuning_variants.inc.  Prints registers / spills per configuration.  Not part of the product."""
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spinrelax_b200 import build  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libct_tune.so")


def main():
    import contextlib
    import io
    err = io.StringIO()
    with contextlib.redirect_stderr(err):
        build.build_tuning(OUT, verbose=True)
    txt = err.getvalue()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"CtCfgILi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)E", line)
        if m and "Compiling entry" in line:
            cur = tuple(int(x) for x in m.groups())
        m2 = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m2 and cur:
            spill = (int(m2.group(1)), int(m2.group(2)))
        m3 = re.search(r"Used (\d+) registers", line)
        if m3 and cur:
            print("R=%d MB=%d FB=%d NW=%d MINB=%d NS=%d FLUSH=%d DIAG=%d  regs=%s spill=%s" % (cur + (m3.group(1), spill)))
            cur = None
    print(OUT)


if __name__ == "__main__":
    main()
