"""Compile tests/integration/speedups_stub.c against libsafelife_b200.so (gcc, in place)."""
import os
import subprocess
import sysconfig

import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def build():
    from safelife_b200 import _lib
    _lib.build()
    libdir = os.path.dirname(_lib.LIB_PATH)
    out = os.path.join(HERE, "speedups_stub" + sysconfig.get_config_var("EXT_SUFFIX"))
    src = os.path.join(HERE, "speedups_stub.c")
    deps = [src, _lib.LIB_PATH, os.path.join(ROOT, "include", "safelife_b200.h")]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    cmd = ["gcc", "-O2", "-shared", "-fPIC", "-I", sysconfig.get_paths()["include"],
           "-I", os.path.join(ROOT, "include"), src, "-o", out,
           "-L", libdir, "-l:libsafelife_b200.so", "-Wl,-rpath," + libdir]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build())
