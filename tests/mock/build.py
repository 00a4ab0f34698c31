"""Builds tests/mock/libb200mock.so (the host test double of the C ABI; see b200mock.c).  Test infrastructure only."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def build():
    so, src = os.path.join(HERE, "libb200mock.so"), os.path.join(HERE, "b200mock.c")
    deps = [src, os.path.join(ROOT, "include", "petscb200.h"), os.path.join(ROOT, "oracle", "oracle.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        from oracle import oracle_py
        oracle_py.build()
        subprocess.check_call(["/usr/bin/gcc", "-O1", "-g", "-fPIC", "-shared", "-std=gnu11", "-Wall", "-Wno-unused-parameter", "-ffp-contract=off", "-o", so, src,
                               "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle"), "-L" + os.path.join(ROOT, "oracle"), "-loracle",
                               "-Wl,-rpath,$ORIGIN/../../oracle", "-lm"])
    return so
