"""Build oracle/_ref/libbitblas_ref.so from the reference's OWN sources where they lie.

TEST INFRASTRUCTURE ONLY.  Compiles oracle/ref_shim.cu, which #includes
/root/reference/testing/cpp/lop3_type_conversion/fast_decoding.hpp (never copied into the repo), with
plain nvcc (the reference's CMake/gtest build is not used).  Output goes only to oracle/_ref/ which is
git-ignored but NOT gpurun-ignored, so the .so travels to the GPU box.  If /root/reference is absent
(i.e. on the GPU box) this is a no-op and the prebuilt file is used.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_INC = "/root/reference/testing/cpp/lop3_type_conversion"
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "libbitblas_ref.so")


def build(force: bool = False) -> str | None:
    if not os.path.isdir(REF_INC):
        return OUT if os.path.exists(OUT) else None
    os.makedirs(OUT_DIR, exist_ok=True)
    src = os.path.join(HERE, "ref_shim.cu")
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(src):
        return OUT
    cmd = ["nvcc", "-O2", "-shared", "-Xcompiler", "-fPIC", "-w",
           "-gencode", "arch=compute_100a,code=sm_100a", "-I", REF_INC, src, "-o", OUT]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
