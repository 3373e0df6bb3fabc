"""Kernel A/B experiments: build a variant of libb200dae.so whose Robertson translation unit is
compiled with extra -D flags, next to the product library.

    python scripts/build_variant.py NAME -DB200DAE_FD_BATCH=2 [-DB200DAE_REG_MINBLOCKS=7 ...]

writes scipy_dae_b200/variants/libb200dae_NAME.so (git-ignored, travels to the GPU box); select it
with B200DAE_LIB=<path> (scipy_dae_b200/_lib.py).  Prints the ptxas resource lines of the Radau
kernels so spills show up before any GPU time is spent."""
from __future__ import annotations

import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scipy_dae_b200 import build as B  # noqa: E402


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    with_stages = "--stages" in flags           # also build the s = 1, 5, 7 kernels (slower)
    flags = [f for f in flags if f != "--stages"]
    if not os.path.exists(os.path.join(B.OBJ_DIR, "b200dae.o")):
        B.build()                               # b200dae.o (C-ABI + registry) comes from the product build
    out_dir = os.path.join(B.HERE, "variants")
    os.makedirs(out_dir, exist_ok=True)
    src = os.path.join(B.CSRC, "problem_robertson.cu")
    obj = os.path.join(out_dir, f"problem_robertson_{name}.o")
    cmd = [B.find_nvcc(), *B.NVCC_FLAGS, *flags, "-Xptxas", "-v", "-c", "-o", obj, src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.exit(r.stdout + r.stderr)
    # ptxas -v: "Compiling entry function '<mangled>'" then "Used N registers, ..." lines
    cur = None
    for line in r.stderr.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = m.group(1)
        elif cur and "radau_kernel" in cur and ("Used" in line or "spill" in line):
            print(cur[:90], "|", line.split(":", 1)[-1].strip())
    stubs = os.path.join(out_dir, "variant_stubs.o")
    stub_src = os.path.join(ROOT, "scripts", "variant_stubs.cu")
    if not os.path.exists(stubs) or os.path.getmtime(stubs) < os.path.getmtime(stub_src):
        subprocess.run([B.find_nvcc(), *B.NVCC_FLAGS, "-c", "-o", stubs, stub_src], check=True)
    # Robertson s = 3 only: every other registry entry is an empty stub (3 MB instead of 33 MB)
    # the C-ABI unit holds the fastmath self-test kernel, so it is compiled with the same flags
    abi = os.path.join(out_dir, f"b200dae_{name}.o")
    subprocess.run([B.find_nvcc(), *B.NVCC_FLAGS, *flags, "-c", "-o", abi,
                    os.path.join(B.CSRC, "b200dae.cu")], check=True)
    objs = [abi, stubs, obj, os.path.join(B.OBJ_DIR, "rt_registry.o")]   # run-time registry: from the product build
    if with_stages:
        st = os.path.join(out_dir, f"problem_robertson_stages_{name}.o")
        subprocess.run([B.find_nvcc(), *B.NVCC_FLAGS, *flags, "-c", "-o", st,
                        os.path.join(B.CSRC, "problem_robertson_stages.cu")], check=True)
        objs.append(st)
    else:
        nost = os.path.join(out_dir, "variant_nostages.o")
        subprocess.run([B.find_nvcc(), *B.NVCC_FLAGS, "-DB200DAE_STUB_STAGES", "-c", "-o", nost,
                        os.path.join(ROOT, "scripts", "variant_stubs.cu")], check=True)
        objs.append(nost)
    lib = os.path.join(out_dir, f"libb200dae_{name}.so")
    subprocess.run([B.find_nvcc(), "-shared", "-o", lib, *objs], check=True)
    print(lib)


if __name__ == "__main__":
    main()
