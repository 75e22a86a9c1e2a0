"""TEST INFRASTRUCTURE: the product's CUDA translation unit compiled by g++ against the SIMT emulator
(tests/simt/) into tests/_build/libr433b_emu.so, so that the kernels' own source runs on the CPU -- one
fibre per CUDA thread, warp collectives as rendezvous points -- under the parity tests of the GPU suite.
This is not a CPU implementation and never part of the product: rtl_433_b200/lib.py loads
csrc/libr433b.so (nvcc, sm_100a) and nothing else unless a test points it here."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EMU_SO = os.path.join(HERE, "_build", "libr433b_emu.so")


def build(extra=()):
    csrc = os.path.join(ROOT, "rtl_433_b200", "csrc")
    deps = [os.path.join(csrc, n) for n in os.listdir(csrc) if n.endswith((".cu", ".cuh", ".hpp"))]
    deps += [os.path.join(HERE, "simt", n) for n in os.listdir(os.path.join(HERE, "simt"))]
    deps += [os.path.join(ROOT, "include", n) for n in ("r433b.h", "r433b_abi.h")]
    if not extra and os.path.exists(EMU_SO) and all(os.path.getmtime(EMU_SO) >= os.path.getmtime(d) for d in deps):
        return EMU_SO
    os.makedirs(os.path.dirname(EMU_SO), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-ffp-contract=off", "-DR433B_SIMT_EMU",
                           "-I" + os.path.join(HERE, "simt"), *extra, "-x", "c++", os.path.join(csrc, "r433b_api.cu"),
                           "-o", EMU_SO])
    return EMU_SO


def use():
    """Point rtl_433_b200.lib at the emulated library (for the rest of this process)."""
    from rtl_433_b200 import lib
    path = build()
    if lib.LIB_PATH != path:
        lib.LIB_PATH = path
        lib._lib = None
    return path
