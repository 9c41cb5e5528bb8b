"""ctypes binding of tests/warp_emu/libstep_kernel_emu.so -- TEST HARNESS ONLY.

The library is ssl-sim_b200/csrc/step_kernel.cu compiled *for the host* against a stand-in cuda_runtime.h: the
kernel's own source, its 32 lanes run as coroutines (tests/warp_emu/).  It lets the CPU-only suite check the
kernel's logic and arithmetic bit-for-bit against the oracle where no GPU exists.  It is not a fallback: nothing
in ssl-sim_b200/ can load it, and the product still fails loudly without a CUDA device.
"""
import ctypes as C
import os
import subprocess

import numpy as np

import oracle_binding as ob

_HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(_HERE, "warp_emu")
_LIB = None


def build():
    subprocess.run(["make", "-s", "-C", EMU_DIR], check=True)


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(os.path.join(EMU_DIR, "libstep_kernel_emu.so"))
        vp = C.c_void_p
        L.emu_step_worlds.argtypes = [C.POINTER(ob.Field), C.POINTER(ob.Params), C.c_int, C.c_int, vp, vp, vp, vp,
                                      C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_longlong, C.c_int, vp, vp]
        L.emu_step_worlds.restype = C.c_int
        _LIB = L
    return _LIB


class EmuWorlds(ob.Worlds):
    """Same initial state / command generator as the oracle's Worlds (those are tested against the CUDA init and
    gen_commands kernels on the GPU); only step() runs the emulated kernel."""

    def step(self, n_steps=1, substeps=2, h=ob.H, lanes=0, threads=8, cmd_seq=None, steps_per_period=0, prof=None):
        cmd = self.cmd if cmd_seq is None else cmd_seq
        stride = self.W * self.R if cmd_seq is not None else 0
        rc = lib().emu_step_worlds(C.byref(self.field), C.byref(self.params), self.R, self.W, ob._ptr(self.pos),
                                   ob._ptr(self.vel), ob._ptr(self.aux), ob._ptr(cmd), n_steps, substeps,
                                   C.c_float(float(h)), lanes, steps_per_period, stride, threads, ob._ptr(prof),
                                   ob._ptr(self.warm))
        assert rc == 0
