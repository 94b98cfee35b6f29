"""ctypes driver of tests/hostsim/libhostsim.so (the device headers compiled for the host; test infrastructure only)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / "pokegym_b200" / "csrc"
STATE_BYTES = 142_610


def build(force: bool = False) -> Path:
    out = HERE / "libhostsim.so"
    srcs = [HERE / "hostsim.cpp"] + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [ROOT / "include" / "gbenv.h"]
    if not force and out.exists() and all(s.stat().st_mtime <= out.stat().st_mtime for s in srcs):
        return out
    subprocess.run(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", str(out), str(HERE / "hostsim.cpp")], check=True)
    return out


class HostSim:
    def __init__(self, n: int, rom: bytes, simt: bool = True):
        self.dll = C.CDLL(str(build()))
        self.dll.hs_create.restype = C.c_void_p
        self.dll.hs_create.argtypes = [C.c_int, C.c_char_p, C.c_size_t]
        for name, args in (("hs_destroy", [C.c_void_p]), ("hs_load_blob", [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t]),
                           ("hs_save_blob", [C.c_void_p, C.c_int, C.c_void_p]), ("hs_run", [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
                           ("hs_counters", [C.c_void_p, C.c_void_p]), ("hs_core_extra", [C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 3)):
            getattr(self.dll, name).argtypes = args
        self.n = n
        self.h = self.dll.hs_create(n, rom, len(rom))
        self.dll.hs_set_simt.argtypes = [C.c_void_p, C.c_int]
        self.dll.hs_set_simt(self.h, 1 if simt else 0)  # which build of the fast loop: k_run_frames (SIMT) or k_run_frames_1
        self.dll.hs_set_defer.argtypes = [C.c_void_p, C.c_int]

    def set_defer(self, on: bool):
        """deferred PPU on (the product's default) or off (lines drawn inside the frame loop)"""
        self.dll.hs_set_defer(self.h, 1 if on else 0)

    def close(self):
        if self.h:
            self.dll.hs_destroy(self.h)
            self.h = None

    def load_blob(self, env: int, blob: bytes):
        rc = self.dll.hs_load_blob(self.h, env, blob, len(blob))
        assert rc == 0, rc

    def save_state(self, env: int) -> bytes:
        out = np.empty(STATE_BYTES, dtype=np.uint8)
        self.dll.hs_save_blob(self.h, env, out.ctypes.data)
        return out.tobytes()

    def tick(self, n_frames: int, render: bool = True):
        self.dll.hs_run(self.h, None, n_frames, 1 if render else 0)

    def run_action(self, actions: np.ndarray, frame_skip: int = 24):
        a = np.ascontiguousarray(actions, dtype=np.uint8)
        self.dll.hs_run(self.h, a.ctypes.data, frame_skip, 2)

    def counters(self):
        c = (C.c_ulonglong * 4)()
        self.dll.hs_counters(self.h, c)
        return list(c)

    def core_extra(self, env: int):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self.dll.hs_core_extra(self.h, env, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value
