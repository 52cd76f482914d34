"""oracle/pyoracle.py -- TEST INFRASTRUCTURE. ctypes binding of oracle/libkmc_oracle.so (the CPU
restatement of main.cpp:461-2202). Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libkmc_oracle.so")


class Params(C.Structure):
    _fields_ = [("box", C.c_double * 3), ("dt", C.c_double), ("pai", C.c_double),
                ("rA", C.c_double), ("DA", C.c_double), ("DrotA", C.c_double),
                ("rB", C.c_double), ("DB", C.c_double), ("DrotB", C.c_double),
                ("mono_cis_on", C.c_double), ("mono_cis_off", C.c_double),
                ("cis_D", C.c_double), ("cis_Drot", C.c_double), ("cis_on", C.c_double), ("cis_off", C.c_double),
                ("bond_D", C.c_double), ("bond_Drot", C.c_double), ("on", C.c_double), ("off", C.c_double),
                ("bond_dist_cut", C.c_double), ("thetapd_cut", C.c_double), ("thetaot_cut", C.c_double),
                ("cis_thetaot_cut", C.c_double), ("cis_dist_cut", C.c_double),
                ("n_receptor", C.c_int32), ("n_ligand", C.c_int32), ("stream_mode", C.c_int32),
                ("use_grid", C.c_int32), ("seed", C.c_uint64), ("rand2_state", C.c_uint64),
                ("rand_state", C.c_uint64), ("order_mode", C.c_int32), ("pad_", C.c_int32),
                ("order_x0", C.c_double), ("order_y0", C.c_double), ("order_inv_edge", C.c_double)]


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("kmc_oracle.cpp", "kmc_oracle.h", "philox.h")]
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(s) for s in srcs):
        subprocess.run(["make", "-C", HERE, "libkmc_oracle.so"], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.kmco_create.restype = C.c_void_p
        _lib.kmco_create.argtypes = [C.POINTER(Params)]
        _lib.kmco_destroy.argtypes = [C.c_void_p]
        _lib.kmco_init_reference.argtypes = [C.c_void_p]
        _lib.kmco_step.argtypes = [C.c_void_p, C.c_int64]
        _lib.kmco_get_counts.restype = C.c_double
        _lib.kmco_get_counts.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        _lib.kmco_set_state.argtypes = [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int64, C.c_int32]
        _lib.kmco_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        _lib.kmco_get_results.restype = C.c_int64
        _lib.kmco_get_results.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        _lib.kmco_get_accept.argtypes = [C.c_void_p, C.c_void_p]
        _lib.kmco_get_events.argtypes = [C.c_void_p, C.c_void_p]
        _lib.kmco_default_params.argtypes = [C.POINTER(Params)]
    return _lib


def default_params(**kw):
    p = Params()
    lib().kmco_default_params(C.byref(p))
    for k, v in kw.items():
        if k == "box":
            p.box[0], p.box[1], p.box[2] = v
        else:
            setattr(p, k, v)
    return p


class Oracle:
    def __init__(self, params):
        self.p = params
        self.n = params.n_receptor + params.n_ligand
        self.h = lib().kmco_create(C.byref(params))

    def __del__(self):
        if getattr(self, "h", None):
            lib().kmco_destroy(self.h)
            self.h = None

    def init_reference(self):
        lib().kmco_init_reference(self.h)

    def step(self, n=1):
        lib().kmco_step(self.h, n)

    def counts(self):
        c = np.zeros(8, dtype=np.int32)
        sd = C.c_int64(); n2 = C.c_uint64(); nr = C.c_uint64()
        cs = lib().kmco_get_counts(self.h, c.ctypes.data, C.addressof(sd), C.addressof(n2), C.addressof(nr))
        return dict(bond_num=int(c[0]), bond_num_rl=int(c[1]), bond_num_cis=int(c[2]), bond_num_mono_cis=int(c[3]),
                    max_complex=int(c[4]), tot_cluster_num=int(c[5]), tot_proteins_in_cluster=int(c[6]),
                    cluster_size=cs, step=sd.value, rand2_draws=n2.value, rand_draws=nr.value)

    def get_state(self):
        """Returns R[N+1,5,5,3], status[N+1,5], res_nei[N+1,7] (reference array shapes)."""
        n = self.n
        X = np.zeros((n + 1, 5, 5)); Y = np.zeros_like(X); Z = np.zeros_like(X)
        st = np.zeros((n + 1, 5), dtype=np.int32); rn = np.zeros((n + 1, 7), dtype=np.int32)
        lib().kmco_get_state(self.h, X.ctypes.data, Y.ctypes.data, Z.ctypes.data, st.ctypes.data, rn.ctypes.data)
        return np.stack([X, Y, Z], axis=-1), st, rn

    def set_state(self, R, status, res_nei, step_done=0, max_complex=0):
        X = np.ascontiguousarray(R[..., 0]); Y = np.ascontiguousarray(R[..., 1]); Z = np.ascontiguousarray(R[..., 2])
        st = np.ascontiguousarray(status, dtype=np.int32); rn = np.ascontiguousarray(res_nei, dtype=np.int32)
        lib().kmco_set_state(self.h, X.ctypes.data, Y.ctypes.data, Z.ctypes.data, st.ctypes.data, rn.ctypes.data,
                             step_done, max_complex)

    def results(self):
        nb = self.p.n_ligand
        rl = np.zeros(nb, dtype=np.int32); mem = np.zeros(self.n + 1, dtype=np.int32)
        tot = lib().kmco_get_results(self.h, rl.ctypes.data, mem.ctypes.data, mem.size)
        rows, o = [], 0
        for l in range(nb):
            rows.append(mem[o:o + rl[l]].tolist()); o += rl[l]
        assert o == tot
        return rows

    def accepted(self):
        a = np.zeros(self.n + 1, dtype=np.int32)
        lib().kmco_get_accept(self.h, a.ctypes.data)
        return a

    def events(self):
        e = np.zeros(16, dtype=np.int64)
        lib().kmco_get_events(self.h, e.ctypes.data)
        return e
