"""kmc_b200 -- thin ctypes binding of libkmc_b200.so (C ABI in include/kmc_b200.h).

This is plumbing for tests and bench.py: the product is the shared library. There is no Python or CPU
implementation of the sweep here; if the library or a B200 is missing every call fails loudly.
"""
import ctypes as C
import importlib.util
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("KMC_LIB") or os.path.join(PKG_DIR, "libkmc_b200.so")

MODE_REPLAY, MODE_PRODUCTION = 0, 1


class Params(C.Structure):
    _fields_ = [("box", C.c_double * 3), ("dt", C.c_double), ("pai", C.c_double),
                ("rA", C.c_double), ("DA", C.c_double), ("DrotA", C.c_double),
                ("rB", C.c_double), ("DB", C.c_double), ("DrotB", C.c_double),
                ("mono_cis_on", C.c_double), ("mono_cis_off", C.c_double),
                ("cis_D", C.c_double), ("cis_Drot", C.c_double), ("cis_on", C.c_double), ("cis_off", C.c_double),
                ("bond_D", C.c_double), ("bond_Drot", C.c_double), ("on", C.c_double), ("off", C.c_double),
                ("bond_dist_cut", C.c_double), ("thetapd_cut", C.c_double), ("thetaot_cut", C.c_double),
                ("cis_thetaot_cut", C.c_double), ("cis_dist_cut", C.c_double),
                ("n_receptor", C.c_int32), ("n_ligand", C.c_int32), ("n_replicas", C.c_int32), ("mode", C.c_int32),
                ("seed", C.c_uint64), ("cell_edge", C.c_double), ("device", C.c_int32), ("min_image", C.c_int32)]


class Series(C.Structure):
    _fields_ = [("step", C.c_int64), ("bond_num_rl", C.c_int32), ("bond_num_mono_cis", C.c_int32),
                ("bond_num_cis", C.c_int32), ("bond_num", C.c_int32), ("max_complex", C.c_int32),
                ("n_complexes", C.c_int32), ("n_in_complexes", C.c_int32), ("reserved", C.c_int32),
                ("cluster_size", C.c_double)]


EXPORTS = ["kmc_abi_version", "kmc_default_params", "kmc_create", "kmc_destroy", "kmc_last_error", "kmc_init_random",
           "kmc_set_state", "kmc_get_state", "kmc_get_packed", "kmc_set_packed", "kmc_step", "kmc_sync", "kmc_get_series",
           "kmc_get_complexes", "kmc_get_complex_labels", "kmc_get_oligomer_hist", "kmc_get_accept", "kmc_get_events", "kmc_get_live_counts", "kmc_get_step_path", "kmc_alignment_window", "kmc_get_packed_async", "kmc_snapshot_wait", "kmc_write_bond_dat",
           "kmc_write_cluster_log", "kmc_run", "kmc_step_timed", "kmc_profile", "kmc_profile_get", "kmc_timeline_print", "kmc_format_bond_dat", "kmc_format_cluster_log", "kmc_get_grid", "kmc_strip_configure", "kmc_strip_load_global",
           "kmc_strip_begin_refresh", "kmc_strip_message", "kmc_strip_rebuild", "kmc_strip_halo_width", "kmc_strip_unique_id",
           "kmc_strip_comm_init", "kmc_strip_refresh", "kmc_strip_refresh_local", "kmc_strip_get_series", "kmc_strip_get_oligomer_hist",
           "kmc_strip_get_records", "kmc_strip_load_records", "kmc_strip_init_random", "kmc_generate_packed", "kmc_gro_append_arrays", "kmc_checkpoint_write_arrays",
           "kmc_checkpoint_read_arrays", "kmc_parameter_log_write", "kmc_write_gro", "kmc_write_checkpoint", "kmc_read_checkpoint",
           "kmc_write_checkpoint_bin", "kmc_read_checkpoint_bin"]


class KmcError(RuntimeError):
    pass


def build(force=False, verbose=False):
    spec = importlib.util.spec_from_file_location("kmc_b200_build", os.path.join(PKG_DIR, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(force=force, verbose=verbose)


_lib = None


def lib():
    """Loads the CUDA library; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KmcError("libkmc_b200.so is missing: run __graft_entry__.build() (nvcc, sm_100a); there is no CPU path")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
        L.kmc_abi_version.restype = C.c_int
        L.kmc_default_params.argtypes = [C.POINTER(Params)]
        L.kmc_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
        L.kmc_destroy.argtypes = [vp]
        L.kmc_last_error.restype = C.c_char_p
        L.kmc_last_error.argtypes = [vp]
        L.kmc_init_random.argtypes = [vp, u64, i32]
        L.kmc_set_state.argtypes = [vp, i32, vp, vp, vp, vp, vp, i64, i32]
        L.kmc_get_state.argtypes = [vp, i32, vp, vp, vp, vp, vp]
        L.kmc_get_packed.argtypes = [vp, vp, vp, vp, vp, vp]
        L.kmc_set_packed.argtypes = [vp, vp, vp, vp, vp, vp, i64]
        L.kmc_get_packed_async.argtypes = [vp, vp, vp, vp, vp, vp]
        L.kmc_snapshot_wait.argtypes = [vp]
        L.kmc_step.argtypes = [vp, i64]
        L.kmc_sync.argtypes = [vp]
        L.kmc_get_series.argtypes = [vp, i32, C.POINTER(Series)]
        L.kmc_get_complexes.restype = i64
        L.kmc_get_complexes.argtypes = [vp, i32, vp, vp, i64]
        L.kmc_get_complex_labels.argtypes = [vp, i32, vp]
        L.kmc_get_oligomer_hist.argtypes = [vp, i32, vp, i32]
        L.kmc_get_accept.argtypes = [vp, i32, vp]
        L.kmc_get_events.argtypes = [vp, vp]
        L.kmc_get_live_counts.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
        L.kmc_get_step_path.argtypes = [vp]
        L.kmc_alignment_window.argtypes = [C.c_double, C.POINTER(C.c_double)]
        L.kmc_strip_configure.argtypes = [vp, i32, i32, C.c_double]
        L.kmc_strip_load_global.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, i64]
        L.kmc_strip_begin_refresh.argtypes = [vp]
        L.kmc_strip_message.restype = i64
        L.kmc_strip_message.argtypes = [vp, i32, C.POINTER(vp)]
        L.kmc_strip_rebuild.argtypes = [vp, vp, i64, vp, i64]
        L.kmc_strip_halo_width.restype = C.c_double
        L.kmc_strip_halo_width.argtypes = [C.POINTER(Params), i32, C.c_double]
        L.kmc_strip_unique_id.argtypes = [vp]
        L.kmc_strip_comm_init.argtypes = [vp, vp, i32]
        L.kmc_strip_refresh.argtypes = [vp]
        L.kmc_strip_refresh_local.argtypes = [C.POINTER(vp), i32, i32]
        L.kmc_strip_get_series.argtypes = [vp, i32, C.POINTER(Series)]
        L.kmc_strip_get_oligomer_hist.argtypes = [vp, i32, vp, i32]
        L.kmc_strip_get_records.argtypes = [vp, i32, vp, i64, C.POINTER(i64), C.POINTER(i64)]
        L.kmc_strip_load_records.argtypes = [vp, vp, i64, i64, i64]
        L.kmc_strip_init_random.argtypes = [vp, i32, i32, u64, i32]
        L.kmc_gro_append_arrays.argtypes = [C.c_char_p, i32, i32, vp, vp, vp, C.c_double, i64, vp]
        L.kmc_checkpoint_write_arrays.argtypes = [C.c_char_p, i32, i32, vp, vp, vp, vp, vp, vp]
        L.kmc_checkpoint_read_arrays.argtypes = [C.c_char_p, i32, i32, vp, vp, vp, vp, vp, vp]
        L.kmc_parameter_log_write.argtypes = [C.POINTER(Params), C.c_char_p]
        L.kmc_write_gro.argtypes = [vp, i32, C.c_char_p]
        L.kmc_write_checkpoint.argtypes = [vp, i32, C.c_char_p]
        L.kmc_read_checkpoint.argtypes = [vp, i32, C.c_char_p]
        L.kmc_write_checkpoint_bin.argtypes = [vp, C.c_char_p]
        L.kmc_read_checkpoint_bin.argtypes = [vp, C.c_char_p]
        L.kmc_generate_packed.argtypes = [C.POINTER(Params), u64, i32, vp, vp]
        L.kmc_get_grid.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i32), C.POINTER(i32)]
        L.kmc_write_bond_dat.argtypes = [vp, i32, C.c_char_p]
        L.kmc_write_cluster_log.argtypes = [vp, i32, C.c_char_p]
        L.kmc_run.argtypes = [vp, i64, i32, C.c_char_p]
        L.kmc_format_bond_dat.argtypes = [C.c_double, C.POINTER(Series), C.c_char_p, i32]
        L.kmc_format_cluster_log.restype = i64
        L.kmc_format_cluster_log.argtypes = [C.c_double, i64, i32, vp, vp, C.c_char_p, i64]
        L.kmc_step_timed.argtypes = [vp, i64, C.POINTER(C.c_double)]
        L.kmc_profile.argtypes = [vp, i32]
        L.kmc_timeline_print.argtypes = [vp]
        L.kmc_profile_get.argtypes = [vp, i32, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(i64)]
        _lib = L
    return _lib


def default_params(**kw):
    p = Params()
    lib().kmc_default_params(C.byref(p))
    for k, v in kw.items():
        if k == "box":
            p.box[0], p.box[1], p.box[2] = v
        else:
            setattr(p, k, v)
    return p


def generate_packed(params, seed=1, sort_cells=True):
    """random non-overlapping start state (the generator of kmc_init_random) as host arrays; needs no GPU"""
    r = params.n_replicas
    rec = np.zeros((r * params.n_receptor, 6)); lig = np.zeros((r * params.n_ligand, 24))
    rc = lib().kmc_generate_packed(C.byref(params), seed, int(sort_cells), rec.ctypes.data, lig.ctypes.data)
    if rc != 0:
        raise KmcError("kmc_generate_packed failed (%d): box too dense?" % rc)
    return rec, lig


def strip_unique_id():
    """128 bytes identifying a new NCCL communicator (rank 0 calls it and hands the bytes to the other ranks)"""
    buf = C.create_string_buffer(128)
    rc = lib().kmc_strip_unique_id(buf)
    if rc != 0:
        raise KmcError("kmc_strip_unique_id failed (%d): %s" % (rc, lib().kmc_last_error(None).decode()))
    return buf.raw


def strip_halo_width(params, refresh_every, complex_extent=400.0):
    return lib().kmc_strip_halo_width(C.byref(params), refresh_every, complex_extent)


def strip_refresh_local(handles, refresh_every=0):
    """one refresh between K handles of this process (logical ranks 0..K-1 on one GPU)"""
    arr = (C.c_void_p * len(handles))(*[k.h for k in handles])
    rc = lib().kmc_strip_refresh_local(arr, len(handles), refresh_every)
    if rc < 0:
        raise KmcError("kmc_strip_refresh_local failed (%d): %s" % (rc, "; ".join(lib().kmc_last_error(k.h).decode() for k in handles)))


def alignment_window(length):
    """[lo, hi] of squared distances q with fabs(sqrt(q) - length) < 1e-8 (host arithmetic; None if not available)"""
    w = (C.c_double * 2)()
    return (w[0], w[1]) if lib().kmc_alignment_window(length, w) == 0 else None


def scaled_box(n_total, z=1000.0):
    """Box edge that keeps the reference's default densities (main.cpp:43-57): L = 5773*sqrt(N/200)."""
    L = 5773.0 * (n_total / 200.0) ** 0.5
    return (L, L, z)


def _xyz(R):
    return [np.ascontiguousarray(R[..., c], dtype=np.float64) for c in range(3)]


def gro_append(path, na, nb, R, dt, step, box):
    """append one test.gro frame (main.cpp:2258-2287) for reference-shaped coordinates R[N+1,5,5,3]; host only"""
    X, Y, Z = _xyz(R); bx = np.array(box, dtype=np.float64)
    rc = lib().kmc_gro_append_arrays(os.fsencode(path), na, nb, X.ctypes.data, Y.ctypes.data, Z.ctypes.data, dt, step, bx.ctypes.data)
    if rc:
        raise KmcError("kmc_gro_append_arrays failed: %d" % rc)


def checkpoint_write(path, na, nb, R, status, res_nei, counters):
    """position.cpt in the reference's format (main.cpp:2206-2244); counters = (bond_num, rl, cis, mono_cis, max_complex, step)"""
    X, Y, Z = _xyz(R); st = np.ascontiguousarray(status, dtype=np.int32); rn = np.ascontiguousarray(res_nei, dtype=np.int32)
    c = np.array(counters, dtype=np.int64)
    rc = lib().kmc_checkpoint_write_arrays(os.fsencode(path), na, nb, X.ctypes.data, Y.ctypes.data, Z.ctypes.data, st.ctypes.data, rn.ctypes.data, c.ctypes.data)
    if rc:
        raise KmcError("kmc_checkpoint_write_arrays failed: %d" % rc)


def checkpoint_read(path, na, nb):
    """-> R[N+1,5,5,3], status, res_nei, counters[6]   (main.cpp:226-268)"""
    n = na + nb
    X = np.zeros((n + 1, 5, 5)); Y = np.zeros_like(X); Z = np.zeros_like(X)
    st = np.zeros((n + 1, 5), dtype=np.int32); rn = np.zeros((n + 1, 7), dtype=np.int32); c = np.zeros(6, dtype=np.int64)
    rc = lib().kmc_checkpoint_read_arrays(os.fsencode(path), na, nb, X.ctypes.data, Y.ctypes.data, Z.ctypes.data, st.ctypes.data, rn.ctypes.data, c.ctypes.data)
    if rc:
        raise KmcError("kmc_checkpoint_read_arrays failed: %d" % rc)
    return np.stack([X, Y, Z], axis=-1), st, rn, c


def parameter_log_write(params, path):
    rc = lib().kmc_parameter_log_write(C.byref(params), os.fsencode(path))
    if rc:
        raise KmcError("kmc_parameter_log_write failed: %d" % rc)


def format_bond_dat(dt, step, bond_num_rl, bond_num_mono_cis, bond_num_cis, bond_num, cluster_size, max_complex):
    """one bond.dat line exactly as main.cpp:2251 prints it (host only, no GPU)"""
    s = Series(step=step, bond_num_rl=bond_num_rl, bond_num_mono_cis=bond_num_mono_cis, bond_num_cis=bond_num_cis,
               bond_num=bond_num, max_complex=max_complex, cluster_size=cluster_size)
    buf = C.create_string_buffer(256)
    n = lib().kmc_format_bond_dat(dt, C.byref(s), buf, 256)
    if n < 0:
        raise KmcError("kmc_format_bond_dat failed: %d" % n)
    return buf.value.decode()


def format_cluster_log(dt, step, rows):
    """one cluster.log frame exactly as main.cpp:2293-2301 prints it (host only, no GPU); rows = list of member lists per ligand"""
    rl = np.array([len(r) for r in rows], dtype=np.int32)
    mem = np.array([m for r in rows for m in r] + [0], dtype=np.int32)
    cap = 64 + 16 * mem.size + rl.size
    buf = C.create_string_buffer(cap)
    n = lib().kmc_format_cluster_log(dt, step, rl.size, rl.ctypes.data, mem.ctypes.data, buf, cap)
    if n < 0:
        raise KmcError("kmc_format_cluster_log failed: %d" % n)
    return buf.value.decode()


class Kmc:
    """One kmc_handle. Mirrors the reference's use: set parameters, (re)start from a state, step, read outputs."""

    def __init__(self, params):
        self.p = params
        self.na, self.nb = params.n_receptor, params.n_ligand
        self.n = self.na + self.nb
        self.h = C.c_void_p()
        rc = lib().kmc_create(C.byref(params), C.byref(self.h))
        if rc != 0:
            raise KmcError("kmc_create failed (%d): %s" % (rc, lib().kmc_last_error(None).decode()))

    def close(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.kmc_destroy(self.h)
        self.h = None

    __del__ = close

    def _ck(self, rc):
        if rc < 0:
            raise KmcError("kmc error %d: %s" % (rc, lib().kmc_last_error(self.h).decode()))
        return rc

    def init_random(self, seed=1, sort_cells=False):
        self._ck(lib().kmc_init_random(self.h, seed, int(sort_cells)))

    def set_state(self, R, status, res_nei, replica=0, step_done=0, max_complex=0):
        X = np.ascontiguousarray(R[..., 0], dtype=np.float64); Y = np.ascontiguousarray(R[..., 1], dtype=np.float64)
        Z = np.ascontiguousarray(R[..., 2], dtype=np.float64)
        st = np.ascontiguousarray(status, dtype=np.int32); rn = np.ascontiguousarray(res_nei, dtype=np.int32)
        assert X.shape == (self.n + 1, 5, 5) and st.shape == (self.n + 1, 5) and rn.shape == (self.n + 1, 7)
        self._ck(lib().kmc_set_state(self.h, replica, X.ctypes.data, Y.ctypes.data, Z.ctypes.data, st.ctypes.data,
                                     rn.ctypes.data, step_done, max_complex))

    def get_state(self, replica=0):
        n = self.n
        X = np.zeros((n + 1, 5, 5)); Y = np.zeros_like(X); Z = np.zeros_like(X)
        st = np.zeros((n + 1, 5), dtype=np.int32); rn = np.zeros((n + 1, 7), dtype=np.int32)
        self._ck(lib().kmc_get_state(self.h, replica, X.ctypes.data, Y.ctypes.data, Z.ctypes.data, st.ctypes.data, rn.ctypes.data))
        return np.stack([X, Y, Z], axis=-1), st, rn

    def get_packed(self, out=None):
        """-> rec[n,6], lig[n,24], rec_lig, rec_site, rec_cis. `out` = the same five arrays preallocated (e.g. pinned host memory)."""
        r = self.p.n_replicas
        if out is not None:
            rec, lig, rl, rs, rc = out
        else:
            rec = np.zeros((r * self.na, 6)); lig = np.zeros((r * self.nb, 24))
            rl = np.zeros(r * self.na, dtype=np.int32); rs = np.zeros_like(rl); rc = np.zeros_like(rl)
        self._ck(lib().kmc_get_packed(self.h, rec.ctypes.data, lig.ctypes.data, rl.ctypes.data, rs.ctypes.data, rc.ctypes.data))
        return rec, lig, rl, rs, rc

    def get_packed_async(self, out):
        """starts the transfer of the state as it stands after the steps enqueued so far into `out` (five preallocated arrays in PINNED
        host memory); stepping may continue at once; snapshot_wait() completes the arrays"""
        rec, lig, rl, rs, rc = out
        self._ck(lib().kmc_get_packed_async(self.h, rec.ctypes.data, lig.ctypes.data, rl.ctypes.data, rs.ctypes.data, rc.ctypes.data))
        return out

    def snapshot_wait(self):
        self._ck(lib().kmc_snapshot_wait(self.h))

    def set_packed(self, rec, lig, rl=None, rs=None, rc=None, step_done=0):
        rec = np.ascontiguousarray(rec, dtype=np.float64); lig = np.ascontiguousarray(lig, dtype=np.float64)
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.int32) for a in (rl, rs, rc)]
        ptr = [None if a is None else a.ctypes.data for a in arrs]
        self._ck(lib().kmc_set_packed(self.h, rec.ctypes.data, lig.ctypes.data, ptr[0], ptr[1], ptr[2], step_done))

    def step(self, n=1):
        self._ck(lib().kmc_step(self.h, n))

    def step_timed(self, n=1):
        """n steps; returns device milliseconds (CUDA events on the handle's stream)."""
        ms = C.c_double()
        self._ck(lib().kmc_step_timed(self.h, n, C.byref(ms)))
        return ms.value

    def profile(self, enable=True):
        self._ck(lib().kmc_profile(self.h, int(enable)))

    def profile_get(self):
        """{kernel name: (total ms, launches)} since profile(True)"""
        out, i = {}, 0
        while True:
            name = C.c_char_p(); ms = C.c_double(); cnt = C.c_int64()
            rc = self._ck(lib().kmc_profile_get(self.h, i, C.byref(name), C.byref(ms), C.byref(cnt)))
            if rc == 1:
                return out
            out[name.value.decode()] = (ms.value, cnt.value)
            i += 1

    def sync(self):
        self._ck(lib().kmc_sync(self.h))

    def series(self, replica=0):
        s = Series()
        self._ck(lib().kmc_get_series(self.h, replica, C.byref(s)))
        return {k: getattr(s, k) for k, _ in Series._fields_ if k != "reserved"}

    def complexes(self, replica=0):
        rl = np.zeros(self.nb, dtype=np.int32); mem = np.zeros(self.n + 1, dtype=np.int32)
        tot = self._ck(lib().kmc_get_complexes(self.h, replica, rl.ctypes.data, mem.ctypes.data, mem.size))
        rows, o = [], 0
        for l in range(self.nb):
            rows.append(mem[o:o + rl[l]].tolist()); o += rl[l]
        assert o == tot
        return rows

    def complex_labels(self, replica=0):
        """root[i] (1-based): reference id of the head of the unit molecule i belongs to after the last step"""
        a = np.zeros(self.n + 1, dtype=np.int32)
        self._ck(lib().kmc_get_complex_labels(self.h, replica, a.ctypes.data))
        return a

    def oligomer_hist(self, replica=-1, nbins=64):
        hist = np.zeros(nbins, dtype=np.int64)
        self._ck(lib().kmc_get_oligomer_hist(self.h, replica, hist.ctypes.data, nbins))
        return hist

    def accepted(self, replica=0):
        a = np.zeros(self.n + 1, dtype=np.int32)
        self._ck(lib().kmc_get_accept(self.h, replica, a.ctypes.data))
        return a

    # ---- strip decomposition (csrc/kmc_strips.cu) ----
    def strip_configure(self, rank, nranks, halo_width):
        self._ck(lib().kmc_strip_configure(self.h, rank, nranks, float(halo_width)))

    def strip_load_global(self, rec, lig, rl=None, rs=None, rc=None, step_done=0):
        rec = np.ascontiguousarray(rec, dtype=np.float64); lig = np.ascontiguousarray(lig, dtype=np.float64)
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.int32) for a in (rl, rs, rc)]
        ptr = [None if a is None else a.ctypes.data for a in arrs]
        self._ck(lib().kmc_strip_load_global(self.h, rec.shape[0], lig.shape[0], rec.ctypes.data, lig.ctypes.data, ptr[0], ptr[1], ptr[2], step_done))

    def strip_begin_refresh(self):
        self._ck(lib().kmc_strip_begin_refresh(self.h))

    def strip_message(self, side):
        """bytes of message `side` (0: to lower-x neighbour, 1: to higher-x neighbour, 2: the owned set) after strip_begin_refresh"""
        ptr = C.c_void_p()
        n = self._ck(lib().kmc_strip_message(self.h, side, C.byref(ptr)))
        return C.string_at(ptr, n) if n else b""

    def strip_rebuild(self, from_low, from_high):
        self._ck(lib().kmc_strip_rebuild(self.h, from_low, len(from_low), from_high, len(from_high)))

    def strip_comm_init(self, id128, refresh_every):
        """collective: NCCL communicator over the configured ranks; refresh_every > 0 = kmc_step refreshes the halos itself"""
        buf = C.create_string_buffer(bytes(id128), 128)
        self._ck(lib().kmc_strip_comm_init(self.h, buf, refresh_every))

    def strip_init_random(self, n_rec, n_lig, seed=1, sort_cells=True):
        """the global start state generated on this rank's GPU, the slab (owned + halo) kept"""
        self._ck(lib().kmc_strip_init_random(self.h, n_rec, n_lig, seed, int(sort_cells)))

    def strip_refresh(self):
        self._ck(lib().kmc_strip_refresh(self.h))

    def strip_series(self, reduce=True):
        s = Series()
        self._ck(lib().kmc_strip_get_series(self.h, int(reduce), C.byref(s)))
        return {k: getattr(s, k) for k, _ in Series._fields_ if k != "reserved"}

    def strip_oligomer_hist(self, reduce=True, nbins=64):
        hist = np.zeros(nbins, dtype=np.int64)
        self._ck(lib().kmc_strip_get_oligomer_hist(self.h, int(reduce), hist.ctypes.data, nbins))
        return hist

    def strip_get_records(self, which, out):
        """which 2 = owned units, 3 = everything local; out = uint8 host array (pinned recommended) -> (n_rec, n_lig)"""
        nr, nl = C.c_int64(), C.c_int64()
        self._ck(lib().kmc_strip_get_records(self.h, which, out.ctypes.data, out.nbytes, C.byref(nr), C.byref(nl)))
        return nr.value, nl.value

    def strip_load_records(self, buf, n_rec, n_lig, step_done=0):
        self._ck(lib().kmc_strip_load_records(self.h, buf.ctypes.data, n_rec, n_lig, step_done))

    def grid(self):
        x0, y0, edge, ncx, ncy = C.c_double(), C.c_double(), C.c_double(), C.c_int32(), C.c_int32()
        self._ck(lib().kmc_get_grid(self.h, C.byref(x0), C.byref(y0), C.byref(edge), C.byref(ncx), C.byref(ncy)))
        return dict(x0=x0.value, y0=y0.value, inv_edge=edge.value, ncx=ncx.value, ncy=ncy.value)

    def live_counts(self):
        a, b = C.c_int32(), C.c_int32()
        self._ck(lib().kmc_get_live_counts(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def path(self):
        """'fused' (small replicas: one CTA per replica, the whole step in one kernel) or 'general' (one CUDA graph per step)"""
        return "fused" if self._ck(lib().kmc_get_step_path(self.h)) == 1 else "general"

    def events(self):
        e = np.zeros(16, dtype=np.int64)
        self._ck(lib().kmc_get_events(self.h, e.ctypes.data))
        return dict(rl_on=int(e[0]), mono_cis_on=int(e[1]), cis_on=int(e[2]), rl_off=int(e[3]), mono_cis_off=int(e[4]),
                    cis_off=int(e[5]), reverted=int(e[6]), tried=int(e[7]), far=int(e[8]), passes=int(e[9]),
                    rebuilds=int(e[10]), launches=int(e[11]), list_pairs=int(e[12]), special_entries=int(e[13]),
                    pending_findings=int(e[14]), reaction_pairs=int(e[15]))

    def write_bond_dat(self, path, replica=0):
        self._ck(lib().kmc_write_bond_dat(self.h, replica, os.fsencode(path)))

    def write_cluster_log(self, path, replica=0):
        self._ck(lib().kmc_write_cluster_log(self.h, replica, os.fsencode(path)))

    def write_gro(self, path, replica=0):
        self._ck(lib().kmc_write_gro(self.h, replica, os.fsencode(path)))

    def write_checkpoint(self, path, replica=0):
        self._ck(lib().kmc_write_checkpoint(self.h, replica, os.fsencode(path)))

    def read_checkpoint(self, path, replica=0):
        self._ck(lib().kmc_read_checkpoint(self.h, replica, os.fsencode(path)))

    def write_checkpoint_bin(self, path):
        self._ck(lib().kmc_write_checkpoint_bin(self.h, os.fsencode(path)))

    def read_checkpoint_bin(self, path):
        self._ck(lib().kmc_read_checkpoint_bin(self.h, os.fsencode(path)))

    def run(self, n_steps, output_every, directory):
        self._ck(lib().kmc_run(self.h, n_steps, output_every, os.fsencode(directory)))
