"""ctypes binding of libnlmc_b200.so (the C ABI declared in include/nlmc_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is usable, every
entry point raises.  The library is built in-tree by ``nonlocal-monte-carlo_b200/build.py``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# NLMC_LIB_PATH: an experiment build of the same library (tools/sweep_ab.py compares kernel variants); never a fallback
LIB_PATH = os.environ.get("NLMC_LIB_PATH") or os.path.join(_HERE, "libnlmc_b200.so")
_lib = None


class NlmcError(RuntimeError):
    pass


_i32 = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64 = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i8 = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")
_u8 = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u32 = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_vp = C.c_void_p
_int = C.c_int
_dbl = C.c_double
_u64 = C.c_uint64

# name -> (argtypes); every function returns int except where noted below
_SIGNATURES = {
    "nlmc_version": [],
    "nlmc_device_count": [],
    "nlmc_device_info": [_int, C.c_char_p, _int, C.POINTER(_int), C.POINTER(_int), C.POINTER(_int),
                         C.POINTER(_u64), C.POINTER(_u64)],
    "nlmc_instance_create": [_int, _i32, _i32, _f64, _f64, _int, C.POINTER(_vp)],
    "nlmc_instance_destroy": [_vp],
    "nlmc_instance_n": [_vp],
    "nlmc_instance_is_integer": [_vp],
    "nlmc_instance_is_symmetric": [_vp],
    "nlmc_replicas_create": [_vp, _int, _vp, C.POINTER(_vp)],
    "nlmc_replicas_destroy": [_vp],
    "nlmc_set_spins": [_vp, _int, _int, _i8],
    "nlmc_get_spins": [_vp, _int, _int, _i8],
    "nlmc_set_phase": [_vp, _int, _vp, _vp, _dbl],
    "nlmc_sweep_replay": [_vp, _int, _i32, _f64, _f64, _vp, _int, _vp, _int, _vp],
    "nlmc_energy": [_vp, _f64],
    "nlmc_energy_states": [_vp, _int, _i8, _f64],
    "nlmc_np_tanh": [_f64, _f64, C.c_int64, _int],
    "nlmc_np_arctanh": [_f64, _f64, C.c_int64, _int],
    "nlmc_lbp_create": [_vp, C.POINTER(_vp)],
    "nlmc_lbp_destroy": [_vp],
    "nlmc_lbp_epsilon": [_vp, _f64],
    "nlmc_lbp_reset": [_vp, _f64],
    "nlmc_lbp_step": [_vp, _dbl, _dbl, _dbl, _int, _vp, C.POINTER(_int)],
    "nlmc_lbp_run": [_vp, _f64, _dbl, _dbl, _int, _vp, C.POINTER(_int)],
    "nlmc_lbp_set_messages": [_vp, _f64, _f64, _f64],
    "nlmc_lbp_get_messages": [_vp, _vp, _vp, _vp],
    "nlmc_lbp_byproducts": [_vp, _dbl, _vp, _vp, _vp],
    "nlmc_icm_clusters": [_vp, _int, _i8, _i8, _i32, _i32],
    "nlmc_msc_create": [_vp, _int, _f64, _int, _int, C.c_ulonglong, C.POINTER(_vp)],
    "nlmc_msc_destroy": [_vp],
    "nlmc_msc_info": [_vp, C.POINTER(_int), C.POINTER(_int), C.POINTER(_int), C.POINTER(C.c_longlong)],
    "nlmc_msc_set_seed": [_vp, C.c_ulonglong, C.c_uint],
    "nlmc_msc_set_betas": [_vp, _f64],
    "nlmc_msc_init_random": [_vp, C.c_uint],
    "nlmc_msc_set_spins": [_vp, _int, _int, _i8],
    "nlmc_msc_get_spins": [_vp, _int, _int, _i8],
    "nlmc_msc_set_packed": [_vp, _vp],
    "nlmc_msc_get_packed": [_vp, _vp],
    "nlmc_msc_sweep": [_vp, _int],
    "nlmc_msc_energies": [_vp, _vp],
    "nlmc_msc_sweep_record": [_vp, _int, _int, _vp, _vp],
    "nlmc_msc_sweep_record_layout": [_vp, _int, _int, _vp, _vp, _int],
    "nlmc_msc_sweep_record_f64": [_vp, _int, _int, _vp, _vp],
    "nlmc_msc_sweep_record_dev": [_vp, _int, _int, _int, C.POINTER(_vp), C.POINTER(_vp)],
    "nlmc_msc_round": [_vp, _int, _int, _vp],
    "nlmc_host_widen_i8_f64": [_vp, _vp, _u64, _int],
    "nlmc_host_prefault": [_vp, _u64, _int],
    "nlmc_host_fetch_widen_blocks": [_vp, _vp, _int, _u64, _vp, _int, _vp],
    "nlmc_msc_round_host": [_vp, _vp, _int, _int, _vp, _vp],
    "nlmc_msc_round_host_async": [_vp, _vp, _int, _int, _vp, _vp],
    "nlmc_msc_swap_count": [_vp, C.POINTER(_int), _int],
    "nlmc_msc_swap_counts": [_vp, _int, _i32],
    "nlmc_msc_create_labelled": [_vp, _int, _f64, _int, _int, _int, _int, C.c_ulonglong, C.POINTER(_vp)],
    "nlmc_msc_set_stream": [_vp, _vp],
    "nlmc_msc_energies_dev": [_vp, _vp],
    "nlmc_msc_exchange_labels": [_vp, _vp, _int],
    "nlmc_msc_get_labels": [_vp, _u8],
    "nlmc_msc_sync": [_vp],
    "nlmc_msc_timer_mark": [_vp, _int],
    "nlmc_msc_timer_elapsed_ms": [_vp, C.POINTER(C.c_float)],
    "nlmc_col_create": [_vp, _int, _f64, _int, C.c_ulonglong, C.POINTER(_vp)],
    "nlmc_col_destroy": [_vp],
    "nlmc_col_info": [_vp, C.POINTER(_int), C.POINTER(_int)],
    "nlmc_col_set_betas": [_vp, _f64],
    "nlmc_col_set_spins": [_vp, _i8],
    "nlmc_col_get_spins": [_vp, _i8],
    "nlmc_col_set_site_modes": [_vp, _vp, _dbl],
    "nlmc_col_best_reset": [_vp],
    "nlmc_col_best_get": [_vp, _vp, _vp],
    "nlmc_col_sweep": [_vp, _int, _vp, _int, _vp, _vp, _int],
    "nlmc_col_energies": [_vp, _f64],
    "nlmc_col_sync": [_vp],
    "nlmc_col_ladders": [_vp, _int, _f64],
    "nlmc_col_exchange": [_vp, _int],
    "nlmc_col_labels": [_vp, _vp, _int, _vp],
    "nlmc_dense_create": [_vp, _int, _f64, _int, C.c_ulonglong, C.POINTER(_vp)],
    "nlmc_dense_destroy": [_vp],
    "nlmc_dense_set_betas": [_vp, _f64],
    "nlmc_dense_set_spins": [_vp, _i8],
    "nlmc_dense_get_spins": [_vp, _i8],
    "nlmc_dense_fields": [_vp, _vp],
    "nlmc_dense_sweep": [_vp, _int],
    "nlmc_dense_energies": [_vp, _f64],
    "nlmc_dense_set_site_modes": [_vp, _vp, _dbl],
    "nlmc_dense_best_reset": [_vp],
    "nlmc_dense_best_update": [_vp, _vp],
    "nlmc_dense_best_get": [_vp, _vp, _vp],
    "nlmc_dense_sync": [_vp],
    "nlmc_dense_ladders": [_vp, _int, _f64],
    "nlmc_dense_exchange": [_vp, _int],
    "nlmc_dense_labels": [_vp, _vp, _int, _vp],
    "nlmc_dense_time_fields": [_vp, _int, C.POINTER(C.c_float)],
    "nlmc_dense_time_sweeps": [_vp, _int, C.POINTER(C.c_float)],
}


def exported_symbols():
    """Names include/nlmc_b200.h declares (used by the CPU-side symbol test)."""
    return ["nlmc_last_error"] + list(_SIGNATURES)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NlmcError(f"{LIB_PATH} is missing: build it with `python nonlocal-monte-carlo_b200/build.py` "
                            "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.nlmc_last_error.restype = C.c_char_p
        L.nlmc_last_error.argtypes = []
        for name, args in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = C.c_int
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc < 0:
        msg = lib().nlmc_last_error().decode(errors="replace")
        raise NlmcError(f"{what or 'nlmc call'} failed ({rc}): {msg}")
    return rc


def require_device(device: int = 0):
    n = lib().nlmc_device_count()
    if n <= device:
        raise NlmcError(f"CUDA device {device} not available ({n} visible); nlmc_b200 has no CPU fallback")
    return n


def np_tanh(x, device: int = 0) -> np.ndarray:
    """np.tanh(x) for float64, evaluated on the device (bit-equal to numpy's AVX-512 routine)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    check(lib().nlmc_np_tanh(x.reshape(-1), out.reshape(-1), x.size, device), "nlmc_np_tanh")
    return out


def np_arctanh(x, device: int = 0) -> np.ndarray:
    """np.arctanh(x) for float64, evaluated on the device (bit-equal to the routine numpy dispatches to)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    check(lib().nlmc_np_arctanh(x.reshape(-1), out.reshape(-1), x.size, device), "nlmc_np_arctanh")
    return out


def _ptr(a):
    return None if a is None else a.ctypes.data


def empty_prefaulted(shape, dtype=np.float64) -> np.ndarray:
    """np.empty whose pages have been touched by the library's host threads (call it while the GPU is busy)."""
    a = np.empty(shape, dtype=dtype)
    if a.nbytes >= (1 << 22):
        check(lib().nlmc_host_prefault(a.ctypes.data, a.nbytes, 0), "nlmc_host_prefault")
    return a


class _ResultCache:
    """Caching host allocator for the large float64 result arrays (the 1 GB M of config C5): a buffer handed out earlier
    is reused once nothing outside the cache refers to it any more -- neither the array itself nor a view of it -- so a
    loop of run() calls pays the page faults of a fresh gigabyte once instead of on every call.  An array the caller still
    holds is never touched: the next request allocates a new one, and the cache keeps only the most recent buffer."""

    def __init__(self):
        self._buf = None

    def take(self, shape, dtype=np.float64) -> np.ndarray:
        import sys
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        if nbytes < (1 << 26):
            return np.empty(shape, dtype=dtype)
        b = self._buf
        # references to a free buffer: self._buf, the local b, getrefcount's argument
        if b is not None and b.nbytes >= nbytes and sys.getrefcount(b) == 3:
            return b[:nbytes].view(dtype).reshape(shape)
        del b
        self._buf = np.empty(nbytes, dtype=np.uint8)
        return self._buf[:nbytes].view(dtype).reshape(shape)

    def release(self):
        self._buf = None


result_cache = _ResultCache()


def release_host_cache():
    """Drop the cached result buffer (see _ResultCache)."""
    result_cache.release()


def fetch_widen_blocks(dev_ptr: int, out: np.ndarray, n_blocks: int, block_elems: int, dst_block=None, device: int = 0,
                       cuda_stream=None):
    """Device int8 (address dev_ptr, n_blocks x block_elems) -> host float64 `out`, block b landing at block dst_block[b]
    (None = identity); chunked through pinned staging and widened by the library's host workers."""
    assert out.dtype == np.float64 and out.flags.c_contiguous and out.size >= n_blocks * block_elems
    perm = None if dst_block is None else np.ascontiguousarray(dst_block, dtype=np.int32)
    check(lib().nlmc_host_fetch_widen_blocks(dev_ptr, out.ctypes.data, int(n_blocks), int(block_elems),
                                             None if perm is None else perm.ctypes.data, int(device), cuda_stream),
          "nlmc_host_fetch_widen_blocks")
    return out


class DevArray:
    """A device buffer given by address as a __cuda_array_interface__ object (torch.as_tensor(DevArray(...), device=...)
    wraps it without a copy).  The memory stays owned by whoever handed out the address."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2, "strides": None}


def widen_to_f64(a: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
    """int8 spins -> float64 (the dtype of every array the reference's API returns), multi-threaded in the library."""
    a = np.ascontiguousarray(a, dtype=np.int8)
    if out is None:
        out = np.empty(a.shape, dtype=np.float64)
    assert out.flags.c_contiguous and out.size == a.size and out.dtype == np.float64
    check(lib().nlmc_host_widen_i8_f64(a.ctypes.data, out.ctypes.data, a.size, 0), "nlmc_host_widen_i8_f64")
    return out


class Instance:
    """Device-resident instance: normalised J in CSR (scipy csr_matrix order) and h."""

    def __init__(self, rp, ci, val, h, device: int = 0):
        require_device(device)
        self.rp = np.ascontiguousarray(rp, dtype=np.int32)
        self.ci = np.ascontiguousarray(ci, dtype=np.int32)
        self.val = np.ascontiguousarray(val, dtype=np.float64)
        self.h = np.ascontiguousarray(np.asarray(h, dtype=np.float64).reshape(-1))
        self.n = len(self.rp) - 1
        self.nnz = len(self.ci)
        self.device = device
        if len(self.h) != self.n:
            raise ValueError(f"h has {len(self.h)} entries, J has {self.n} rows")
        ci_arg = self.ci if len(self.ci) else np.zeros(1, np.int32)
        val_arg = self.val if len(self.val) else np.zeros(1, np.float64)
        handle = _vp()
        check(lib().nlmc_instance_create(self.n, self.rp, ci_arg, val_arg, self.h, device, C.byref(handle)),
              "nlmc_instance_create")
        self._h = handle
        self.is_integer = bool(lib().nlmc_instance_is_integer(self._h))

    def energy_states(self, states) -> np.ndarray:
        states = np.ascontiguousarray(states, dtype=np.int8).reshape(-1, self.n)
        out = np.empty(states.shape[0], dtype=np.float64)
        check(lib().nlmc_energy_states(self._h, states.shape[0], states, out), "nlmc_energy_states")
        return out

    def close(self):
        if getattr(self, "_h", None):
            lib().nlmc_instance_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Lbp:
    """Device state of the LBP backbone search (K5) for one instance."""

    def __init__(self, inst: Instance):
        self.inst = inst
        handle = _vp()
        check(lib().nlmc_lbp_create(inst._h, C.byref(handle)), "nlmc_lbp_create")
        self._h = handle

    def epsilon(self) -> np.ndarray:
        out = np.empty(self.inst.n, dtype=np.float64)
        check(lib().nlmc_lbp_epsilon(self._h, out), "nlmc_lbp_epsilon")
        return out

    def reset(self, m_star):
        m = np.ascontiguousarray(np.asarray(m_star, dtype=np.float64).reshape(-1))
        check(lib().nlmc_lbp_reset(self._h, m), "nlmc_lbp_reset")

    def step(self, lam: float, beta: float, tol: float, max_iter: int):
        """One LoopyBeliefPropagation call; returns (marginal [n], iteration)."""
        marg = np.empty(self.inst.n, dtype=np.float64)
        it = _int(0)
        check(lib().nlmc_lbp_step(self._h, float(lam), float(beta), float(tol), int(max_iter), marg.ctypes.data,
                                  C.byref(it)), "nlmc_lbp_step")
        return marg, it.value

    def run(self, h_field, beta: float, tol: float, max_iter: int):
        """One LoopyBeliefPropagation call with the caller's field h[n]; returns (marginal [n], iteration)."""
        hf = np.ascontiguousarray(np.asarray(h_field, dtype=np.float64).reshape(-1))
        assert len(hf) == self.inst.n
        marg = np.empty(self.inst.n, dtype=np.float64)
        it = _int(0)
        check(lib().nlmc_lbp_run(self._h, hf, float(beta), float(tol), int(max_iter), marg.ctypes.data,
                                 C.byref(it)), "nlmc_lbp_run")
        return marg, it.value

    def set_messages(self, h_edge, u_edge, tot):
        he = np.ascontiguousarray(h_edge, dtype=np.float64).reshape(-1)
        ue = np.ascontiguousarray(u_edge, dtype=np.float64).reshape(-1)
        t = np.ascontiguousarray(tot, dtype=np.float64).reshape(-1)
        assert len(he) == len(ue) == self.inst.nnz and len(t) == self.inst.n
        check(lib().nlmc_lbp_set_messages(self._h, he, ue, t), "nlmc_lbp_set_messages")

    def get_messages(self):
        """(h_edge [nnz], u_edge [nnz], tot [n]): messages on the stored entries plus the off-entry row value."""
        he = np.empty(self.inst.nnz, dtype=np.float64)
        ue = np.empty(self.inst.nnz, dtype=np.float64)
        t = np.empty(self.inst.n, dtype=np.float64)
        check(lib().nlmc_lbp_get_messages(self._h, he.ctypes.data, ue.ctypes.data, t.ctypes.data),
              "nlmc_lbp_get_messages")
        return he, ue, t

    def byproducts(self, beta: float, want_corr: bool = True, want_J_tilde: bool = True):
        """(correlations [n][n] | None, h_tilde [n], J_tilde [n][n] | None) of the last call (NMC/nmc.py:217-226)."""
        n = self.inst.n
        corr = np.empty((n, n), dtype=np.float64) if want_corr else None
        jt = np.empty((n, n), dtype=np.float64) if want_J_tilde else None
        ht = np.empty(n, dtype=np.float64)
        check(lib().nlmc_lbp_byproducts(self._h, float(beta), _ptr(corr), ht.ctypes.data, _ptr(jt)),
              "nlmc_lbp_byproducts")
        return corr, ht, jt

    def close(self):
        if getattr(self, "_h", None):
            lib().nlmc_lbp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def icm_clusters(inst: Instance, s1, s2):
    """K7 for a batch of state pairs: returns (labels int32 [P][n], n_clusters int32 [P])."""
    s1 = np.ascontiguousarray(s1, dtype=np.int8).reshape(-1, inst.n)
    s2 = np.ascontiguousarray(s2, dtype=np.int8).reshape(-1, inst.n)
    P = s1.shape[0]
    labels = np.empty((P, inst.n), dtype=np.int32)
    counts = np.empty(P, dtype=np.int32)
    check(lib().nlmc_icm_clusters(inst._h, P, s1, s2, labels, counts), "nlmc_icm_clusters")
    return labels, counts


class Replicas:
    """R int8 spin configurations of one instance on the device (exact-replay path)."""

    def __init__(self, inst: Instance, n_replicas: int, init_spins=None):
        self.inst = inst
        self.R = int(n_replicas)
        self.n = inst.n
        init = None
        if init_spins is not None:
            init = np.ascontiguousarray(init_spins, dtype=np.int8).reshape(self.R, self.n)
        handle = _vp()
        check(lib().nlmc_replicas_create(inst._h, self.R, _ptr(init), C.byref(handle)), "nlmc_replicas_create")
        self._h = handle

    def set_spins(self, spins, first: int = 0):
        spins = np.ascontiguousarray(spins, dtype=np.int8).reshape(-1, self.n)
        check(lib().nlmc_set_spins(self._h, first, spins.shape[0], spins), "nlmc_set_spins")

    def get_spins(self, first: int = 0, count: int | None = None) -> np.ndarray:
        count = self.R - first if count is None else count
        out = np.empty((count, self.n), dtype=np.int8)
        check(lib().nlmc_get_spins(self._h, first, count, out), "nlmc_get_spins")
        return out

    def set_phase(self, r: int, h_eff=None, row_scaled=None, temp_x: float = 1.0):
        he = None if h_eff is None else np.ascontiguousarray(np.asarray(h_eff, dtype=np.float64).reshape(-1))
        rs = None if row_scaled is None else np.ascontiguousarray(row_scaled, dtype=np.uint8).reshape(-1)
        check(lib().nlmc_set_phase(self._h, r, _ptr(he), _ptr(rs), float(temp_x)), "nlmc_set_phase")

    def sweep_replay(self, perm, u, beta, tanh_lut=None, lut_half: int = 0, record_from: int | None = 0,
                     want_energy: bool = True):
        """perm,u: [R][S][n]; beta: [R][S].  Returns (M int8 [R][S-record_from][n] or None, E [R][S] or None)."""
        perm = np.ascontiguousarray(perm, dtype=np.int32).reshape(self.R, -1, self.n)
        S = perm.shape[1]
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(self.R, S, self.n)
        beta = np.ascontiguousarray(beta, dtype=np.float64).reshape(self.R, S)
        lut = None
        if tanh_lut is not None:
            lut = np.ascontiguousarray(tanh_lut, dtype=np.float64).reshape(self.R, S, 2 * lut_half + 1)
        M = None
        if record_from is not None:
            M = np.empty((self.R, S - record_from, self.n), dtype=np.int8)
        E = np.empty((self.R, S), dtype=np.float64) if want_energy else None
        check(lib().nlmc_sweep_replay(self._h, S, perm, u, beta, _ptr(lut), int(lut_half), _ptr(M),
                                      int(record_from or 0), _ptr(E)), "nlmc_sweep_replay")
        return M, E

    def energy(self) -> np.ndarray:
        out = np.empty(self.R, dtype=np.float64)
        check(lib().nlmc_energy(self._h, out), "nlmc_energy")
        return out

    def close(self):
        if getattr(self, "_h", None):
            lib().nlmc_replicas_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Msc:
    """Bit-packed production state (K2/K4'/K6): n_beta x n_ladders replicas of a +-J instance."""

    def __init__(self, inst: Instance, betas, n_ladders: int, seed: int = 0, ladder_offset: int = 0,
                 labelled: bool = False, slot_begin: int = 0, slot_count: int | None = None):
        """labelled=False: slot b sits at betas[b] and exchanges move configuration bits.  labelled=True: `betas` is the
        WHOLE ladder, the handle owns its slots [slot_begin, slot_begin + slot_count) and exchanges permute beta labels
        (north_star 4); a block of a ladder sharded over GPUs is driven by distributed.ShardedBetaLadder."""
        self.inst = inst
        self.ladder_offset = int(ladder_offset)
        self.n = inst.n
        self.labelled = bool(labelled)
        self.betas_total = np.ascontiguousarray(betas, dtype=np.float64).reshape(-1)
        self.n_beta_total = len(self.betas_total)
        self.slot_begin = int(slot_begin) if labelled else 0
        count = self.n_beta_total - self.slot_begin if slot_count is None else int(slot_count)
        self.betas = self.betas_total[self.slot_begin:self.slot_begin + count] if labelled else self.betas_total
        self.n_beta = len(self.betas)
        handle = _vp()
        if labelled:
            check(lib().nlmc_msc_create_labelled(inst._h, self.n_beta_total, self.betas_total, self.slot_begin, self.n_beta,
                                                 int(n_ladders), self.ladder_offset, int(seed) & (2**64 - 1),
                                                 C.byref(handle)), "nlmc_msc_create_labelled")
        else:
            check(lib().nlmc_msc_create(inst._h, self.n_beta, self.betas, int(n_ladders), self.ladder_offset,
                                        int(seed) & (2**64 - 1), C.byref(handle)), "nlmc_msc_create")
        self._h = handle
        w, lad, col, nb = _int(), _int(), _int(), C.c_longlong()
        check(lib().nlmc_msc_info(self._h, C.byref(w), C.byref(lad), C.byref(col), C.byref(nb)), "nlmc_msc_info")
        self.n_words, self.n_ladders, self.n_colours, self.n_bonds = w.value, lad.value, col.value, nb.value
        self.n_ladders_requested = int(n_ladders)

    def set_seed(self, seed: int, sweep_counter: int = 0):
        check(lib().nlmc_msc_set_seed(self._h, int(seed) & (2**64 - 1), int(sweep_counter)), "nlmc_msc_set_seed")

    def set_betas(self, betas):
        b = np.ascontiguousarray(betas, dtype=np.float64).reshape(-1)
        assert len(b) == self.n_beta
        check(lib().nlmc_msc_set_betas(self._h, b), "nlmc_msc_set_betas")
        self.betas = b

    def init_random(self, stream_id: int = 0):
        check(lib().nlmc_msc_init_random(self._h, int(stream_id)), "nlmc_msc_init_random")

    def set_spins(self, beta_idx: int, ladder: int, spins):
        s = np.ascontiguousarray(np.asarray(spins).reshape(-1), dtype=np.int8)
        check(lib().nlmc_msc_set_spins(self._h, int(beta_idx), int(ladder), s), "nlmc_msc_set_spins")

    def get_spins(self, beta_idx: int, ladder: int) -> np.ndarray:
        out = np.empty(self.n, dtype=np.int8)
        check(lib().nlmc_msc_get_spins(self._h, int(beta_idx), int(ladder), out), "nlmc_msc_get_spins")
        return out

    def packed_shape(self):
        return (self.n, self.n_words)

    def set_packed(self, packed: np.ndarray):
        assert packed.dtype == np.uint32 and packed.shape == self.packed_shape() and packed.flags.c_contiguous
        check(lib().nlmc_msc_set_packed(self._h, packed.ctypes.data), "nlmc_msc_set_packed")
        self.sync()

    def get_packed(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.packed_shape(), dtype=np.uint32)
        check(lib().nlmc_msc_get_packed(self._h, out.ctypes.data), "nlmc_msc_get_packed")
        return out

    def sweep(self, n_sweeps: int):
        check(lib().nlmc_msc_sweep(self._h, int(n_sweeps)), "nlmc_msc_sweep")

    def energies(self, fetch: bool = True):
        out = np.empty((self.n_beta, self.n_ladders), dtype=np.float64) if fetch else None
        check(lib().nlmc_msc_energies(self._h, _ptr(out)), "nlmc_msc_energies")
        return out

    def sweep_record(self, n_sweeps: int, ladder: int | None = 0, energies: bool = True, rows_of_M: bool = False):
        """n_sweeps sweeps recorded on the device: (M int8 [n_sweeps][n_beta][n] of `ladder` or None,
        E float64 [n_sweeps][n_beta][n_ladders] or None).  rows_of_M=True: M comes as [n_beta][n][n_sweeps], the
        layout of the reference's M, so that the host only has to widen it."""
        shape = (self.n_beta, self.n, n_sweeps) if rows_of_M else (n_sweeps, self.n_beta, self.n)
        Mrec = np.empty(shape, dtype=np.int8) if ladder is not None else None
        Erec = np.empty((n_sweeps, self.n_beta, self.n_ladders), dtype=np.float64) if energies else None
        check(lib().nlmc_msc_sweep_record_layout(self._h, int(n_sweeps), int(ladder or 0), _ptr(Mrec), _ptr(Erec),
                                                 1 if rows_of_M else 0), "nlmc_msc_sweep_record_layout")
        return Mrec, Erec

    def sweep_record_f64(self, n_sweeps: int, ladder: int = 0, out: np.ndarray | None = None):
        """n_sweeps recorded sweeps of `ladder` as the float64 rows of the reference's M: (M float64 [n_beta][n][n_sweeps],
        E float64 [n_sweeps][n_beta][n_ladders]); the int8 record travels through a pinned buffer of the library.
        `out` may be a preallocated (and prefaulted) C-contiguous float64 array of that size."""
        Mf = np.empty((self.n_beta, self.n, n_sweeps), dtype=np.float64) if out is None else out
        assert Mf.dtype == np.float64 and Mf.flags.c_contiguous and Mf.size == self.n_beta * self.n * n_sweeps
        Erec = np.empty((n_sweeps, self.n_beta, self.n_ladders), dtype=np.float64)
        check(lib().nlmc_msc_sweep_record_f64(self._h, int(n_sweeps), int(ladder), Mf.ctypes.data, Erec.ctypes.data),
              "nlmc_msc_sweep_record_f64")
        return Mf, Erec

    def sweep_record_dev(self, n_sweeps: int, ladder: int = 0, rows_of_M: bool = True, states: bool = True,
                         energies: bool = True):
        """n_sweeps recorded sweeps left on the device, no synchronisation: (address of the int8 states -- [n_beta][n]
        [n_sweeps] with rows_of_M, else [n_sweeps][n_beta][n] -- or None, address of the float64 energies [n_sweeps][n_beta]
        [n_ladders] or None).  The buffers belong to the handle (valid until its next record call)."""
        pm, pe = _vp(), _vp()
        check(lib().nlmc_msc_sweep_record_dev(self._h, int(n_sweeps), int(ladder), 1 if rows_of_M else 0,
                                              C.byref(pm) if states else None, C.byref(pe) if energies else None),
              "nlmc_msc_sweep_record_dev")
        return (pm.value if states else None), (pe.value if energies else None)

    def round(self, n_sweeps: int, num_swapping_pairs: int, fetch_energies: bool = False):
        out = np.empty((self.n_beta, self.n_ladders), dtype=np.float64) if fetch_energies else None
        check(lib().nlmc_msc_round(self._h, int(n_sweeps), int(num_swapping_pairs), _ptr(out)), "nlmc_msc_round")
        return out

    def round_host(self, packed_in_ptr, n_sweeps: int, num_swapping_pairs: int, packed_out_ptr, out_E_ptr):
        """Raw-pointer variant for pinned host buffers (bench.py's end-to-end leg)."""
        check(lib().nlmc_msc_round_host(self._h, packed_in_ptr, int(n_sweeps), int(num_swapping_pairs),
                                        packed_out_ptr, out_E_ptr), "nlmc_msc_round_host")

    def round_host_async(self, packed_in_ptr, n_sweeps: int, num_swapping_pairs: int, packed_out_ptr, out_E_ptr):
        """round_host without the final wait: outputs are valid after sync().  Two handles used alternately with
        pinned buffers overlap one batch's copies with the other's sweeps."""
        check(lib().nlmc_msc_round_host_async(self._h, packed_in_ptr, int(n_sweeps), int(num_swapping_pairs),
                                              packed_out_ptr, out_E_ptr), "nlmc_msc_round_host_async")

    def swap_count(self, reset: bool = False) -> int:
        v = _int()
        check(lib().nlmc_msc_swap_count(self._h, C.byref(v), int(reset)), "nlmc_msc_swap_count")
        return v.value

    def swap_counts(self, n_rounds: int) -> np.ndarray:
        """Accepted exchanges of each of the last n_rounds rounds (oldest first), counted on the device."""
        out = np.zeros(int(n_rounds), dtype=np.int32)
        if n_rounds:
            check(lib().nlmc_msc_swap_counts(self._h, int(n_rounds), out), "nlmc_msc_swap_counts")
        return out

    def set_stream(self, cuda_stream_ptr):
        """Run on the caller's CUDA stream (an int/pointer, e.g. torch.cuda.Stream().cuda_stream); None = own stream."""
        check(lib().nlmc_msc_set_stream(self._h, cuda_stream_ptr), "nlmc_msc_set_stream")

    def energies_dev(self, out_dev_ptr):
        """K4' into a device buffer [n_beta][n_ladders] float64 (no synchronisation)."""
        check(lib().nlmc_msc_energies_dev(self._h, out_dev_ptr), "nlmc_msc_energies_dev")

    def exchange_labels(self, E_full_dev_ptr, num_swapping_pairs: int):
        """Label exchange on the gathered energies [n_beta_total][n_ladders] (device pointer; no synchronisation)."""
        check(lib().nlmc_msc_exchange_labels(self._h, E_full_dev_ptr, int(num_swapping_pairs)), "nlmc_msc_exchange_labels")

    def energies_into(self, t):
        """energies_dev into a torch CUDA tensor [n_beta][n_ladders] float64 (contiguous)."""
        assert t.is_cuda and t.is_contiguous() and tuple(t.shape) == (self.n_beta, self.n_ladders)
        self.energies_dev(t.data_ptr())

    def exchange_labels_from(self, E_full, num_swapping_pairs: int):
        """exchange_labels on a torch CUDA tensor [n_beta_total][n_ladders] float64 (contiguous)."""
        assert E_full.is_cuda and E_full.is_contiguous() and tuple(E_full.shape) == (self.n_beta_total, self.n_ladders)
        self.exchange_labels(E_full.data_ptr(), num_swapping_pairs)

    def labels(self) -> np.ndarray:
        """labels[slot][ladder]: index of the temperature the configuration in (slot, ladder) is simulated at."""
        out = np.empty((self.n_beta_total, self.n_ladders), dtype=np.uint8)
        check(lib().nlmc_msc_get_labels(self._h, out), "nlmc_msc_get_labels")
        return out

    def sync(self):
        check(lib().nlmc_msc_sync(self._h), "nlmc_msc_sync")

    def timer_mark(self, which: int):
        check(lib().nlmc_msc_timer_mark(self._h, int(which)), "nlmc_msc_timer_mark")

    def timer_elapsed_ms(self) -> float:
        v = C.c_float()
        check(lib().nlmc_msc_timer_elapsed_ms(self._h, C.byref(v)), "nlmc_msc_timer_elapsed_ms")
        return float(v.value)

    def close(self):
        if getattr(self, "_h", None):
            lib().nlmc_msc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _LabelExchange:
    """Device-side replica exchange of the generic engines (rows grouped into ladders, row = ladder*n_beta + slot)."""
    _prefix = ""

    def ladders(self, betas):
        b = np.ascontiguousarray(betas, dtype=np.float64).reshape(-1)
        check(getattr(lib(), f"nlmc_{self._prefix}_ladders")(self._h, len(b), b), f"nlmc_{self._prefix}_ladders")
        self.ladder_betas = b

    def exchange(self, num_swapping_pairs: int):
        check(getattr(lib(), f"nlmc_{self._prefix}_exchange")(self._h, int(num_swapping_pairs)), f"nlmc_{self._prefix}_exchange")

    def labels(self, n_rounds: int = 0):
        """(labels int32 [R], accepted exchanges of each of the last n_rounds rounds)."""
        lab = np.empty(self.R, dtype=np.int32)
        cnt = np.zeros(max(int(n_rounds), 0), dtype=np.int32)
        check(getattr(lib(), f"nlmc_{self._prefix}_labels")(self._h, lab.ctypes.data, int(n_rounds), _ptr(cnt) if n_rounds else None),
              f"nlmc_{self._prefix}_labels")
        return lab, cnt



class Dense(_LabelExchange):
    _prefix = "dense"
    """Dense-J production state (K3): R replicas, field contraction H = S.J on tcgen05 tensor cores."""

    def __init__(self, inst: Instance, betas, n_split: int = 3, seed: int = 0):
        self.inst = inst
        self.n = inst.n
        self.betas = np.ascontiguousarray(betas, dtype=np.float64).reshape(-1)
        self.R = len(self.betas)
        self.n_split = int(n_split)
        handle = _vp()
        check(lib().nlmc_dense_create(inst._h, self.R, self.betas, self.n_split, int(seed) & (2**64 - 1),
                                      C.byref(handle)), "nlmc_dense_create")
        self._h = handle

    def set_betas(self, betas):
        b = np.ascontiguousarray(betas, dtype=np.float64).reshape(-1)
        assert len(b) == self.R
        check(lib().nlmc_dense_set_betas(self._h, b), "nlmc_dense_set_betas")
        self.betas = b

    def set_spins(self, spins):
        s = np.ascontiguousarray(spins, dtype=np.int8).reshape(self.R, self.n)
        check(lib().nlmc_dense_set_spins(self._h, s), "nlmc_dense_set_spins")

    def get_spins(self) -> np.ndarray:
        out = np.empty((self.R, self.n), dtype=np.int8)
        check(lib().nlmc_dense_get_spins(self._h, out), "nlmc_dense_get_spins")
        return out

    def fields(self, fetch: bool = True):
        out = np.empty((self.R, self.n), dtype=np.float32) if fetch else None
        check(lib().nlmc_dense_fields(self._h, _ptr(out)), "nlmc_dense_fields")
        return out

    def sweep(self, n_sweeps: int):
        check(lib().nlmc_dense_sweep(self._h, int(n_sweeps)), "nlmc_dense_sweep")

    def energies(self) -> np.ndarray:
        out = np.empty(self.R, dtype=np.float64)
        check(lib().nlmc_dense_energies(self._h, out), "nlmc_dense_energies")
        return out

    def set_site_modes(self, modes, temp_x: float = 1.0):
        m = None if modes is None else np.ascontiguousarray(modes, dtype=np.uint8).reshape(self.R, self.n)
        check(lib().nlmc_dense_set_site_modes(self._h, _ptr(m), float(temp_x)), "nlmc_dense_set_site_modes")

    def best_reset(self):
        check(lib().nlmc_dense_best_reset(self._h), "nlmc_dense_best_reset")

    def best_update(self, fetch: bool = True):
        out = np.empty(self.R, dtype=np.float64) if fetch else None
        check(lib().nlmc_dense_best_update(self._h, _ptr(out)), "nlmc_dense_best_update")
        return out

    def best_get(self):
        spins = np.empty((self.R, self.n), dtype=np.int8)
        E = np.empty(self.R, dtype=np.float64)
        check(lib().nlmc_dense_best_get(self._h, spins.ctypes.data, E.ctypes.data), "nlmc_dense_best_get")
        return spins, E

    def sync(self):
        check(lib().nlmc_dense_sync(self._h), "nlmc_dense_sync")

    def time_fields(self, repeats: int = 10) -> float:
        v = C.c_float()
        check(lib().nlmc_dense_time_fields(self._h, int(repeats), C.byref(v)), "nlmc_dense_time_fields")
        return float(v.value)

    def time_sweeps(self, n_sweeps: int = 2) -> float:
        v = C.c_float()
        check(lib().nlmc_dense_time_sweeps(self._h, int(n_sweeps), C.byref(v)), "nlmc_dense_time_sweeps")
        return float(v.value)

    def close(self):
        if getattr(self, "_h", None):
            lib().nlmc_dense_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Col(_LabelExchange):
    _prefix = "col"
    """Sparse production state (K2a): R replicas, graph-coloured parallel heat bath, one CTA per replica."""

    def __init__(self, inst: Instance, betas, seed: int = 0, replica_offset: int = 0):
        self.inst = inst
        self.n = inst.n
        self.betas = np.ascontiguousarray(betas, dtype=np.float64).reshape(-1)
        self.R = len(self.betas)
        handle = _vp()
        check(lib().nlmc_col_create(inst._h, self.R, self.betas, int(replica_offset), int(seed) & (2**64 - 1),
                                    C.byref(handle)), "nlmc_col_create")
        self._h = handle
        nc, sm = _int(), _int()
        check(lib().nlmc_col_info(self._h, C.byref(nc), C.byref(sm)), "nlmc_col_info")
        self.n_colours, self.csr_in_smem = nc.value, bool(sm.value)

    def set_betas(self, betas):
        b = np.ascontiguousarray(betas, dtype=np.float64).reshape(-1)
        assert len(b) == self.R
        check(lib().nlmc_col_set_betas(self._h, b), "nlmc_col_set_betas")
        self.betas = b

    def set_spins(self, spins):
        s = np.ascontiguousarray(spins, dtype=np.int8).reshape(self.R, self.n)
        check(lib().nlmc_col_set_spins(self._h, s), "nlmc_col_set_spins")

    def get_spins(self) -> np.ndarray:
        out = np.empty((self.R, self.n), dtype=np.int8)
        check(lib().nlmc_col_get_spins(self._h, out), "nlmc_col_get_spins")
        return out

    def set_site_modes(self, modes, temp_x: float = 1.0):
        m = None if modes is None else np.ascontiguousarray(modes, dtype=np.uint8).reshape(self.R, self.n)
        check(lib().nlmc_col_set_site_modes(self._h, _ptr(m), float(temp_x)), "nlmc_col_set_site_modes")

    def best_reset(self):
        check(lib().nlmc_col_best_reset(self._h), "nlmc_col_best_reset")

    def best_get(self):
        spins = np.empty((self.R, self.n), dtype=np.int8)
        E = np.empty(self.R, dtype=np.float64)
        check(lib().nlmc_col_best_get(self._h, spins.ctypes.data, E.ctypes.data), "nlmc_col_best_get")
        return spins, E

    def sweep(self, n_sweeps: int):
        check(lib().nlmc_col_sweep(self._h, int(n_sweeps), None, 0, None, None, 0), "nlmc_col_sweep")

    def sweep_record(self, n_sweeps: int, record_every: int = 1, track_best: bool = False, beta_sched=None,
                     want_states: bool = True, want_energies: bool = True):
        """n_sweeps sweeps in one launch -> (states int8 [n_rec][R][n] or None, E [n_sweeps][R] or None)."""
        n_rec = (n_sweeps + record_every - 1) // record_every if want_states else 0
        states = np.empty((n_rec, self.R, self.n), dtype=np.int8) if want_states else None
        E = np.empty((n_sweeps, self.R), dtype=np.float64) if want_energies else None
        sched = None
        if beta_sched is not None:
            sched = np.ascontiguousarray(beta_sched, dtype=np.float64).reshape(n_sweeps, self.R)
        check(lib().nlmc_col_sweep(self._h, int(n_sweeps), _ptr(sched), int(record_every), _ptr(states), _ptr(E),
                                   int(track_best)), "nlmc_col_sweep")
        return states, E

    def energies(self) -> np.ndarray:
        out = np.empty(self.R, dtype=np.float64)
        check(lib().nlmc_col_energies(self._h, out), "nlmc_col_energies")
        return out

    def sync(self):
        check(lib().nlmc_col_sync(self._h), "nlmc_col_sync")

    def close(self):
        if getattr(self, "_h", None):
            lib().nlmc_col_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
