"""ctypes binding of the CPU oracle (oracle/nix_oracle.h).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
``--impl reference`` legs -- never by the product package ``nix_b200``.

Two interchangeable back ends implement the same C API:
  ``port``       oracle/libnixoracle.so       plain-C restatement (always available)
  ``ref``        oracle/_ref/libnixref.so     the reference's own templates (parity build)
  ``ref_v3/v4``  oracle/_ref/libnixref_v{3,4}.so  the reference's templates, -O3 AVX2 / AVX-512 (timing)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

MODE_FIELD, MODE_CURRENT, MODE_PARTICLE, MODE_MOMENT = 0, 1, 2, 3

_LIBS = {
    "port": os.path.join(HERE, "libnixoracle.so"),
    "port_fast": os.path.join(HERE, "libnixoracle_fast.so"),
    "ref": os.path.join(HERE, "_ref", "libnixref.so"),
    "ref_v3": os.path.join(HERE, "_ref", "libnixref_v3.so"),
    "ref_v4": os.path.join(HERE, "_ref", "libnixref_v4.so"),
}


class Geom(C.Structure):
    _fields_ = [
        ("dims", C.c_int * 3),
        ("nb", C.c_int),
        ("order", C.c_int),
        ("offset", C.c_int * 3),
        ("gdims", C.c_int * 3),
        ("del_", C.c_double * 3),
    ]


def build(which=("port", "ref")):
    """Compile the oracle libraries (the checker) with oracle/Makefile."""
    for target in which:
        subprocess.run(["make", "-s", "-C", HERE, target], check=True)


def available(name):
    return os.path.exists(_LIBS[name])


def cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def best_timing_backend():
    """Fastest back end this host can execute: reference AVX-512 > reference AVX2 > port."""
    flags = cpu_flags()
    if available("ref_v4") and {"avx512f", "avx512dq", "avx512bw", "avx512vl", "avx512cd"} <= flags:
        return "ref_v4"
    if available("ref_v3") and {"avx2", "fma", "bmi2"} <= flags:
        return "ref_v3"
    if available("ref"):
        return "ref"
    if available("port_fast") and {"avx2", "fma", "bmi2"} <= flags:
        return "port_fast"
    return "port"


_cache = {}


def load(name="port"):
    if name in _cache:
        return _cache[name]
    path = _LIBS[name]
    if not os.path.exists(path) and name.startswith("port"):
        build(("port",))
    lib = C.CDLL(path)
    P, I, D = C.c_void_p, C.c_int, C.c_double
    PI, PD = C.POINTER(C.c_int), C.POINTER(C.c_double)

    def sig(fn, res, *args):
        f = getattr(lib, fn)
        f.restype = res
        f.argtypes = list(args)

    sig("nixo_impl_name", C.c_char_p)
    sig("nixo_simd_lanes", I)
    sig("nixo_chunk_create", P, C.POINTER(Geom), I, PI, PD, PD)
    sig("nixo_chunk_destroy", None, P)
    sig("nixo_chunk_uf", PD, P)
    sig("nixo_chunk_uj", PD, P)
    sig("nixo_chunk_set_nb_valid", None, P, I, I, I, I)
    for fn in ("ng", "np", "np_total"):
        sig("nixo_particle_" + fn, I, P, I)
    sig("nixo_particle_set_np", None, P, I, I)
    sig("nixo_particle_xu", PD, P, I)
    sig("nixo_particle_xv", PD, P, I)
    for fn in ("gindex", "pindex", "pcount"):
        sig("nixo_particle_" + fn, C.POINTER(C.c_int32), P, I)
    sig("nixo_particle_resize", None, P, I, I)
    sig("nixo_particle_count", None, P, I, I, I, I, I)
    sig("nixo_particle_sort", None, P, I)
    sig("nixo_particle_set_boundary_periodic", None, P, I, I, I)
    sig("nixo_digitize", I, D, D, D)
    sig("nixo_shape_mc", None, I, D, D, D, PD)
    sig("nixo_push_boris", None, PD, PD, D)
    sig("nixo_push_vay", None, PD, PD, D)
    sig("nixo_push_higuera_cary", None, PD, PD, D)
    sig("nixo_set_pusher", None, I)
    sig("nixo_get_pusher", I)
    sig("nixo_lorentz_factor", D, D, D, D, D)
    sig("nixo_deposit3d", None, I, D, D, D, D, PD, PD)
    sig("nixo_interp3d", D, I, PD, I, I, I, I, I, I, PD, PD, PD, D)
    sig("nixo_interp_shift_weights", None, I, I, PD)
    sig("nixo_append_current3d", None, I, PD, I, I, I, I, I, PD)
    sig("nixo_append_moment3d", None, I, PD, I, I, I, I, I, I, I, PD)
    sig("nixo_esirkepov_shift_weights", None, I, C.POINTER(C.c_int), PD)
    sig("nixo_chunk_push_deposit", None, P, D, D, I)
    sig("nixo_chunk_halo_pack", None, P, I)
    sig("nixo_chunk_halo_unpack", None, P, I)
    sig("nixo_chunk_bufsize", I, P, I, I, I, I)
    sig("nixo_chunk_bufaddr", I, P, I, I, I, I)
    sig("nixo_chunk_sendbuf", C.POINTER(C.c_uint8), P, I)
    sig("nixo_chunk_sendbuf_size", I, P, I)
    sig("nixo_chunk_set_recv_sizes", None, P, I, PI)
    sig("nixo_chunk_recvbuf", C.POINTER(C.c_uint8), P, I)
    sig("nixo_chunk_recvbuf_size", I, P, I)
    sig("nixo_domain_create", P, PI, PI, I, I, PD, I, PD, PD, PI, PI)
    sig("nixo_domain_destroy", None, P)
    sig("nixo_domain_nchunk", I, P)
    sig("nixo_domain_chunk", P, P, I)
    sig("nixo_domain_neighbor", I, P, I, I, I, I)
    sig("nixo_domain_clear_current", None, P)
    sig("nixo_domain_push_deposit", None, P, D, D, I)
    sig("nixo_domain_exchange", None, P, I)
    sig("nixo_domain_sort_only", None, P)
    sig("nixo_domain_step", None, P, D, D, I)
    sig("nixo_domain_total_particles", C.c_int64, P)
    sig("nixo_shape_wt", None, I, D, D, D, D, D, PD)
    sig("nixo_chunk_um", PD, P)
    sig("nixo_chunk_deposit_moment", None, P, D)
    sig("nixo_chunk_pack_field", I, P, I, PD)
    sig("nixo_chunk_pack_moment", I, P, I, I, PD)
    sig("nixo_chunk_pack_tracer", I, P, I, PD)
    sig("nixo_domain_deposit_moment", None, P, D)
    sig("nixo_fdtd_push_bfd", None, PD, PI, I, PD, D, D, I)
    sig("nixo_fdtd_push_efd", None, PD, PD, PI, I, PD, D, D, D)
    sig("nixo_fdtd_energy", None, PD, PI, I, PD)
    sig("nixo_domain_push_bfd", None, P, D, D, I)
    sig("nixo_domain_push_efd", None, P, D, D, D)
    sig("nixo_domain_step_em", None, P, D, D, D, I)
    sig("nixo_set_num_threads", None, I)
    sig("nixo_get_num_threads", I)
    _cache[name] = lib
    return lib


def _iarr(v):
    a = np.ascontiguousarray(v, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int))


def _darr(v):
    a = np.ascontiguousarray(v, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


class Chunk:
    """One chunk of the oracle (owned unless it belongs to a Domain)."""

    def __init__(self, lib, dims, nb, order, ns=1, np_required=None, q=None, m=None, offset=(0, 0, 0),
                 gdims=None, delh=(1.0, 1.0, 1.0), handle=None):
        self.lib = lib
        self.dims = tuple(int(v) for v in dims)
        self.nb, self.order, self.ns = int(nb), int(order), int(ns)
        self.M = tuple(d + 2 * self.nb for d in self.dims)
        self.owned = handle is None
        if handle is None:
            g = Geom()
            g.dims[:] = self.dims
            g.nb, g.order = self.nb, self.order
            g.offset[:] = [int(v) for v in offset]
            g.gdims[:] = [int(v) for v in (gdims if gdims is not None else dims)]
            g.del_[:] = [float(v) for v in delh]
            npr, pnpr = _iarr(np_required if np_required is not None else [0] * ns)
            qa, pq = _darr(q if q is not None else [1.0] * ns)
            ma, pm = _darr(m if m is not None else [1.0] * ns)
            handle = lib.nixo_chunk_create(C.byref(g), ns, pnpr, pq, pm)
        self.h = handle

    def __del__(self):
        if getattr(self, "owned", False) and self.h:
            self.lib.nixo_chunk_destroy(self.h)
            self.h = None

    # --- grid arrays (numpy views on the oracle's memory) ---
    @property
    def uf(self):
        return np.ctypeslib.as_array(self.lib.nixo_chunk_uf(self.h), shape=self.M + (6,))

    @property
    def uj(self):
        return np.ctypeslib.as_array(self.lib.nixo_chunk_uj(self.h), shape=self.M + (4,))

    # --- particle container ---
    @property
    def um(self):
        return np.ctypeslib.as_array(self.lib.nixo_chunk_um(self.h), shape=self.M + (self.ns, 14))

    def deposit_moment(self, cc=1.0):
        self.lib.nixo_chunk_deposit_moment(self.h, float(cc))

    def pack_field(self, decimate=1):
        n = self.lib.nixo_chunk_pack_field(self.h, int(decimate), None)
        out = np.zeros(n)
        self.lib.nixo_chunk_pack_field(self.h, int(decimate), out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def pack_moment(self, which, decimate=1):
        n = self.lib.nixo_chunk_pack_moment(self.h, int(which), int(decimate), None)
        out = np.zeros(n)
        self.lib.nixo_chunk_pack_moment(self.h, int(which), int(decimate), out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def pack_tracer(self, s=0):
        n = self.lib.nixo_chunk_pack_tracer(self.h, int(s), None)
        out = np.zeros((n, 7))
        if n:
            self.lib.nixo_chunk_pack_tracer(self.h, int(s), out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def ng(self, s=0):
        return self.lib.nixo_particle_ng(self.h, s)

    def np(self, s=0):
        return self.lib.nixo_particle_np(self.h, s)

    def set_np(self, s, n):
        self.lib.nixo_particle_set_np(self.h, s, int(n))

    def np_total(self, s=0):
        return self.lib.nixo_particle_np_total(self.h, s)

    def xu(self, s=0):
        return np.ctypeslib.as_array(self.lib.nixo_particle_xu(self.h, s), shape=(self.np_total(s), 7))

    def xv(self, s=0):
        return np.ctypeslib.as_array(self.lib.nixo_particle_xv(self.h, s), shape=(self.np_total(s), 7))

    def gindex(self, s=0):
        return np.ctypeslib.as_array(self.lib.nixo_particle_gindex(self.h, s), shape=(self.np_total(s),))

    def pindex(self, s=0):
        return np.ctypeslib.as_array(self.lib.nixo_particle_pindex(self.h, s), shape=(self.ng(s) + 1,))

    def pcount(self, s=0):
        return np.ctypeslib.as_array(self.lib.nixo_particle_pcount(self.h, s), shape=(self.ng(s) + 1, 8))

    def set_particles(self, s, xu):
        """Load an [n][7] AoS array as the active particles of species s."""
        xu = np.ascontiguousarray(xu, dtype=np.float64)
        n = xu.shape[0]
        self.lib.nixo_particle_resize(self.h, s, n)
        assert self.np_total(s) > n or n == 0 or self.np_total(s) >= n
        self.xu(s)[:n] = xu
        self.set_np(s, n)

    def particles(self, s=0):
        return self.xu(s)[: self.np(s)].copy()

    def count(self, s, lbp, ubp, reset=True, order=None):
        self.lib.nixo_particle_count(self.h, s, lbp, ubp, int(reset), self.order if order is None else order)

    def sort(self, s=0):
        self.lib.nixo_particle_sort(self.h, s)

    def set_boundary_periodic(self, s, lbp, ubp):
        self.lib.nixo_particle_set_boundary_periodic(self.h, s, lbp, ubp)

    def push_deposit(self, delt, cc, simd=False):
        self.lib.nixo_chunk_push_deposit(self.h, delt, cc, int(simd))

    # --- halo engine ---
    def set_nb_valid(self, iz, iy, ix, valid):
        self.lib.nixo_chunk_set_nb_valid(self.h, iz, iy, ix, int(valid))

    def halo_pack(self, mode):
        self.lib.nixo_chunk_halo_pack(self.h, mode)

    def halo_unpack(self, mode):
        self.lib.nixo_chunk_halo_unpack(self.h, mode)

    def bufsize(self, mode):
        return np.array([self.lib.nixo_chunk_bufsize(self.h, mode, s // 9, (s // 3) % 3, s % 3) for s in range(27)],
                        dtype=np.int32)

    def bufaddr(self, mode):
        return np.array([self.lib.nixo_chunk_bufaddr(self.h, mode, s // 9, (s // 3) % 3, s % 3) for s in range(27)],
                        dtype=np.int32)

    def sendbuf(self, mode):
        n = self.lib.nixo_chunk_sendbuf_size(self.h, mode)
        if n == 0:
            return np.zeros(0, dtype=np.uint8)
        return np.ctypeslib.as_array(self.lib.nixo_chunk_sendbuf(self.h, mode), shape=(n,))

    def recvbuf(self, mode):
        n = self.lib.nixo_chunk_recvbuf_size(self.h, mode)
        if n == 0:
            return np.zeros(0, dtype=np.uint8)
        return np.ctypeslib.as_array(self.lib.nixo_chunk_recvbuf(self.h, mode), shape=(n,))

    def set_recv_sizes(self, mode, sizes):
        a, p = _iarr(sizes)
        self.lib.nixo_chunk_set_recv_sizes(self.h, mode, p)


class Domain:
    """Periodic box of Cz*Cy*Cx equal chunks with loop-back exchange (oracle/domain_driver.c)."""

    def __init__(self, lib, cdims, dims, nb, order, ns, q, m, coord, np_required, delh=(1.0, 1.0, 1.0)):
        self.lib = lib
        self.cdims = tuple(int(v) for v in cdims)
        self.dims = tuple(int(v) for v in dims)
        self.nb, self.order, self.ns = int(nb), int(order), int(ns)
        self.nchunk = int(np.prod(self.cdims))
        coord = np.ascontiguousarray(coord, dtype=np.int32).reshape(self.nchunk, 3)
        self.coord = coord
        a_c, p_c = _iarr(self.cdims)
        a_d, p_d = _iarr(self.dims)
        a_h, p_h = _darr(delh)
        a_q, p_q = _darr(q)
        a_m, p_m = _darr(m)
        a_k, p_k = _iarr(coord)
        a_n, p_n = _iarr(np.broadcast_to(np.asarray(np_required, dtype=np.int32), (self.nchunk, ns)))
        self.h = lib.nixo_domain_create(p_c, p_d, nb, order, p_h, ns, p_q, p_m, p_k, p_n)
        self.chunks = [
            Chunk(lib, dims, nb, order, ns, handle=lib.nixo_domain_chunk(self.h, k)) for k in range(self.nchunk)
        ]

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.nixo_domain_destroy(self.h)
            self.h = None

    def neighbor(self, k, iz, iy, ix):
        return self.lib.nixo_domain_neighbor(self.h, k, iz, iy, ix)

    def clear_current(self):
        self.lib.nixo_domain_clear_current(self.h)

    def push_deposit(self, delt, cc, simd=False):
        self.lib.nixo_domain_push_deposit(self.h, delt, cc, int(simd))

    def exchange(self, mode):
        self.lib.nixo_domain_exchange(self.h, mode)

    def sort_only(self):
        self.lib.nixo_domain_sort_only(self.h)

    def step(self, delt, cc, simd=False):
        self.lib.nixo_domain_step(self.h, delt, cc, int(simd))

    # Yee field update between the J halo and the E/B halo (oracle/field_solver.c; not in the reference tree)
    def push_bfd(self, delt, cc, ext=0):
        self.lib.nixo_domain_push_bfd(self.h, delt, cc, int(ext))

    def push_efd(self, delt, cc, cfj=1.0):
        self.lib.nixo_domain_push_efd(self.h, delt, cc, cfj)

    def step_em(self, delt, cc, cfj=1.0, simd=False):
        self.lib.nixo_domain_step_em(self.h, delt, cc, cfj, int(simd))

    def deposit_moment(self, cc=1.0):
        """every chunk's moments + XtensorHaloMoment3D exchange"""
        self.lib.nixo_domain_deposit_moment(self.h, float(cc))

    def field_energy(self):
        """[nchunk][2]: sum E^2, sum B^2 over the interior cells of every chunk"""
        out = np.zeros((self.nchunk, 2))
        a_d, p_d = _iarr(self.dims)
        for k, c in enumerate(self.chunks):
            e = np.zeros(2)
            self.lib.nixo_fdtd_energy(c.uf.ctypes.data_as(C.POINTER(C.c_double)), p_d, self.nb,
                                      e.ctypes.data_as(C.POINTER(C.c_double)))
            out[k] = e
        return out

    def total_particles(self):
        return int(self.lib.nixo_domain_total_particles(self.h))
