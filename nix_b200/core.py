"""ctypes binding of libnixb200.so (include/nixb200.h) and the host-side `Domain` object.

The CUDA library is the product; this module is plumbing.  There is no CPU fallback: loading fails
loudly when the library has not been built, and every entry point fails when no sm_100 device is
present.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NIXB200_LIB") or os.path.join(HERE, "libnixb200.so")

MODE_FIELD, MODE_CURRENT, MODE_PARTICLE = 0, 1, 2
FIELD_UF, FIELD_UJ = 0, 1
ERR_UNSORTED, ERR_CFL, ERR_CAPACITY = 1, 2, 4

# every symbol include/nixb200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "nixb200_last_error", "nixb200_version", "nixb200_launch_count", "nixb200_domain_create",
    "nixb200_domain_destroy", "nixb200_domain_set_stream", "nixb200_domain_synchronize",
    "nixb200_domain_check", "nixb200_chunk_field_upload", "nixb200_chunk_field_download",
    "nixb200_domain_set_particles", "nixb200_domain_get_np", "nixb200_chunk_get_particles",
    "nixb200_chunk_get_pindex", "nixb200_chunk_get_pcount", "nixb200_domain_sort",
    "nixb200_domain_clear_current", "nixb200_domain_push_deposit", "nixb200_domain_exchange_current",
    "nixb200_domain_exchange_field", "nixb200_domain_migrate_sort", "nixb200_domain_step",
    "nixb200_halo_layout", "nixb200_chunk_halo_pack", "nixb200_chunk_halo_unpack",
    "nixb200_domain_get_load", "nixb200_domain_total_particles", "nixb200_domain_field_upload_async",
    "nixb200_domain_field_download_async", "nixb200_domain_field_upload_overlapped",
    "nixb200_domain_field_download_overlapped", "nixb200_domain_interior_upload_overlapped",
    "nixb200_domain_interior_download_overlapped", "nixb200_domain_copy_synchronize", "nixb200_domain_set_profiling", "nixb200_domain_get_phase_ms",
    "nixb200_plan_create", "nixb200_plan_destroy", "nixb200_plan_npeer", "nixb200_plan_peer",
    "nixb200_plan_entries", "nixb200_domain_set_ranks", "nixb200_comm_unique_id", "nixb200_domain_comm_init",
    "nixb200_domain_set_comm", "nixb200_domain_peer_traffic", "nixb200_domain_reserve",
    "nixb200_domain_get_capacity", "nixb200_domain_push_bfd", "nixb200_domain_push_efd", "nixb200_domain_step_em",
    "nixb200_domain_field_energy", "nixb200_domain_set_strict_fp", "nixb200_comm_create", "nixb200_comm_destroy",
    "nixb200_device_count", "nixb200_domain_deposit_moment", "nixb200_chunk_moment_download",
    "nixb200_chunk_pack_field", "nixb200_chunk_pack_moment", "nixb200_chunk_pack_tracer", "nixb200_shape_eval",
    "nixb200_chunk_wire_size", "nixb200_chunk_wire_pack", "nixb200_domain_rebalance", "nixb200_domain_history_async",
    "nixb200_rebalance_moves", "nixb200_halo_layout_dims", "nixb200_wire_size_dims",
    "nixb200_wire_particle_header",
]

PHASES = ("push_deposit", "exchange_current", "exchange_field", "migrate_sort", "sort", "k_push", "k_deposit",
          "field_solve")


PUSH_BORIS, PUSH_VAY, PUSH_HIGUERA_CARY = 0, 1, 2  # primitives.hpp:165-253


class DomainDesc(C.Structure):
    _fields_ = [
        ("cdims", C.c_int * 3),
        ("dims", C.c_int * 3),
        ("nb", C.c_int),
        ("order", C.c_int),
        ("ns", C.c_int),
        ("del_", C.c_double * 3),
        ("cc", C.c_double),
        ("id_begin", C.c_int),
        ("id_end", C.c_int),
        ("device", C.c_int),
        ("strict_fp", C.c_int),
        ("capacity_factor", C.c_double),
        ("pusher", C.c_int),
        ("fp32", C.c_int),
    ]


class NixB200Error(RuntimeError):
    pass


_lib = None


def load_library():
    """Load libnixb200.so; raises if the CUDA extension has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NixB200Error(
            f"{LIB_PATH} is missing: build it with `python -m nix_b200.build` "
            "(there is no CPU fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    P, I, D = C.c_void_p, C.c_int, C.c_double
    PI, PD, PL = C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_int64)

    def sig(fn, res, *args):
        f = getattr(lib, fn)
        f.restype = res
        f.argtypes = list(args)

    sig("nixb200_last_error", C.c_char_p)
    sig("nixb200_version", C.c_char_p)
    sig("nixb200_launch_count", C.c_int64)
    sig("nixb200_domain_create", I, C.POINTER(DomainDesc), PI, PD, PD, C.POINTER(P))
    sig("nixb200_domain_destroy", I, P)
    sig("nixb200_domain_set_stream", I, P, P)
    sig("nixb200_domain_synchronize", I, P)
    sig("nixb200_domain_check", I, P, PI)
    sig("nixb200_chunk_field_upload", I, P, I, I, PD)
    sig("nixb200_chunk_field_download", I, P, I, I, PD)
    sig("nixb200_domain_set_particles", I, P, I, PD, PL)
    sig("nixb200_domain_get_np", I, P, I, PL)
    sig("nixb200_chunk_get_particles", I, P, I, I, PD, C.c_int64, PL)
    sig("nixb200_chunk_get_pindex", I, P, I, I, C.POINTER(C.c_int32))
    sig("nixb200_chunk_get_pcount", I, P, I, I, C.POINTER(C.c_int32))
    for fn in ("sort", "clear_current", "exchange_current", "exchange_field", "migrate_sort"):
        sig("nixb200_domain_" + fn, I, P)
    sig("nixb200_domain_push_deposit", I, P, D)
    sig("nixb200_domain_step", I, P, D)
    sig("nixb200_halo_layout", I, P, I, PI, PI)
    sig("nixb200_halo_layout_dims", I, PI, I, I, PI, PI)
    sig("nixb200_chunk_halo_pack", I, P, I, I, P)
    sig("nixb200_chunk_halo_unpack", I, P, I, I, P, PI)
    sig("nixb200_domain_field_upload_async", I, P, I, P)
    sig("nixb200_domain_field_download_async", I, P, I, P)
    sig("nixb200_domain_field_upload_overlapped", I, P, I, P)
    sig("nixb200_domain_field_download_overlapped", I, P, I, P)
    sig("nixb200_domain_interior_upload_overlapped", I, P, I, P)
    sig("nixb200_domain_interior_download_overlapped", I, P, I, P)
    sig("nixb200_domain_copy_synchronize", I, P)
    sig("nixb200_domain_set_profiling", I, P, I)
    sig("nixb200_domain_get_phase_ms", I, P, I, PD, PI)
    sig("nixb200_domain_get_load", I, P, PD)
    sig("nixb200_domain_total_particles", C.c_int64, P)
    sig("nixb200_plan_create", I, PI, PI, I, PI, I, PI, I, C.POINTER(P))
    sig("nixb200_plan_destroy", I, P)
    sig("nixb200_plan_npeer", I, P)
    sig("nixb200_plan_peer", I, P, I, PI, PI, PI)
    sig("nixb200_plan_entries", I, P, I, PI, PI)
    sig("nixb200_domain_set_ranks", I, P, I, PI, I)
    sig("nixb200_comm_unique_id", I, P)
    sig("nixb200_domain_comm_init", I, P, P)
    sig("nixb200_domain_set_comm", I, P, P)
    sig("nixb200_domain_peer_traffic", I, P, PL, PL, PL)
    sig("nixb200_domain_push_bfd", I, P, D, I)
    sig("nixb200_domain_push_efd", I, P, D, D)
    sig("nixb200_domain_step_em", I, P, D, D)
    sig("nixb200_domain_field_energy", I, P, PD)
    sig("nixb200_domain_set_strict_fp", I, P, I)
    sig("nixb200_domain_history_async", I, P, P, P)
    sig("nixb200_domain_deposit_moment", I, P)
    sig("nixb200_chunk_moment_download", I, P, I, PD)
    sig("nixb200_chunk_pack_field", I, P, I, I, PD, PL)
    sig("nixb200_chunk_pack_moment", I, P, I, I, I, PD, PL)
    sig("nixb200_chunk_pack_tracer", I, P, I, I, PD, C.c_int64, PL)
    sig("nixb200_shape_eval", I, I, I, I, I, PD, PD, D, D, D, PD)
    sig("nixb200_chunk_wire_size", I, P, I, PL)
    sig("nixb200_wire_size_dims", I, PI, I, I, PI, PL)
    sig("nixb200_wire_particle_header", I, PI, I, PD, PI, PI, D, D, I, P)
    sig("nixb200_chunk_wire_pack", I, P, I, P, C.c_int64)
    sig("nixb200_domain_rebalance", I, P, I, PI, I)
    sig("nixb200_rebalance_moves", I, I, I, I, I, PI)
    sig("nixb200_comm_create", I, I, I, P, I, C.POINTER(C.c_void_p))
    sig("nixb200_comm_destroy", I, P)
    sig("nixb200_device_count", I, PI)
    sig("nixb200_domain_reserve", I, P, I, C.c_int64, C.c_int64)
    sig("nixb200_domain_get_capacity", I, P, I, PL, PL)
    _lib = lib
    return lib


def shape_eval(kind, order, x, X, rdx, dt=0.0, rdt=0.0, device=0):
    """shape_mc<order> (kind 0) / shape_wt<order> (kind 1) of the reference, evaluated on the device: [n][order+1]"""
    lib = load_library()
    x = np.ascontiguousarray(x, dtype=np.float64)
    X = np.ascontiguousarray(X, dtype=np.float64)
    out = np.empty((len(x), order + 1), dtype=np.float64)
    PD = C.POINTER(C.c_double)
    rc = lib.nixb200_shape_eval(int(device), int(kind), int(order), len(x), x.ctypes.data_as(PD), X.ctypes.data_as(PD),
                                float(rdx), float(dt), float(rdt), out.ctypes.data_as(PD))
    if rc != 0:
        raise NixB200Error(lib.nixb200_last_error().decode())
    return out


def launch_count():
    return int(load_library().nixb200_launch_count())


def lexicographic_coords(cdims):
    """id -> (cz,cy,cx) in plain z-major order (a valid, if not locality-preserving, chunk order)."""
    cz, cy, cx = cdims
    return np.array([(z, y, x) for z in range(cz) for y in range(cy) for x in range(cx)], dtype=np.int32)


class Plan:
    """Which slabs of which local chunk cross to which rank (host logic only, no device needed).

    peers[q] = dict(rank, send=[(chunk id, direction, cells)], recv=[(chunk id, receive slot, cells)]);
    entry j of my send list to rank r pairs with entry j of r's receive list from me."""

    def __init__(self, cdims, dims, nb, coord, boundary, rank):
        lib = load_library()
        cd = np.ascontiguousarray(cdims, dtype=np.int32)
        dm = np.ascontiguousarray(dims, dtype=np.int32)
        co = np.ascontiguousarray(coord, dtype=np.int32)
        bd = np.ascontiguousarray(boundary, dtype=np.int32)
        PI = C.POINTER(C.c_int)
        h = C.c_void_p()
        rc = lib.nixb200_plan_create(cd.ctypes.data_as(PI), dm.ctypes.data_as(PI), int(nb), co.ctypes.data_as(PI),
                                     len(bd) - 1, bd.ctypes.data_as(PI), int(rank), C.byref(h))
        if rc != 0:
            raise NixB200Error(lib.nixb200_last_error().decode())
        self.peers = []
        for q in range(lib.nixb200_plan_npeer(h)):
            r, ns, nr = C.c_int(), C.c_int(), C.c_int()
            lib.nixb200_plan_peer(h, q, C.byref(r), C.byref(ns), C.byref(nr))
            snd = np.zeros((ns.value, 3), dtype=np.int32)
            rcv = np.zeros((nr.value, 3), dtype=np.int32)
            lib.nixb200_plan_entries(h, q, snd.ctypes.data_as(PI), rcv.ctypes.data_as(PI))
            self.peers.append(dict(rank=r.value, send=[tuple(int(v) for v in e) for e in snd],
                                   recv=[tuple(int(v) for v in e) for e in rcv]))
        lib.nixb200_plan_destroy(h)


from .sfc import chunk_coords  # noqa: E402  (the reference's chunk order; re-exported)


def rebalance_moves(old, new):
    """(send_left, send_right, recv_left, recv_right, keep) id ranges of a rank whose range changes old -> new"""
    lib = load_library()
    out = (C.c_int * 10)()
    rc = lib.nixb200_rebalance_moves(int(old[0]), int(old[1]), int(new[0]), int(new[1]), out)
    v = list(out)
    return rc, [(v[0], v[1]), (v[2], v[3]), (v[4], v[5]), (v[6], v[7]), (v[8], v[9])]


def halo_layout_dims(dims, nb, mode):
    """(bufsize[27], bufaddr[27]) of a chunk's MpiBuffer (Chunk::set_mpi_buffer, chunk.cpp:257-286) from its shape"""
    lib = load_library()
    d = (C.c_int * 3)(*[int(v) for v in dims])
    bs, ba = (C.c_int * 27)(), (C.c_int * 27)()
    if lib.nixb200_halo_layout_dims(d, int(nb), int(mode), bs, ba):
        raise NixB200Error(lib.nixb200_last_error().decode())
    return np.array(bs, dtype=np.int32), np.array(ba, dtype=np.int32)


def wire_size_dims(dims, nb, np_species):
    """bytes of a chunk's payload in the reference's wire format (what follows nix::Chunk::pack's own header)"""
    lib = load_library()
    d = (C.c_int * 3)(*[int(v) for v in dims])
    n = (C.c_int * len(np_species))(*[int(v) for v in np_species])
    out = C.c_int64(0)
    if lib.nixb200_wire_size_dims(d, int(nb), len(np_species), n, C.byref(out)):
        raise NixB200Error(lib.nixb200_last_error().decode())
    return out.value


def wire_particle_header(dims, nb, delh, offset, gdims, q, m, np_):
    """the 175 scalar bytes XtensorParticle::pack opens a species with, as the device-made chunk record carries them"""
    lib = load_library()
    buf = (C.c_ubyte * 175)()
    rc = lib.nixb200_wire_particle_header((C.c_int * 3)(*[int(v) for v in dims]), int(nb), (C.c_double * 3)(*[float(v) for v in delh]),
                                          (C.c_int * 3)(*[int(v) for v in offset]), (C.c_int * 3)(*[int(v) for v in gdims]),
                                          float(q), float(m), int(np_), C.cast(buf, C.c_void_p))
    if rc:
        raise NixB200Error(lib.nixb200_last_error().decode())
    return bytes(buf)


def uniform_boundary(nchunk, nrank):
    """Rank boundaries for equal loads: what Balancer::assign_initial (balancer.cpp:101-124) returns
    for a uniform load vector (unittest/test_balancer.cpp:41-44: boundary[i] == i * nchunk / nrank)."""
    return np.array([(nchunk * r) // nrank for r in range(nrank + 1)], dtype=np.int32)


class Domain:
    """The chunks of one rank, resident on one B200 (mirrors the rank-local ChunkVec of
    nix::Application, application.hpp:114, behind the Chunk API)."""

    def __init__(self, cdims, dims, nb, order, q, m, delh=(1.0, 1.0, 1.0), cc=1.0, coord=None,
                 id_range=None, device=0, strict_fp=True, capacity_factor=1.25, stream=None, pusher=0, fp32=False):
        self.lib = load_library()
        self.cdims = tuple(int(v) for v in cdims)
        self.dims = tuple(int(v) for v in dims)
        self.nb, self.order = int(nb), int(order)
        self.q = np.ascontiguousarray(q, dtype=np.float64)
        self.m = np.ascontiguousarray(m, dtype=np.float64)
        self.ns = len(self.q)
        self.M = tuple(d + 2 * self.nb for d in self.dims)
        self.Ng = int(np.prod(self.M))
        ncid = int(np.prod(self.cdims))
        self.coord = np.ascontiguousarray(coord if coord is not None else lexicographic_coords(self.cdims),
                                          dtype=np.int32).reshape(ncid, 3)
        self.id_begin, self.id_end = (0, ncid) if id_range is None else (int(id_range[0]), int(id_range[1]))
        self.nchunk = self.id_end - self.id_begin
        desc = DomainDesc()
        desc.cdims[:] = self.cdims
        desc.dims[:] = self.dims
        desc.nb, desc.order, desc.ns = self.nb, self.order, self.ns
        desc.del_[:] = [float(v) for v in delh]
        desc.cc = float(cc)
        desc.id_begin, desc.id_end = self.id_begin, self.id_end
        desc.device = int(device)
        desc.strict_fp = int(bool(strict_fp))
        desc.capacity_factor = float(capacity_factor)
        desc.pusher = int(pusher)  # PUSH_BORIS / PUSH_VAY / PUSH_HIGUERA_CARY
        desc.fp32 = int(bool(fp32))  # fp32 on the device (the C ABI stays fp64)
        self.fp32 = bool(fp32)
        self.h = C.c_void_p()
        self._ck(self.lib.nixb200_domain_create(
            C.byref(desc), self.coord.ctypes.data_as(C.POINTER(C.c_int)),
            self.q.ctypes.data_as(C.POINTER(C.c_double)), self.m.ctypes.data_as(C.POINTER(C.c_double)),
            C.byref(self.h)))
        if stream is not None:
            self.set_stream(stream)

    # ---- several ranks ----
    def set_ranks(self, boundary, rank):
        """Chunks partitioned over ranks along the id order; boundary[rank:rank+2] must be my id range."""
        bd = np.ascontiguousarray(boundary, dtype=np.int32)
        self._ck(self.lib.nixb200_domain_set_ranks(self.h, len(bd) - 1, bd.ctypes.data_as(C.POINTER(C.c_int)),
                                                   int(rank)))

    def comm_init_torch(self, group=None):
        """NCCL bootstrap through an initialised torch.distributed group: rank 0 draws the unique id,
        the group broadcasts its 128 bytes, every rank joins the library's own communicator."""
        import torch
        import torch.distributed as dist
        buf = (C.c_ubyte * 128)()
        if dist.get_rank(group) == 0:
            self._ck(self.lib.nixb200_comm_unique_id(C.cast(buf, C.c_void_p)))
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=0, group=group)
        raw = bytes(t.cpu().tolist())
        self._ck(self.lib.nixb200_domain_comm_init(self.h, C.c_char_p(raw)))

    def rebalance(self, boundary, rank):
        """Collective: move to the new rank boundaries; chunks travel GPU to GPU in the reference's wire format."""
        bd = np.ascontiguousarray(boundary, dtype=np.int32)
        self._ck(self.lib.nixb200_domain_rebalance(self.h, len(bd) - 1, bd.ctypes.data_as(C.POINTER(C.c_int)), int(rank)))
        self.id_begin, self.id_end = int(bd[rank]), int(bd[rank + 1])
        self.nchunk = self.id_end - self.id_begin

    def wire_pack(self, k):
        """payload of chunk k in the reference's wire format (bytes after nix::Chunk::pack's header)"""
        n = C.c_int64(0)
        self._ck(self.lib.nixb200_chunk_wire_size(self.h, k, C.byref(n)))
        buf = np.zeros(n.value, dtype=np.uint8)
        self._ck(self.lib.nixb200_chunk_wire_pack(self.h, k, buf.ctypes.data_as(C.c_void_p), n.value))
        return buf

    def peer_traffic(self):
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self._ck(self.lib.nixb200_domain_peer_traffic(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(halo_cells_sent=a.value, particles_sent=b.value, particles_received=c.value)

    def _ck(self, rc):
        if rc != 0:
            raise NixB200Error(self.lib.nixb200_last_error().decode())

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.nixb200_domain_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing ----
    def set_stream(self, cuda_stream_ptr):
        self._ck(self.lib.nixb200_domain_set_stream(self.h, C.c_void_p(int(cuda_stream_ptr))))

    def synchronize(self):
        self._ck(self.lib.nixb200_domain_synchronize(self.h))

    def check(self):
        e = C.c_int(0)
        self._ck(self.lib.nixb200_domain_check(self.h, C.byref(e)))
        return e.value

    # ---- data in / out ----
    def set_field(self, k, uf):
        a = np.ascontiguousarray(uf, dtype=np.float64)
        assert a.shape == self.M + (6,)
        self._ck(self.lib.nixb200_chunk_field_upload(self.h, k, FIELD_UF, a.ctypes.data_as(C.POINTER(C.c_double))))

    def set_current(self, k, uj):
        a = np.ascontiguousarray(uj, dtype=np.float64)
        assert a.shape == self.M + (4,)
        self._ck(self.lib.nixb200_chunk_field_upload(self.h, k, FIELD_UJ, a.ctypes.data_as(C.POINTER(C.c_double))))

    def get_field(self, k):
        a = np.empty(self.M + (6,), dtype=np.float64)
        self._ck(self.lib.nixb200_chunk_field_download(self.h, k, FIELD_UF, a.ctypes.data_as(C.POINTER(C.c_double))))
        return a

    def get_current(self, k):
        a = np.empty(self.M + (4,), dtype=np.float64)
        self._ck(self.lib.nixb200_chunk_field_download(self.h, k, FIELD_UJ, a.ctypes.data_as(C.POINTER(C.c_double))))
        return a

    def set_particles(self, s, per_chunk):
        """per_chunk: list (one per local chunk) of [n][7] AoS arrays."""
        assert len(per_chunk) == self.nchunk
        npc = np.array([len(p) for p in per_chunk], dtype=np.int64)
        if npc.sum() > 0:
            flat = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, 7)
                                                        for p in per_chunk]))
        else:
            flat = np.zeros((0, 7))
        self.set_particles_flat(s, flat, npc)

    def set_particles_flat(self, s, flat, npc):
        flat = np.ascontiguousarray(flat, dtype=np.float64)
        npc = np.ascontiguousarray(npc, dtype=np.int64)
        self._ck(self.lib.nixb200_domain_set_particles(
            self.h, s, flat.ctypes.data_as(C.POINTER(C.c_double)), npc.ctypes.data_as(C.POINTER(C.c_int64))))

    def reserve(self, s, np_=0, nmove=0):
        """Room for np_ particles of species s (all chunks) and nmove migrating particles per step."""
        self._ck(self.lib.nixb200_domain_reserve(self.h, s, int(np_), int(nmove)))

    def capacity(self, s):
        a, b = C.c_int64(0), C.c_int64(0)
        self._ck(self.lib.nixb200_domain_get_capacity(self.h, s, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_particles_ptr(self, s, ptr, npc):
        """ptr: address of an [n][7] float64 AoS array in host OR device memory."""
        npc = np.ascontiguousarray(npc, dtype=np.int64)
        self._ck(self.lib.nixb200_domain_set_particles(
            self.h, s, C.cast(C.c_void_p(int(ptr)), C.POINTER(C.c_double)), npc.ctypes.data_as(C.POINTER(C.c_int64))))

    def get_np(self, s):
        a = np.zeros(self.nchunk, dtype=np.int64)
        self._ck(self.lib.nixb200_domain_get_np(self.h, s, a.ctypes.data_as(C.POINTER(C.c_int64))))
        return a

    def get_particles(self, k, s):
        n = C.c_int64(0)
        self._ck(self.lib.nixb200_chunk_get_particles(self.h, k, s, None, 0, C.byref(n)))
        out = np.empty((n.value, 7), dtype=np.float64)
        if n.value:
            self._ck(self.lib.nixb200_chunk_get_particles(
                self.h, k, s, out.ctypes.data_as(C.POINTER(C.c_double)), n.value, C.byref(n)))
        return out

    def get_pindex(self, k, s):
        a = np.empty(self.Ng + 1, dtype=np.int32)
        self._ck(self.lib.nixb200_chunk_get_pindex(self.h, k, s, a.ctypes.data_as(C.POINTER(C.c_int32))))
        return a

    def get_pcount(self, k, s):
        a = np.empty((self.Ng + 1, 8), dtype=np.int32)
        self._ck(self.lib.nixb200_chunk_get_pcount(self.h, k, s, a.ctypes.data_as(C.POINTER(C.c_int32))))
        return a

    # ---- the hot path ----
    def sort(self):
        self._ck(self.lib.nixb200_domain_sort(self.h))

    def clear_current(self):
        self._ck(self.lib.nixb200_domain_clear_current(self.h))

    def push_deposit(self, delt):
        self._ck(self.lib.nixb200_domain_push_deposit(self.h, float(delt)))

    def exchange_current(self):
        self._ck(self.lib.nixb200_domain_exchange_current(self.h))

    def exchange_field(self):
        self._ck(self.lib.nixb200_domain_exchange_field(self.h))

    def migrate_sort(self):
        self._ck(self.lib.nixb200_domain_migrate_sort(self.h))

    def step(self, delt):
        self._ck(self.lib.nixb200_domain_step(self.h, float(delt)))

    # ---- Yee field update on the device (row N1; the reference ships none) ----
    def push_bfd(self, delt, ext=0):
        self._ck(self.lib.nixb200_domain_push_bfd(self.h, float(delt), int(ext)))

    def push_efd(self, delt, cfj=1.0):
        self._ck(self.lib.nixb200_domain_push_efd(self.h, float(delt), float(cfj)))

    def step_em(self, delt, cfj=1.0):
        self._ck(self.lib.nixb200_domain_step_em(self.h, float(delt), float(cfj)))

    def field_energy(self):
        out = np.zeros((self.nchunk, 2), dtype=np.float64)
        self._ck(self.lib.nixb200_domain_field_energy(self.h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def history_async(self, e2b2_ptr, np_ptr):
        """field energies [nchunk][2] f64 and particle counts [ns][nchunk] i64 into PINNED host memory, in stream
        order, without synchronising (addresses; 0 / None skips one)"""
        self._ck(self.lib.nixb200_domain_history_async(self.h, C.c_void_p(int(e2b2_ptr or 0)), C.c_void_p(int(np_ptr or 0))))

    def set_strict_fp(self, on):
        self._ck(self.lib.nixb200_domain_set_strict_fp(self.h, int(bool(on))))

    # ---- diagnostics / output on the device (rows N3, N4) ----
    def deposit_moment(self):
        self._ck(self.lib.nixb200_domain_deposit_moment(self.h))

    def get_moment(self, k):
        a = np.empty(self.M + (self.ns, 14), dtype=np.float64)
        self._ck(self.lib.nixb200_chunk_moment_download(self.h, k, a.ctypes.data_as(C.POINTER(C.c_double))))
        return a

    def pack_field(self, k, decimate=1):
        n = C.c_int64(0)
        self._ck(self.lib.nixb200_chunk_pack_field(self.h, k, int(decimate), None, C.byref(n)))
        out = np.empty(n.value, dtype=np.float64)
        self._ck(self.lib.nixb200_chunk_pack_field(self.h, k, int(decimate), out.ctypes.data_as(C.POINTER(C.c_double)), C.byref(n)))
        return out

    def pack_moment(self, k, which, decimate=1):
        n = C.c_int64(0)
        self._ck(self.lib.nixb200_chunk_pack_moment(self.h, k, int(which), int(decimate), None, C.byref(n)))
        out = np.empty(n.value, dtype=np.float64)
        self._ck(self.lib.nixb200_chunk_pack_moment(self.h, k, int(which), int(decimate),
                                                    out.ctypes.data_as(C.POINTER(C.c_double)), C.byref(n)))
        return out

    def pack_tracer(self, k, s):
        n = C.c_int64(0)
        self._ck(self.lib.nixb200_chunk_pack_tracer(self.h, k, s, None, 0, C.byref(n)))
        out = np.empty((n.value, 7), dtype=np.float64)
        if n.value:
            self._ck(self.lib.nixb200_chunk_pack_tracer(self.h, k, s, out.ctypes.data_as(C.POINTER(C.c_double)), n.value, C.byref(n)))
        return out

    # ---- per-chunk halo buffers in the reference's MpiBuffer layout ----
    def halo_layout(self, mode):
        bs = np.zeros(27, dtype=np.int32)
        ba = np.zeros(27, dtype=np.int32)
        self._ck(self.lib.nixb200_halo_layout(self.h, mode, bs.ctypes.data_as(C.POINTER(C.c_int)),
                                              ba.ctypes.data_as(C.POINTER(C.c_int))))
        return bs, ba

    def halo_pack(self, k, mode):
        bs, ba = self.halo_layout(mode)
        buf = np.zeros(int(ba[26] + bs[26]), dtype=np.uint8)
        self._ck(self.lib.nixb200_chunk_halo_pack(self.h, k, mode, buf.ctypes.data_as(C.c_void_p)))
        return buf

    def halo_unpack(self, k, mode, buf, nbvalid=None):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        pv = None
        if nbvalid is not None:
            v = np.ascontiguousarray(nbvalid, dtype=np.int32)
            pv = v.ctypes.data_as(C.POINTER(C.c_int))
        self._ck(self.lib.nixb200_chunk_halo_unpack(self.h, k, mode, buf.ctypes.data_as(C.c_void_p), pv))

    def field_upload_async(self, which, host_ptr):
        """host_ptr: address of a (pinned) host buffer holding all local chunks back to back."""
        self._ck(self.lib.nixb200_domain_field_upload_async(self.h, which, C.c_void_p(int(host_ptr))))

    def field_download_async(self, which, host_ptr):
        self._ck(self.lib.nixb200_domain_field_download_async(self.h, which, C.c_void_p(int(host_ptr))))

    def field_upload_overlapped(self, which, host_ptr):
        """Host -> device on the domain's copy stream, ordered by events against the phases that use
        the array (host_ptr: page-locked memory)."""
        self._ck(self.lib.nixb200_domain_field_upload_overlapped(self.h, which, C.c_void_p(int(host_ptr))))

    def field_download_overlapped(self, which, host_ptr):
        self._ck(self.lib.nixb200_domain_field_download_overlapped(self.h, which, C.c_void_p(int(host_ptr))))

    def interior_upload_overlapped(self, which, host_ptr):
        """Interior cells only, host layout [chunk][Nz][Ny][Nx][6 or 4] (page-locked)."""
        self._ck(self.lib.nixb200_domain_interior_upload_overlapped(self.h, which, C.c_void_p(int(host_ptr))))

    def interior_download_overlapped(self, which, host_ptr):
        self._ck(self.lib.nixb200_domain_interior_download_overlapped(self.h, which, C.c_void_p(int(host_ptr))))

    def copy_synchronize(self):
        self._ck(self.lib.nixb200_domain_copy_synchronize(self.h))

    def set_profiling(self, on=True):
        self._ck(self.lib.nixb200_domain_set_profiling(self.h, int(on)))

    def phase_ms(self):
        """{phase: (ms_sum, calls)} accumulated since the last call (synchronises the stream)."""
        out = {}
        for i, name in enumerate(PHASES):
            ms, n = C.c_double(0), C.c_int(0)
            self._ck(self.lib.nixb200_domain_get_phase_ms(self.h, i, C.byref(ms), C.byref(n)))
            out[name] = (ms.value, n.value)
        return out

    def get_load(self):
        ms = C.c_double(0)
        self._ck(self.lib.nixb200_domain_get_load(self.h, C.byref(ms)))
        return ms.value

    def total_particles(self):
        return int(self.lib.nixb200_domain_total_particles(self.h))
