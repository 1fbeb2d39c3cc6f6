"""Python mirror of the reference's cs_matrix / cs_sles_it / cs_multigrid interface,
bound with ctypes to the C-ABI of ``csrc/libcsb200.so`` (include/csb200.h).

This is what the parity tests and bench.py call; it adds no arithmetic of its
own.  Names and argument meaning follow the reference:

* ``Matrix``            cs_matrix_t (MSR)          src/alge/cs_matrix.cpp:4518,:5451,:6328,:6405;
                        ``compute_coeffs`` = cs_matrix_compute_coeffs, src/alge/cs_matrix_building.cpp:1948
* ``SlesIt``            cs_sles_it_t               src/alge/cs_sles_it.cpp:4659,:4964,:5194;
                        ``transfer_pc`` = cs_sles_it_transfer_pc
* ``Smoother``          cs_multigrid_smoother_*    src/alge/cs_multigrid_smoother.cpp:1726,:1825,:2008
* ``Multigrid``         cs_multigrid_t             src/alge/cs_multigrid.cpp:4644,:5032,:5275,:5362;
                        ``Multigrid(pc=True)`` = cs_multigrid_pc_create (:5682), V- or K-cycle
* ``ConvergenceState``  cs_sles_convergence_state_t  src/alge/cs_sles.h:57-68

The library has no CPU fallback: importing works anywhere (so that the symbol
table can be checked), but ``init()`` raises unless a CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import enum
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "csrc", "libcsb200.so")
# CSB_LIB: another build of the same sources -- tuning variants (-DCSB_PHASE_TRACE ...) and the
# host-thread stand-in that tests/cudasim builds to desk-check kernels without a GPU (test
# infrastructure only; nothing in this package builds or selects it)
if os.environ.get("CSB_LIB"):
    LIB_PATH = os.path.abspath(os.environ["CSB_LIB"])
HEADER_PATH = os.path.join(_ROOT, "include", "csb200.h")


class CsbError(RuntimeError):
    pass


class ConvergenceState(enum.IntEnum):
    DIVERGED = -3
    BREAKDOWN = -2
    MAX_ITERATION = -1
    ITERATING = 0
    CONVERGED = 1


class SlesItType(enum.IntEnum):
    PCG = 0
    FCG = 1
    IPCG = 2
    JACOBI = 3
    BICGSTAB = 4
    GCR = 6
    # process-local Gauss-Seidel family (cs_sles_it.h:64-65, :84-85)
    P_GAUSS_SEIDEL = 8
    P_SYM_GAUSS_SEIDEL = 9
    TS_F_GAUSS_SEIDEL = 17
    TS_B_GAUSS_SEIDEL = 18
    # multigrid smoothers (cs_sles_it.h:79-83)
    L1_JACOBI = 13
    R_JACOBI = 14
    RJ2 = 15
    RJ3 = 16


_lib = None


def header_symbols() -> list[str]:
    """Every function name declared in include/csb200.h."""
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(csb_[a-z0-9_]+)\s*\(", txt)))


def load_library() -> C.CDLL:
    """dlopen libcsb200.so and declare prototypes.  Fails loudly if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CsbError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, ci, cd, cs = C.c_void_p, C.c_int, C.c_double, C.c_size_t
    L.csb_last_error.restype = C.c_char_p
    L.csb_version.restype = C.c_char_p
    L.csb_spmv_kernel_name.restype = C.c_char_p
    L.csb_init.argtypes = [ci]
    L.csb_set_stream.argtypes = [vp]
    L.csb_malloc_device.restype = vp
    L.csb_malloc_device.argtypes = [cs]
    L.csb_free_device.argtypes = [vp]
    L.csb_copy_h2d.argtypes = [vp, vp, cs]
    L.csb_copy_d2h.argtypes = [vp, vp, cs]
    L.csb_malloc_pinned.restype = vp
    L.csb_malloc_pinned.argtypes = [cs]
    L.csb_free_pinned.argtypes = [vp]
    L.csb_launch_count.restype = C.c_longlong
    L.csb_launch_count.argtypes = [ci]
    L.csb_matrix_create_msr_block.restype = vp
    L.csb_matrix_create_msr_block.argtypes = [ci, ci, ci, vp, vp, vp, vp, ci]
    L.csb_matrix_create_csr.restype = vp
    L.csb_matrix_create_csr.argtypes = [ci, ci, vp, vp, vp, ci]
    cll = C.c_longlong
    L.csb_matrix_assembler_create.restype = vp
    L.csb_matrix_assembler_create.argtypes = [vp, ci]
    L.csb_matrix_assembler_destroy.argtypes = [C.POINTER(vp)]
    L.csb_matrix_assembler_add_g_ids.argtypes = [vp, ci, vp, vp]
    L.csb_matrix_assembler_compute.argtypes = [vp]
    L.csb_matrix_assembler_get_n_rows.argtypes = [vp]
    L.csb_matrix_assembler_get_structure.argtypes = [vp, vp, vp]
    L.csb_matrix_create_from_assembler.restype = vp
    L.csb_matrix_create_from_assembler.argtypes = [vp]
    L.csb_matrix_assembler_values_init.restype = vp
    L.csb_matrix_assembler_values_init.argtypes = [vp, vp, ci, ci]
    L.csb_matrix_assembler_values_add_g.argtypes = [vp, ci, vp, vp, vp]
    L.csb_matrix_assembler_values_done.argtypes = [vp]
    L.csb_matrix_assembler_values_add_declared.argtypes = [vp, vp]
    L.csb_matrix_assembler_values_finalize.argtypes = [C.POINTER(vp)]
    L.csb_matrix_create_native.restype = vp
    L.csb_matrix_create_native.argtypes = [ci, ci, ci, vp, vp, vp, ci]
    L.csb_matrix_set_coefficients_native.argtypes = [vp, vp, vp]
    L.csb_matrix_create_msr.restype = vp
    L.csb_matrix_create_msr.argtypes = [ci, ci, vp, vp, vp, vp, ci]
    L.csb_matrix_set_coefficients.argtypes = [vp, ci, vp, vp]
    L.csb_matrix_share_host_arrays.argtypes = [vp, vp, vp, vp]
    L.csb_host_pool_trim.restype = None
    L.csb_matrix_set_mesh_association.argtypes = [vp, ci, vp, vp, vp, vp, vp, vp]
    L.csb_matrix_set_halo.argtypes = [vp, vp]
    L.csb_matrix_destroy.argtypes = [C.POINTER(vp)]
    for f in (L.csb_matrix_get_n_rows, L.csb_matrix_get_n_columns, L.csb_matrix_get_n_entries):
        f.argtypes = [vp]
    L.csb_matrix_vector_multiply.argtypes = [vp, vp, vp]
    L.csb_matrix_vector_multiply_exdiag.argtypes = [vp, vp, vp]
    L.csb_matrix_spmv_bytes.restype = cd
    L.csb_matrix_spmv_bytes.argtypes = [vp]
    L.csb_dot.restype = cd
    L.csb_dot.argtypes = [ci, vp, vp]
    L.csb_dot_xx.restype = cd
    L.csb_dot_xx.argtypes = [ci, vp]
    L.csb_dot_xx_xy_yz.argtypes = [ci, vp, vp, vp, vp, vp, vp]
    L.csb_comm_get_unique_id.argtypes = [vp]
    L.csb_comm_init_rank.restype = vp
    L.csb_comm_init_rank.argtypes = [vp, ci, ci]
    L.csb_comm_destroy.argtypes = [vp]
    L.csb_halo_create.restype = vp
    L.csb_halo_create.argtypes = [vp, ci, ci, vp, ci, vp, vp, vp]
    L.csb_halo_destroy.argtypes = [C.POINTER(vp)]
    L.csb_halo_sync_var.argtypes = [vp, vp]
    L.csb_pcomm_create.restype = vp
    L.csb_pcomm_create.argtypes = [ci, ci, cs, cs]
    L.csb_pcomm_get_ipc_handle.argtypes = [vp, vp]
    L.csb_pcomm_open_peers.argtypes = [vp, vp]
    L.csb_pcomm_group_create.restype = vp
    L.csb_pcomm_group_create.argtypes = [ci]
    L.csb_pcomm_group_destroy.argtypes = [C.POINTER(vp)]
    L.csb_pcomm_group_destroy.restype = None
    L.csb_pcomm_create_in_group.restype = vp
    L.csb_pcomm_create_in_group.argtypes = [vp, ci, cs, cs]
    L.csb_pcomm_destroy.argtypes = [C.POINTER(vp)]
    L.csb_pcomm_destroy.restype = None
    L.csb_pcomm_rank.argtypes = [vp]
    L.csb_pcomm_n_ranks.argtypes = [vp]
    L.csb_pcomm_allreduce_sum.argtypes = [vp, vp, ci]
    L.csb_pcomm_barrier.argtypes = [vp]
    L.csb_halo_set_pcomm.argtypes = [vp, vp]
    L.csb_sles_it_set_pcomm.argtypes = [vp, vp]
    L.csb_multigrid_set_pcomm.argtypes = [vp, vp]
    L.csb_multigrid_get_level_halo.argtypes = [vp, ci, vp, vp, vp, vp, vp]
    L.csb_sles_it_create.restype = vp
    L.csb_sles_it_create.argtypes = [ci, ci, ci, ci]
    L.csb_sles_it_set_reduce_comm.argtypes = [vp, vp]
    L.csb_sles_it_set_fallback_threshold.argtypes = [vp, ci, ci]
    L.csb_sles_it_set_fallback_threshold.restype = None
    L.csb_sles_it_set_restart_interval.argtypes = [vp, ci]
    L.csb_sles_it_set_restart_interval.restype = None
    L.csb_sles_it_used_fallback.argtypes = [vp]
    L.csb_sles_it_setup.argtypes = [vp, C.c_char_p, vp, ci]
    L.csb_sles_it_setup.restype = None
    solve_args = [vp, C.c_char_p, vp, ci, cd, cd, C.POINTER(ci), C.POINTER(cd), vp, vp, vp,
                  cs, vp]
    L.csb_sles_it_solve.argtypes = solve_args
    L.csb_sles_it_free.argtypes = [vp]
    L.csb_sles_it_free.restype = None
    L.csb_sles_it_destroy.argtypes = [C.POINTER(vp)]
    L.csb_sles_it_destroy.restype = None
    L.csb_sles_it_get_last_initial_residual.restype = cd
    L.csb_sles_it_get_last_initial_residual.argtypes = [vp]
    L.csb_sles_it_get_residual_history.argtypes = [vp, vp, ci]
    L.csb_sles_it_log.argtypes = [vp, ci]
    L.csb_sles_it_log.restype = None
    L.csb_sles_it_copy.restype = vp
    L.csb_sles_it_copy.argtypes = [vp]
    L.csb_multigrid_create.restype = vp
    L.csb_multigrid_create.argtypes = [ci]
    L.csb_multigrid_set_solver_options.argtypes = [vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci,
                                                   cd, cd, cd]
    L.csb_multigrid_set_solver_options.restype = None
    L.csb_multigrid_set_coarsening_options.argtypes = [vp, ci, ci, ci, C.c_ulonglong, cd, ci]
    L.csb_multigrid_set_coarsening_options.restype = None
    L.csb_multigrid_set_aggregation_blocks.argtypes = [vp, ci]
    L.csb_multigrid_set_aggregation_blocks.restype = None
    L.csb_multigrid_set_reduce_comm.argtypes = [vp, vp]
    L.csb_multigrid_setup.argtypes = [vp, C.c_char_p, vp, ci]
    L.csb_multigrid_setup.restype = None
    L.csb_multigrid_solve.argtypes = solve_args
    L.csb_multigrid_free.argtypes = [vp]
    L.csb_multigrid_free.restype = None
    L.csb_multigrid_destroy.argtypes = [C.POINTER(vp)]
    L.csb_multigrid_destroy.restype = None
    L.csb_multigrid_get_n_levels.argtypes = [vp]
    L.csb_multigrid_is_setup.argtypes = [vp]
    L.csb_multigrid_get_n_cycles.argtypes = [vp]
    L.csb_multigrid_get_level_info.argtypes = [vp, ci, vp]
    L.csb_multigrid_get_level_arrays.argtypes = [vp, ci, vp, vp, vp, vp, vp]
    L.csb_multigrid_get_cycle_residuals.argtypes = [vp, vp, ci]
    L.csb_multigrid_get_setup_times.argtypes = [vp, vp]
    L.csb_multigrid_get_phase_times.argtypes = [vp, vp, ci]
    L.csb_multigrid_get_smoother_times.argtypes = [vp, vp, ci]
    L.csb_multigrid_log.argtypes = [vp, ci]
    L.csb_multigrid_log.restype = None
    L.csb_multigrid_copy.restype = vp
    L.csb_multigrid_copy.argtypes = [vp]
    L.csb_multigrid_pc_create.restype = vp
    L.csb_multigrid_pc_create.argtypes = [ci]
    L.csb_sles_it_transfer_pc.argtypes = [vp, C.POINTER(vp)]
    L.csb_sles_it_get_pc.restype = vp
    L.csb_sles_it_get_pc.argtypes = [vp]
    L.csb_multigrid_smoother_create.restype = vp
    L.csb_multigrid_smoother_create.argtypes = [ci, ci, ci]
    L.csb_multigrid_smoother_setup.argtypes = [vp, C.c_char_p, vp, ci]
    L.csb_multigrid_smoother_setup.restype = None
    L.csb_multigrid_smoother_solve.argtypes = solve_args
    L.csb_host_aggregate_dx_msr.argtypes = [ci, ci, vp, vp, vp, vp, ci, ci, vp]
    L.csb_host_coarse_structure.argtypes = [ci, vp, vp, vp, ci, vp, vp, vp, vp, vp]
    L.csb_host_coarsen_pairwise.argtypes = [ci, ci, vp, vp, vp, vp, ci, vp, vp, vp, vp, vp]
    _lib = L
    return L


def _err(what: str):
    L = load_library()
    raise CsbError(f"{what}: {L.csb_last_error().decode()}")


def init(device: int = 0) -> None:
    """csb_init: select the device.  Raises without a CUDA device (no fallback)."""
    if load_library().csb_init(device) != 0:
        _err("csb_init")


def finalize() -> None:
    load_library().csb_finalize()


def set_stream(cuda_stream: int | None) -> None:
    """Run all kernels on the given cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream)."""
    if load_library().csb_set_stream(C.c_void_p(cuda_stream or 0)) != 0:
        _err("csb_set_stream")


def synchronize() -> None:
    if load_library().csb_stream_synchronize() != 0:
        _err("csb_stream_synchronize")


def launch_count(reset: bool = False) -> int:
    return int(load_library().csb_launch_count(1 if reset else 0))


class PComm:
    """Peer-memory communicator (csb_pcomm_t): ghost exchange and global sums between
    the GPUs of one box without MPI/NCCL calls.  Build it with `from_handles` (one
    process per GPU) or `in_group` (ranks as threads of one process)."""

    def __init__(self, handle, n_ranks, rank):
        self._h = handle
        self.n_ranks = n_ranks
        self.rank = rank

    @property
    def handle(self):
        return C.c_void_p(self._h)

    @staticmethod
    def default_sizes(n_ghost, n_levels=25):
        """(arena_bytes, window_bytes) large enough for a hierarchy whose finest level
        has n_ghost ghost rows (ghost buffers shrink ~2x per level)."""
        # + room for the merge boundary (merged right-hand side of the first merged level)
        return 64 * (n_ghost + 64) + 8192 * n_levels + (16 << 20), 32 * (n_ghost + 64) + (1 << 20)

    @classmethod
    def from_handles(cls, n_ranks, rank, arena_bytes, window_bytes, all_gather):
        """One process per GPU.  `all_gather(bytes64) -> list of n_ranks bytes objects`
        exchanges the cudaIpc handles (torch.distributed.all_gather_object, MPI_Allgather)."""
        L = load_library()
        h = L.csb_pcomm_create(n_ranks, rank, arena_bytes, window_bytes)
        if not h:
            _err("csb_pcomm_create")
        self = cls(h, n_ranks, rank)
        if n_ranks > 1:
            buf = C.create_string_buffer(64)
            if L.csb_pcomm_get_ipc_handle(self.handle, buf) != 0:
                _err("csb_pcomm_get_ipc_handle")
            handles = all_gather(buf.raw)
            blob = b"".join(handles)
            assert len(blob) == 64 * n_ranks
            if L.csb_pcomm_open_peers(self.handle, blob) != 0:
                _err("csb_pcomm_open_peers")
        return self

    @classmethod
    def in_group(cls, group, n_ranks, rank, arena_bytes, window_bytes):
        """Collective over the threads of `group` (an int from `PComm.group_create`)."""
        h = load_library().csb_pcomm_create_in_group(C.c_void_p(group), rank, arena_bytes,
                                                     window_bytes)
        if not h:
            _err("csb_pcomm_create_in_group")
        return cls(h, n_ranks, rank)

    @staticmethod
    def group_create(n_ranks):
        g = load_library().csb_pcomm_group_create(n_ranks)
        if not g:
            raise CsbError("csb_pcomm_group_create failed")
        return g

    @staticmethod
    def group_destroy(group):
        g = C.c_void_p(group)
        load_library().csb_pcomm_group_destroy(C.byref(g))

    def allreduce_sum(self, vals):
        a = np.ascontiguousarray(vals, dtype=np.float64).copy()
        if load_library().csb_pcomm_allreduce_sum(self.handle, _ptr(a), len(a)) != 0:
            _err("csb_pcomm_allreduce_sum")
        return a

    def barrier(self):
        if load_library().csb_pcomm_barrier(self.handle) != 0:
            _err("csb_pcomm_barrier")

    def destroy(self):
        if self._h:
            h = C.c_void_p(self._h)
            load_library().csb_pcomm_destroy(C.byref(h))
            self._h = None


class Halo:
    """cs_halo_t: ghost sections and send lists of one rank (mesh_cartesian.HaloDesc)."""

    def __init__(self, desc, rank, pcomm=None, nccl_comm=None):
        L = load_library()
        ranks = np.ascontiguousarray(desc.c_domain_rank, dtype=np.int32)
        si = np.ascontiguousarray(desc.send_index, dtype=np.int32)
        sl = np.ascontiguousarray(desc.send_list, dtype=np.int32)
        gi = np.ascontiguousarray(desc.index, dtype=np.int32)
        self.n_ghost = int(gi[-1])
        self.n_local = int(desc.n_local_elts)
        self._h = L.csb_halo_create(C.c_void_p(nccl_comm or 0), rank, len(ranks), _ptr(ranks),
                                    self.n_local, _ptr(si), _ptr(sl), _ptr(gi))
        if not self._h:
            _err("csb_halo_create")
        if pcomm is not None:
            if L.csb_halo_set_pcomm(C.c_void_p(self._h), pcomm.handle) != 0:
                _err("csb_halo_set_pcomm")

    @property
    def handle(self):
        return C.c_void_p(self._h)

    def sync_var(self, var_d):
        """cs_halo_sync_var on a device array of n_local + n_ghost doubles (torch tensor)."""
        if load_library().csb_halo_sync_var(self.handle, _ptr(var_d)) != 0:
            _err("csb_halo_sync_var")

    def destroy(self):
        if self._h:
            h = C.c_void_p(self._h)
            load_library().csb_halo_destroy(C.byref(h))
            self._h = None


def _ptr(a):
    """Address of a numpy array, a torch tensor (host or device), an int or None."""
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if isinstance(a, int):
        return C.c_void_p(a)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(f"unsupported buffer type {type(a)}")


def _i32(a):
    return a if not isinstance(a, np.ndarray) else np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return a if not isinstance(a, np.ndarray) else np.ascontiguousarray(a, dtype=np.float64)


class MatrixAssembler:
    """cs_matrix_assembler_t on one rank (MSR target): declare (row, column) global-id pairs,
    compute the structure, create the matrix, add coefficient values."""

    def __init__(self, l_range):
        L = load_library()
        lr = np.array(l_range, dtype=np.uint64)
        self._h = L.csb_matrix_assembler_create(_ptr(lr), 1)
        if not self._h:
            _err("csb_matrix_assembler_create")
        self._mav = None

    @property
    def handle(self):
        return C.c_void_p(self._h)

    def add_g_ids(self, rows, cols):
        r = np.ascontiguousarray(rows, dtype=np.uint64)
        c = np.ascontiguousarray(cols, dtype=np.uint64)
        if load_library().csb_matrix_assembler_add_g_ids(self.handle, len(r), _ptr(r), _ptr(c)) != 0:
            _err("csb_matrix_assembler_add_g_ids")

    def compute(self):
        if load_library().csb_matrix_assembler_compute(self.handle) != 0:
            _err("csb_matrix_assembler_compute")

    def structure(self):
        L = load_library()
        n = L.csb_matrix_assembler_get_n_rows(self.handle)
        nnz = L.csb_matrix_assembler_get_structure(self.handle, None, None)
        if nnz < 0:
            _err("csb_matrix_assembler_get_structure")
        ri = np.zeros(n + 1, dtype=np.int32)
        ci = np.zeros(max(nnz, 1), dtype=np.int32)
        L.csb_matrix_assembler_get_structure(self.handle, _ptr(ri), _ptr(ci))
        return ri, ci[:nnz]

    def create_matrix(self):
        """cs_matrix_create_from_assembler(CS_MATRIX_MSR, ma): values zero"""
        L = load_library()
        h = L.csb_matrix_create_from_assembler(self.handle)
        if not h:
            _err("csb_matrix_create_from_assembler")
        m = Matrix.__new__(Matrix)
        m._h = h
        m.n_rows = m.n_cols_ext = int(L.csb_matrix_get_n_rows(h))
        m.nnz = int(L.csb_matrix_get_n_entries(h)) - m.n_rows
        m._halo = None
        return m

    def values_init(self, m):
        L = load_library()
        self._mav = L.csb_matrix_assembler_values_init(m.handle, self.handle, 1, 1)
        if not self._mav:
            _err("csb_matrix_assembler_values_init")

    def values_add_g(self, rows, cols, vals):
        r = np.ascontiguousarray(rows, dtype=np.uint64)
        c = np.ascontiguousarray(cols, dtype=np.uint64)
        v = _f64(np.asarray(vals))
        if load_library().csb_matrix_assembler_values_add_g(C.c_void_p(self._mav), len(r), _ptr(r),
                                                            _ptr(c), _ptr(v)) != 0:
            _err("csb_matrix_assembler_values_add_g")

    def values_add_declared(self, vals):
        """vals[k] for the k-th declared pair; numpy array or a device tensor"""
        v = _f64(vals) if isinstance(vals, np.ndarray) else vals
        if load_library().csb_matrix_assembler_values_add_declared(C.c_void_p(self._mav),
                                                                   _ptr(v)) != 0:
            _err("csb_matrix_assembler_values_add_declared")

    def values_done(self):
        if load_library().csb_matrix_assembler_values_done(C.c_void_p(self._mav)) != 0:
            _err("csb_matrix_assembler_values_done")

    def values_finalize(self):
        h = C.c_void_p(self._mav)
        r = load_library().csb_matrix_assembler_values_finalize(C.byref(h))
        self._mav = None
        if r != 0:
            _err("csb_matrix_assembler_values_finalize")

    def destroy(self):
        if self._h:
            h = C.c_void_p(self._h)
            load_library().csb_matrix_assembler_destroy(C.byref(h))
            self._h = None


class Matrix:
    """cs_matrix_t in MSR format, resident in HBM."""

    def __init__(self, n_rows, row_index, col_id, d_val, x_val, symmetric=True,
                 n_cols_ext=None):
        L = load_library()
        self.n_rows = int(n_rows)
        self.n_cols_ext = int(n_cols_ext if n_cols_ext is not None else n_rows)
        ri, ci, d, x = _i32(row_index), _i32(col_id), _f64(d_val), _f64(x_val)
        self._h = L.csb_matrix_create_msr(self.n_rows, self.n_cols_ext, _ptr(ri), _ptr(ci),
                                          _ptr(d), _ptr(x), int(bool(symmetric)))
        if not self._h:
            _err("csb_matrix_create_msr")
        self.nnz = int(L.csb_matrix_get_n_entries(self._h)) - self.n_rows
        self._halo = None

    db = 1      # diagonal block size (3 for from_block matrices)

    @classmethod
    def from_block(cls, n_rows, db, row_index, col_id, d_val, x_val, symmetric=False,
                   n_cols_ext=None):
        """MSR with db x db diagonal blocks (d_val: n_rows*db*db) and scalar extra-diagonal
        entries (CS_MATRIX_BLOCK_D); vectors hold db values per row."""
        L = load_library()
        self = cls.__new__(cls)
        self.n_rows = int(n_rows)
        self.db = int(db)
        self.n_cols_ext = int(n_cols_ext if n_cols_ext is not None else n_rows)
        ri, ci, d, x = _i32(row_index), _i32(col_id), _f64(d_val), _f64(x_val)
        assert d.size == self.n_rows * self.db * self.db
        self._h = L.csb_matrix_create_msr_block(self.n_rows, self.n_cols_ext, self.db, _ptr(ri),
                                                _ptr(ci), _ptr(d), _ptr(x), int(bool(symmetric)))
        if not self._h:
            _err("csb_matrix_create_msr_block")
        self.nnz = int(L.csb_matrix_get_n_entries(self._h)) - self.n_rows
        self._halo = None
        return self

    @classmethod
    def from_csr(cls, n_rows, row_index, col_id, val, symmetric=True, n_cols_ext=None):
        """CS_MATRIX_CSR (diagonal stored as an ordinary entry): vector_multiply sums in CSR
        order like _mat_vec_p_l_csr; the solvers use the MSR split made at creation."""
        L = load_library()
        self = cls.__new__(cls)
        self.n_rows = int(n_rows)
        self.n_cols_ext = int(n_cols_ext if n_cols_ext is not None else n_rows)
        ri, ci, v = _i32(row_index), _i32(col_id), _f64(val)
        self._h = L.csb_matrix_create_csr(self.n_rows, self.n_cols_ext, _ptr(ri), _ptr(ci),
                                          _ptr(v), int(bool(symmetric)))
        if not self._h:
            _err("csb_matrix_create_csr")
        self.nnz = int(L.csb_matrix_get_n_entries(self._h)) - self.n_rows
        self._halo = None
        return self

    @classmethod
    def from_native(cls, n_rows, edges, da, xa, n_cols_ext=None):
        """csb_matrix_create_native: the matrix of a caller holding face-based arrays
        (cs_matrix_set_coefficients(edges, da, xa)); symmetric, one value per edge."""
        L = load_library()
        self = cls.__new__(cls)
        self.n_rows = int(n_rows)
        self.n_cols_ext = int(n_cols_ext if n_cols_ext is not None else n_rows)
        e = np.ascontiguousarray(edges, dtype=np.int32).reshape(-1, 2)
        d, x = _f64(np.asarray(da)), _f64(np.asarray(xa))
        self.n_edges = len(e)
        self._h = L.csb_matrix_create_native(self.n_rows, self.n_cols_ext, self.n_edges, _ptr(e),
                                             _ptr(d), _ptr(x), 1)
        if not self._h:
            _err("csb_matrix_create_native")
        self.nnz = int(L.csb_matrix_get_n_entries(self._h)) - self.n_rows
        self._halo = None
        return self

    @classmethod
    def from_system_native(cls, s, da=None, xa=None, with_mesh=True):
        """The same system as from_system, entered through its face arrays."""
        if da is None or xa is None:
            xa = np.zeros(max(s.n_faces, 1))
            up = np.asarray(s.c2f_sgn) > 0
            xa[np.asarray(s.c2f)[up]] = np.asarray(s.x_val)[up]
            da, xa = s.d_val, xa[:s.n_faces]
        m = cls.from_native(s.n_rows, s.face_cells, da, xa)
        if with_mesh and s.n_faces > 0:
            m.set_mesh_association(s.n_faces, None, None, s.cell_cen, s.cell_vol,
                                   s.face_u_normal, s.face_surf)
        return m

    def set_coefficients_native(self, da, xa):
        d, x = _f64(np.asarray(da)), _f64(np.asarray(xa))
        if load_library().csb_matrix_set_coefficients_native(self.handle, _ptr(d), _ptr(x)) != 0:
            _err("csb_matrix_set_coefficients_native")

    @classmethod
    def from_system(cls, s, with_mesh=True, share_host=False):
        """Build from a mesh_cartesian.CubeSystem (host arrays).  share_host: lend the
        system's host arrays to the matrix (csb_matrix_share_host_arrays; s must outlive it)."""
        m = cls(s.n_rows, s.row_index, s.col_id, s.d_val, s.x_val, True,
                n_cols_ext=s.meta.get("n_cols_ext", s.n_rows))
        if with_mesh and s.n_faces > 0:
            m.set_mesh_association(s.n_faces, s.c2f, s.c2f_sgn, s.cell_cen, s.cell_vol,
                                   s.face_u_normal, s.face_surf)
        if share_host:
            m.share_host_arrays(s.col_id, s.d_val, s.x_val)
        return m

    def share_host_arrays(self, col_id, d_val, x_val):
        """csb_matrix_share_host_arrays: the arrays are kept alive with this object."""
        ci, d, xv = _i32(col_id), _f64(d_val), _f64(x_val)
        self._lent = (ci, d, xv)
        if load_library().csb_matrix_share_host_arrays(self.handle, _ptr(ci), _ptr(d), _ptr(xv)) != 0:
            _err("csb_matrix_share_host_arrays")

    @property
    def handle(self):
        return C.c_void_p(self._h)

    def set_coefficients(self, d_val, x_val, symmetric=True):
        d, x = _f64(d_val), _f64(x_val)
        if load_library().csb_matrix_set_coefficients(self.handle, int(bool(symmetric)),
                                                      _ptr(d), _ptr(x)) != 0:
            _err("csb_matrix_set_coefficients")

    def set_mesh_association(self, n_faces, c2f, c2f_sgn, cell_cen, cell_vol, face_u_normal,
                             face_surf):
        c2f = None if c2f is None else _i32(c2f)
        sgn = None if c2f_sgn is None else np.ascontiguousarray(c2f_sgn, dtype=np.int16)
        cen, vol = _f64(cell_cen), _f64(cell_vol)
        fun, fs = _f64(face_u_normal), _f64(face_surf)
        if load_library().csb_matrix_set_mesh_association(self.handle, int(n_faces), _ptr(c2f),
                                                          _ptr(sgn), _ptr(cen), _ptr(vol),
                                                          _ptr(fun), _ptr(fs)) != 0:
            _err("csb_matrix_set_mesh_association")

    def set_boundary_faces(self, b_face_cells):
        bfc = _i32(b_face_cells)
        if load_library().csb_matrix_set_boundary_faces(self.handle, len(bfc), _ptr(bfc)) != 0:
            _err("csb_matrix_set_boundary_faces")

    def compute_coeffs(self, rovsdt, i_visc, b_visc=None, cofbfp=None, idiffp=1, ndircp=1,
                       thetap=1., relaxp=1., iconvp=0, coefbp=None, i_massflux=None,
                       b_massflux=None, xcpp=None):
        """cs_matrix_compute_coeffs (scalar cases): fill the MSR coefficients in HBM from
        face viscosities / mass fluxes, boundary coefficients and the unsteady term."""
        L = load_library()
        L.csb_matrix_compute_coeffs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                C.c_double, C.c_double, C.c_int] + [C.c_void_p] * 8
        args = [a if a is None or not isinstance(a, np.ndarray) else _f64(a)
                for a in (coefbp, cofbfp, rovsdt, i_massflux, b_massflux, i_visc, b_visc, xcpp)]
        if L.csb_matrix_compute_coeffs(self.handle, iconvp, idiffp, ndircp, float(thetap),
                                       float(relaxp), 0 if xcpp is None else 1,
                                       *[_ptr(a) for a in args]) != 0:
            _err("csb_matrix_compute_coeffs")

    def diffusion_potential(self, pvar, i_visc, b_visc=None, cofaf=None, cofbf=None, init=1, inc=1,
                            diverg=None):
        """cs_diffusion_potential without flux reconstruction on this matrix's mesh."""
        L = load_library()
        L.csb_diffusion_potential.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6
        div = np.zeros(self.n_rows) if diverg is None else _f64(np.asarray(diverg)).copy()
        args = [a if a is None or not isinstance(a, np.ndarray) else _f64(a)
                for a in (pvar, cofaf, cofbf, i_visc, b_visc)]
        if L.csb_diffusion_potential(self.handle, int(init), int(inc), *[_ptr(a) for a in args],
                                     _ptr(div)) != 0:
            _err("csb_diffusion_potential")
        return div

    def divergence(self, i_massflux, b_massflux=None, init=1, diverg=None):
        """cs_divergence on this matrix's mesh"""
        L = load_library()
        L.csb_divergence.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 3
        div = np.zeros(self.n_rows) if diverg is None else _f64(np.asarray(diverg)).copy()
        a = [x if x is None or not isinstance(x, np.ndarray) else _f64(x)
             for x in (i_massflux, b_massflux)]
        if L.csb_divergence(self.handle, int(init), *[_ptr(x) for x in a], _ptr(div)) != 0:
            _err("csb_divergence")
        return div

    def face_diffusion_potential(self, n_faces, n_b_faces, pvar, cofaf, cofbf, i_visc, b_visc,
                                 coefa=None, coefb=None, visel=None, ircflu=0, init=1, inc=1,
                                 massflux=None):
        """cs_face_diffusion_potential; returns (i_massflux, b_massflux)"""
        L = load_library()
        L.csb_face_diffusion_potential.argtypes = [C.c_void_p] + [C.c_int] * 3 + [C.c_void_p] * 10
        im = np.zeros(max(n_faces, 1)) if massflux is None else _f64(np.asarray(massflux[0])).copy()
        bm = np.zeros(max(n_b_faces, 1)) if massflux is None else _f64(np.asarray(massflux[1])).copy()
        a = [x if x is None or not isinstance(x, np.ndarray) else _f64(x)
             for x in (pvar, coefa, coefb, cofaf, cofbf, i_visc, b_visc, visel)]
        if L.csb_face_diffusion_potential(self.handle, int(ircflu), int(init), int(inc),
                                          *[_ptr(x) for x in a], _ptr(im), _ptr(bm)) != 0:
            _err("csb_face_diffusion_potential")
        return im[:n_faces], bm[:n_b_faces]

    def set_face_cog(self, i_face_cog, b_face_cog):
        """face centres, for the hydrostatic pressure treatment"""
        L = load_library()
        L.csb_matrix_set_face_cog.argtypes = [C.c_void_p] * 3
        a = [_f64(np.asarray(x)) for x in (i_face_cog, b_face_cog)]
        if L.csb_matrix_set_face_cog(self.handle, *[_ptr(x) for x in a]) != 0:
            _err("csb_matrix_set_face_cog")

    def gradient_hyd(self, imrgra, f_ext, pvar, coefa=None, coefb=None, inc=1, hyd_p_flag=1):
        """cs_gradient_scalar_synced_input with the hydrostatic pressure treatment"""
        L = load_library()
        L.csb_gradient_scalar_hyd.argtypes = [C.c_void_p] + [C.c_int] * 3 + [C.c_void_p] * 5
        grad = np.zeros(3 * self.n_rows)
        a = [x if x is None or not isinstance(x, np.ndarray) else _f64(x)
             for x in (f_ext, pvar, coefa, coefb)]
        if L.csb_gradient_scalar_hyd(self.handle, int(imrgra), int(inc), int(hyd_p_flag),
                                     *[_ptr(x) for x in a], _ptr(grad)) != 0:
            _err("csb_gradient_scalar_hyd")
        return grad.reshape(-1, 3)

    def diffusion_potential_hyd(self, frcxt, pvar, coefa, coefb, cofaf, cofbf, i_visc, b_visc,
                                visel, irecmf=0, init=1, inc=1, iphydp=1, diverg=None):
        """cs_diffusion_potential with flux reconstruction, iphydp / frcxt given"""
        L = load_library()
        L.csb_diffusion_potential_hyd.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] * 10
        div = np.zeros(self.n_rows) if diverg is None else _f64(np.asarray(diverg)).copy()
        a = [x if x is None or not isinstance(x, np.ndarray) else _f64(x)
             for x in (frcxt, pvar, coefa, coefb, cofaf, cofbf, i_visc, b_visc, visel)]
        if L.csb_diffusion_potential_hyd(self.handle, int(irecmf), int(init), int(inc),
                                         int(iphydp), *[_ptr(x) for x in a], _ptr(div)) != 0:
            _err("csb_diffusion_potential_hyd")
        return div

    def face_diffusion_potential_hyd(self, n_faces, n_b_faces, frcxt, pvar, coefa, coefb, cofaf,
                                     cofbf, i_visc, b_visc, visel, ircflu=1, init=1, inc=1,
                                     iphydp=1, massflux=None):
        L = load_library()
        L.csb_face_diffusion_potential_hyd.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] * 11
        im = np.zeros(max(n_faces, 1)) if massflux is None else _f64(np.asarray(massflux[0])).copy()
        bm = np.zeros(max(n_b_faces, 1)) if massflux is None else _f64(np.asarray(massflux[1])).copy()
        a = [x if x is None or not isinstance(x, np.ndarray) else _f64(x)
             for x in (frcxt, pvar, coefa, coefb, cofaf, cofbf, i_visc, b_visc, visel)]
        if L.csb_face_diffusion_potential_hyd(self.handle, int(ircflu), int(init), int(inc),
                                              int(iphydp), *[_ptr(x) for x in a], _ptr(im),
                                              _ptr(bm)) != 0:
            _err("csb_face_diffusion_potential_hyd")
        return im[:n_faces], bm[:n_b_faces]

    def set_face_geometry(self, i_dist, diipf, djjpf, b_face_u_normal, b_dist):
        L = load_library()
        L.csb_matrix_set_face_geometry.argtypes = [C.c_void_p] * 6
        a = [_f64(np.asarray(x)) for x in (i_dist, diipf, djjpf, b_face_u_normal, b_dist)]
        if L.csb_matrix_set_face_geometry(self.handle, *[_ptr(x) for x in a]) != 0:
            _err("csb_matrix_set_face_geometry")

    def set_boundary_iprime(self, diipb):
        """II' of the boundary faces (cs_mesh_quantities_t::diipb); None resets it to zero"""
        L = load_library()
        L.csb_matrix_set_boundary_iprime.argtypes = [C.c_void_p, C.c_void_p]
        d = None if diipb is None else _f64(np.asarray(diipb))
        if L.csb_matrix_set_boundary_iprime(self.handle, _ptr(d)) != 0:
            _err("csb_matrix_set_boundary_iprime")

    def set_green_geometry(self, weight, dofij, b_face_surf, corr_grad_lin=None):
        """quantities of the Green-Gauss reconstruction (imrgra = 4)"""
        L = load_library()
        L.csb_matrix_set_green_geometry.argtypes = [C.c_void_p] * 5
        a = [None if x is None else _f64(np.asarray(x))
             for x in (weight, dofij, b_face_surf, corr_grad_lin)]
        if L.csb_matrix_set_green_geometry(self.handle, *[_ptr(x) for x in a]) != 0:
            _err("csb_matrix_set_green_geometry")

    def set_gradient_type(self, imrgra):
        L = load_library()
        L.csb_matrix_set_gradient_type.argtypes = [C.c_void_p, C.c_int]
        if L.csb_matrix_set_gradient_type(self.handle, int(imrgra)) != 0:
            _err("csb_matrix_set_gradient_type")

    def set_gradient_limiter(self, imligr=-1, climgr=-1., b_rc_clip_factor=1.):
        L = load_library()
        L.csb_matrix_set_gradient_limiter.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        if L.csb_matrix_set_gradient_limiter(self.handle, int(imligr), float(climgr),
                                             float(b_rc_clip_factor)) != 0:
            _err("csb_matrix_set_gradient_limiter")

    def set_gradient_sweeps(self, nswrgr=100, epsrgr=1e-4):
        L = load_library()
        L.csb_matrix_set_gradient_sweeps.argtypes = [C.c_void_p, C.c_int, C.c_double]
        if L.csb_matrix_set_gradient_sweeps(self.handle, int(nswrgr), float(epsrgr)) != 0:
            _err("csb_matrix_set_gradient_sweeps")

    def gradient_sweeps(self):
        L = load_library()
        L.csb_matrix_get_gradient_sweeps.argtypes = [C.c_void_p]
        return int(L.csb_matrix_get_gradient_sweeps(self.handle))

    def gradient(self, imrgra, pvar, coefa=None, coefb=None, inc=1):
        """cs_gradient_scalar_synced_input for imrgra 1 (LSQ) or 4 (Green-Gauss, LSQ face values)"""
        L = load_library()
        L.csb_gradient_scalar.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 4
        grad = np.zeros(3 * self.n_rows)
        a = [x if x is None or not isinstance(x, np.ndarray) else _f64(x) for x in (pvar, coefa, coefb)]
        if L.csb_gradient_scalar(self.handle, int(imrgra), int(inc), *[_ptr(x) for x in a],
                                 _ptr(grad)) != 0:
            _err("csb_gradient_scalar")
        return grad.reshape(-1, 3)

    def gradient_lsq(self, pvar, coefa=None, coefb=None, inc=1):
        """least-squares cell gradient (cs_gradient_scalar_synced_input, CS_GRADIENT_LSQ)"""
        L = load_library()
        L.csb_gradient_scalar_lsq.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
        grad = np.zeros(3 * self.n_rows)
        a = [x if x is None or not isinstance(x, np.ndarray) else _f64(x) for x in (pvar, coefa, coefb)]
        if L.csb_gradient_scalar_lsq(self.handle, int(inc), *[_ptr(x) for x in a], _ptr(grad)) != 0:
            _err("csb_gradient_scalar_lsq")
        return grad.reshape(-1, 3)

    def diffusion_potential_rc(self, pvar, coefa, coefb, cofaf, cofbf, i_visc, b_visc, visel,
                               irecmf=0, init=1, inc=1, diverg=None):
        """cs_diffusion_potential with flux reconstruction (least-squares gradient)"""
        L = load_library()
        L.csb_diffusion_potential_rc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 9
        div = np.zeros(self.n_rows) if diverg is None else _f64(np.asarray(diverg)).copy()
        a = [x if x is None or not isinstance(x, np.ndarray) else _f64(x)
             for x in (pvar, coefa, coefb, cofaf, cofbf, i_visc, b_visc, visel)]
        if L.csb_diffusion_potential_rc(self.handle, int(irecmf), int(init), int(inc),
                                        *[_ptr(x) for x in a], _ptr(div)) != 0:
            _err("csb_diffusion_potential_rc")
        return div

    def compute_coeffs_vector(self, i_visc, b_visc=None, cofbfp=None, coefbp=None, fimp=None,
                              i_massflux=None, b_massflux=None, iconvp=0, idiffp=1, ndircp=1,
                              thetap=1., relaxp=1.):
        """cs_matrix_compute_coeffs<3>: vector variable, 3x3 diagonal blocks."""
        L = load_library()
        L.csb_matrix_compute_coeffs_vector.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                       C.c_double, C.c_double] + [C.c_void_p] * 7
        args = [a if a is None or not isinstance(a, np.ndarray) else _f64(a)
                for a in (coefbp, cofbfp, fimp, i_massflux, b_massflux, i_visc, b_visc)]
        if L.csb_matrix_compute_coeffs_vector(self.handle, iconvp, idiffp, ndircp, float(thetap),
                                              float(relaxp), *[_ptr(a) for a in args]) != 0:
            _err("csb_matrix_compute_coeffs_vector")

    def get_coefficients(self):
        """(d_val, x_val) copied back from HBM."""
        L = load_library()
        d, x = np.empty(self.n_rows * self.db * self.db), np.empty(max(self.nnz, 1))
        L.csb_matrix_get_coefficients.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        if L.csb_matrix_get_coefficients(self.handle, _ptr(d), _ptr(x)) != 0:
            _err("csb_matrix_get_coefficients")
        return d, x[:self.nnz]

    def set_halo(self, halo):
        self._halo = halo
        load_library().csb_matrix_set_halo(self.handle, halo.handle if halo else None)

    def vector_multiply(self, x, y=None, exclude_diag=False):
        """y = A.x (cs_matrix_vector_multiply).  numpy in -> numpy out; torch device
        tensors are used in place."""
        L = load_library()
        fn = L.csb_matrix_vector_multiply_exdiag if exclude_diag else L.csb_matrix_vector_multiply
        if isinstance(x, np.ndarray):
            x = _f64(x)
            if y is None:
                y = np.empty(self.n_rows * self.db, dtype=np.float64)
        elif y is None:
            raise ValueError("y is required with device buffers")
        if fn(self.handle, _ptr(x), _ptr(y)) != 0:
            _err("csb_matrix_vector_multiply")
        return y

    def spmv_bytes(self) -> float:
        return float(load_library().csb_matrix_spmv_bytes(self.handle))

    def destroy(self):
        if self._h:
            h = C.c_void_p(self._h)
            load_library().csb_matrix_destroy(C.byref(h))
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def dot(x, y) -> float:
    x, y = _f64(x), _f64(y)
    return float(load_library().csb_dot(len(x), _ptr(x), _ptr(y)))


def dot_xx(x) -> float:
    x = _f64(x)
    return float(load_library().csb_dot_xx(len(x), _ptr(x)))


def dot_xx_xy_yz(x, y, z):
    x, y, z = _f64(x), _f64(y), _f64(z)
    r = (C.c_double * 3)()
    if load_library().csb_dot_xx_xy_yz(len(x), _ptr(x), _ptr(y), _ptr(z), C.byref(r, 0),
                                       C.byref(r, 8), C.byref(r, 16)) != 0:
        _err("csb_dot_xx_xy_yz")
    return tuple(r)


def host_aggregate(level, row_index, col_id, d_val, x_val, max_aggregation=3, n_blocks=1):
    """Host-side DX aggregation used by Multigrid.setup (no GPU needed)."""
    ri, ci, d, x = _i32(row_index), _i32(col_id), _f64(d_val), _f64(x_val)
    n = len(ri) - 1
    f_c_row = np.empty(max(n, 1), dtype=np.int32)
    nc = load_library().csb_host_aggregate_dx_msr(level, n, _ptr(ri), _ptr(ci), _ptr(d), _ptr(x),
                                                  max_aggregation, n_blocks, _ptr(f_c_row))
    return nc, f_c_row[:n]


def host_coarsen_pairwise(level, row_index, col_id, d_val, x_val, aggregation_limit=4):
    """One CS_GRID_COARSENING_SPD_PW + Galerkin coarsening step on the host (no GPU needed)."""
    ri, ci, d, x = _i32(row_index), _i32(col_id), _f64(d_val), _f64(x_val)
    n = len(ri) - 1
    nnz = int(ri[-1])
    f_c_row = np.full(max(n, 1), -2, dtype=np.int32)
    c_ri = np.zeros(n + 1, dtype=np.int32)
    c_ci = np.zeros(max(nnz, 1), dtype=np.int32)
    c_d = np.zeros(max(n, 1))
    c_x = np.zeros(max(nnz, 1))
    nc = load_library().csb_host_coarsen_pairwise(level, n, _ptr(ri), _ptr(ci), _ptr(d), _ptr(x),
                                                  aggregation_limit, _ptr(f_c_row), _ptr(c_ri),
                                                  _ptr(c_ci), _ptr(c_d), _ptr(c_x))
    c_nnz = int(c_ri[nc]) if nc > 0 else 0
    return dict(n_rows=nc, coarse_row=f_c_row[:n], row_index=c_ri[:nc + 1], col_id=c_ci[:c_nnz],
                d_val=c_d[:nc], x_val=c_x[:c_nnz])


def host_coarse_structure(row_index, col_id, f_c_row, c_n_rows):
    """Host-side coarse MSR structure + coarse face numbering (no GPU needed)."""
    ri, ci, fc = _i32(row_index), _i32(col_id), _i32(f_c_row)
    n = len(ri) - 1
    nnz = int(ri[-1])
    c_ri = np.zeros(c_n_rows + 1, dtype=np.int32)
    c_ci = np.zeros(max(nnz, 1), dtype=np.int32)
    c_cf = np.zeros(max(nnz, 1), dtype=np.int32)
    c_sg = np.zeros(max(nnz, 1), dtype=np.int16)
    nf = C.c_int(0)
    c_nnz = load_library().csb_host_coarse_structure(n, _ptr(ri), _ptr(ci), _ptr(fc), c_n_rows,
                                                     _ptr(c_ri), _ptr(c_ci), _ptr(c_cf),
                                                     _ptr(c_sg), C.byref(nf))
    return dict(row_index=c_ri, col_id=c_ci[:c_nnz], cell_face=c_cf[:c_nnz],
                cell_face_sgn=c_sg[:c_nnz], n_faces=nf.value)


class _SolverBase:
    _solve_name = ""

    def _solve(self, a: Matrix, rhs, precision, r_norm, vx=None, use_vx_ini=False,
               verbosity=0, name="system"):
        L = load_library()
        fn = getattr(L, self._solve_name)
        host = isinstance(rhs, np.ndarray)
        if host:
            rhs = _f64(rhs)
            if vx is None:
                vx = np.zeros(a.n_cols_ext * a.db, dtype=np.float64)
                use_vx_ini = False
            else:
                vx = _f64(vx)
        elif vx is None:
            raise ValueError("vx is required with device buffers")
        n_iter = C.c_int(0)
        residual = C.c_double(0.)
        cvg = fn(C.c_void_p(self._h), name.encode(), a.handle, verbosity, float(precision),
                 float(r_norm), C.byref(n_iter), C.byref(residual), _ptr(rhs),
                 _ptr(vx) if use_vx_ini else None, _ptr(vx), 0, None)
        if cvg == ConvergenceState.BREAKDOWN and L.csb_last_error():
            msg = L.csb_last_error().decode()
            if msg:
                raise CsbError(f"{self._solve_name}: {msg}")
        return dict(cvg=ConvergenceState(cvg), n_iter=n_iter.value, residual=residual.value,
                    vx=vx)


class SlesIt(_SolverBase):
    """cs_sles_it_t: PCG / Jacobi / BiCGStab with Jacobi preconditioning."""

    _solve_name = "csb_sles_it_solve"

    def __init__(self, solver_type, n_max_iter, poly_degree=0, update_stats=False):
        self._h = load_library().csb_sles_it_create(int(solver_type), int(poly_degree),
                                                    int(n_max_iter), int(update_stats))
        if not self._h:
            _err("csb_sles_it_create")

    def set_pcomm(self, pcomm):
        load_library().csb_sles_it_set_pcomm(C.c_void_p(self._h), pcomm.handle)

    def set_fallback_threshold(self, threshold, n_iter_max):
        """cs_sles_it_set_fallback_threshold: fall back to GCR below this convergence state."""
        load_library().csb_sles_it_set_fallback_threshold(C.c_void_p(self._h), int(threshold),
                                                          int(n_iter_max))

    def set_restart_interval(self, interval):
        load_library().csb_sles_it_set_restart_interval(C.c_void_p(self._h), int(interval))

    def used_fallback(self):
        return bool(load_library().csb_sles_it_used_fallback(C.c_void_p(self._h)))

    def transfer_pc(self, mg: "Multigrid"):
        """cs_sles_it_transfer_pc: the solver takes ownership of the multigrid preconditioner
        (FCG created with poly_degree < 0)."""
        h = C.c_void_p(mg._h)
        if load_library().csb_sles_it_transfer_pc(C.c_void_p(self._h), C.byref(h)) != 0:
            _err("csb_sles_it_transfer_pc")
        mg._h = None

    def pc(self):
        """cs_sles_it_get_pc: the transferred multigrid preconditioner (borrowed), or None."""
        L = load_library()
        L.csb_sles_it_get_pc.restype = C.c_void_p
        h = L.csb_sles_it_get_pc(C.c_void_p(self._h))
        if not h:
            return None
        mg = Multigrid.__new__(Multigrid)
        mg._h = h
        mg._borrowed = True
        return mg

    def setup(self, a: Matrix, name="system", verbosity=0):
        load_library().csb_sles_it_setup(C.c_void_p(self._h), name.encode(), a.handle, verbosity)

    def solve(self, a, rhs, precision, r_norm, vx=None, use_vx_ini=False, verbosity=0):
        out = self._solve(a, rhs, precision, r_norm, vx, use_vx_ini, verbosity)
        out["initial_residual"] = float(
            load_library().csb_sles_it_get_last_initial_residual(C.c_void_p(self._h)))
        out["history"] = self.residual_history()
        return out

    def residual_history(self):
        buf = np.zeros(2048)
        n = load_library().csb_sles_it_get_residual_history(C.c_void_p(self._h), _ptr(buf), 2048)
        return buf[:n].copy()

    def free(self):
        load_library().csb_sles_it_free(C.c_void_p(self._h))

    def log(self):
        load_library().csb_sles_it_log(C.c_void_p(self._h), 0)

    def destroy(self):
        if self._h:
            h = C.c_void_p(self._h)
            load_library().csb_sles_it_destroy(C.byref(h))
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Smoother(SlesIt):
    """cs_multigrid_smoother_*: a fixed number of Jacobi-family sweeps, no convergence test."""

    _solve_name = "csb_multigrid_smoother_solve"

    def __init__(self, smoother_type, n_iter, poly_degree=0):
        self._h = load_library().csb_multigrid_smoother_create(int(smoother_type),
                                                               int(poly_degree), int(n_iter))
        if not self._h:
            _err("csb_multigrid_smoother_create")

    def setup(self, a: Matrix, name="system", verbosity=0):
        load_library().csb_multigrid_smoother_setup(C.c_void_p(self._h), name.encode(),
                                                    a.handle, verbosity)

    def solve(self, a, rhs, vx=None, use_vx_ini=False, verbosity=0):
        return self._solve(a, rhs, 0., 1., vx, use_vx_ini, verbosity)


class Multigrid(_SolverBase):
    """cs_multigrid_t, V-cycle (solver mode defaults of cs_multigrid_define);
    ``Multigrid(pc=True)`` = cs_multigrid_pc_create (single cycle, device smoother settings)."""

    _solve_name = "csb_multigrid_solve"

    def __init__(self, pc=False, mg_type=0):
        """mg_type: 0 V-cycle, 1 K-cycle (cs_multigrid_type_t)."""
        L = load_library()
        self._h = L.csb_multigrid_pc_create(mg_type) if pc else L.csb_multigrid_create(mg_type)
        if not self._h:
            _err("csb_multigrid_create")

    def set_solver_options(self, descent=SlesItType.PCG, ascent=SlesItType.PCG,
                           coarse=SlesItType.PCG, n_max_cycles=100, n_max_iter_descent=2,
                           n_max_iter_ascent=10, n_max_iter_coarse=500, poly_degree=(0, 0, 0),
                           precision_mult=(1., 1., 1.)):
        load_library().csb_multigrid_set_solver_options(
            C.c_void_p(self._h), int(descent), int(ascent), int(coarse), n_max_cycles,
            n_max_iter_descent, n_max_iter_ascent, n_max_iter_coarse, *[int(p) for p in poly_degree],
            *[float(p) for p in precision_mult])

    def set_coarsening_options(self, aggregation_limit=3, coarsening_type=0, n_max_levels=25,
                               min_g_rows=30, p0p1_relax=0.95):
        load_library().csb_multigrid_set_coarsening_options(
            C.c_void_p(self._h), aggregation_limit, coarsening_type, n_max_levels, min_g_rows,
            p0p1_relax, 0)

    def set_pcomm(self, pcomm):
        load_library().csb_multigrid_set_pcomm(C.c_void_p(self._h), pcomm.handle)

    def level_halo(self, level):
        """Halo of one level of a distributed hierarchy, or None."""
        L = load_library()
        sizes = np.zeros(3, dtype=np.int32)
        if L.csb_multigrid_get_level_halo(C.c_void_p(self._h), level, _ptr(sizes), None, None,
                                          None, None) != 0:
            return None
        nd, ns = int(sizes[0]), int(sizes[1])
        ranks = np.zeros(max(nd, 1), dtype=np.int32)
        si = np.zeros(nd + 1, dtype=np.int32)
        gi = np.zeros(nd + 1, dtype=np.int32)
        sl = np.zeros(max(ns, 1), dtype=np.int32)
        L.csb_multigrid_get_level_halo(C.c_void_p(self._h), level, _ptr(sizes), _ptr(ranks),
                                       _ptr(si), _ptr(sl), _ptr(gi))
        return dict(ranks=ranks[:nd], send_index=si, send_list=sl[:ns], index=gi)

    def set_merge_options(self, rank_stride, rows_mean_threshold=300, rows_glob_threshold=500):
        """cs_multigrid_set_merge_options: merge coarse levels onto rank 0 (rank_stride >= the
        number of ranks) once a level's mean rows per rank / global rows fall under the
        thresholds."""
        L = load_library()
        L.csb_multigrid_set_merge_options.argtypes = [C.c_void_p, C.c_int, C.c_int,
                                                      C.c_ulonglong]
        L.csb_multigrid_set_merge_options(C.c_void_p(self._h), int(rank_stride),
                                          int(rows_mean_threshold), int(rows_glob_threshold))

    def merge_level(self):
        return load_library().csb_multigrid_get_merge_level(C.c_void_p(self._h))

    def set_aggregation_blocks(self, n_blocks):
        load_library().csb_multigrid_set_aggregation_blocks(C.c_void_p(self._h), int(n_blocks))

    def setup(self, a: Matrix, name="system", verbosity=0):
        L = load_library()
        L.csb_multigrid_setup(C.c_void_p(self._h), name.encode(), a.handle, verbosity)
        if (L.csb_multigrid_get_n_levels(C.c_void_p(self._h)) < 1
                or not L.csb_multigrid_is_setup(C.c_void_p(self._h))):
            _err("csb_multigrid_setup")

    def solve(self, a, rhs, precision, r_norm, vx=None, use_vx_ini=False, verbosity=0):
        out = self._solve(a, rhs, precision, r_norm, vx, use_vx_ini, verbosity)
        L = load_library()
        out["n_cycles"] = L.csb_multigrid_get_n_cycles(C.c_void_p(self._h))
        buf = np.zeros(512)
        n = L.csb_multigrid_get_cycle_residuals(C.c_void_p(self._h), _ptr(buf), 512)
        out["cycle_residuals"] = buf[:n].copy()
        return out

    def n_levels(self):
        return load_library().csb_multigrid_get_n_levels(C.c_void_p(self._h))

    def setup_times(self):
        t = np.zeros(4)
        load_library().csb_multigrid_get_setup_times(C.c_void_p(self._h), _ptr(t))
        return dict(total=t[0], aggregation=t[1], assembly=t[2], transfers=t[3])

    def phase_times(self):
        """Seconds per (level, phase) of the last solve, from the device-side timers of
        the persistent kernel.  Columns: descent, residual, restrict, coarsest, prolong,
        ascent."""
        buf = np.zeros((32, 6))
        n = load_library().csb_multigrid_get_phase_times(C.c_void_p(self._h), _ptr(buf), 32)
        return buf[:n].copy()

    def smoother_times(self):
        """Seconds per (level, kind) inside the PCG smoothers of the last solve.  Columns:
        residual SpMV, SpMV + 3 dots, vector updates."""
        buf = np.zeros((32, 3))
        n = load_library().csb_multigrid_get_smoother_times(C.c_void_p(self._h), _ptr(buf), 32)
        return buf[:n].copy()

    def levels(self):
        """Copy the hierarchy to host arrays (for the parity tests)."""
        L = load_library()
        out = []
        parent_n = 0
        for l in range(self.n_levels()):
            info = np.zeros(3, dtype=np.int32)
            L.csb_multigrid_get_level_info(C.c_void_p(self._h), l, _ptr(info))
            n, nce, nnz = int(info[0]), int(info[1]), int(info[2])
            cr = np.zeros(max(parent_n, 1), dtype=np.int32)
            ri = np.zeros(n + 1, dtype=np.int32)
            ci = np.zeros(max(nnz, 1), dtype=np.int32)
            d = np.zeros(max(n, 1))
            xv = np.zeros(max(nnz, 1))
            if L.csb_multigrid_get_level_arrays(C.c_void_p(self._h), l, _ptr(cr), _ptr(ri),
                                                _ptr(ci), _ptr(d), _ptr(xv)) != 0:
                _err("csb_multigrid_get_level_arrays")
            out.append(dict(n_rows=n, n_cols_ext=nce, nnz=nnz, coarse_row=cr[:parent_n],
                            row_index=ri, col_id=ci[:nnz], d_val=d[:n], x_val=xv[:nnz]))
            parent_n = nce      # the map covers the parent's ghost rows too
        return out

    def free(self):
        load_library().csb_multigrid_free(C.c_void_p(self._h))

    def log(self):
        load_library().csb_multigrid_log(C.c_void_p(self._h), 0)

    def destroy(self):
        if self._h and not getattr(self, "_borrowed", False):
            h = C.c_void_p(self._h)
            load_library().csb_multigrid_destroy(C.byref(h))
        self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
