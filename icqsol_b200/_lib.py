"""ctypes binding of include/icqb.h (the counterpart of
``cdll.LoadLibrary(getSharedLibraryName('icqLaplaceMatricesCpp'))`` at
bem/icqBaseLaplaceSolver.py:26-27, with explicit argtypes/restype)."""
import ctypes
import os
from ctypes import (POINTER, byref, c_char_p, c_double, c_int, c_int64, c_uint64,
                    c_void_p)

from icqsol_b200.util.icqSharedLibraryUtils import getSharedLibraryName

LIB_NAME = 'libicqLaplaceMatricesB200'

ICQB_OK = 0
ICQB_ECUDA = -1
ICQB_EARG = -2
ICQB_EORDER = -3
ICQB_ESTATE = -4
ICQB_ECOMM = -5

c_double_p = POINTER(c_double)
c_int64_p = POINTER(c_int64)

# (name, restype, argtypes) for every symbol include/icqb.h declares
SIGNATURES = [
    ('icqb_create', c_int, [POINTER(c_void_p), c_int]),
    ('icqb_destroy', c_int, [c_void_p]),
    ('icqb_last_error', c_char_p, [c_void_p]),
    ('icqb_device', c_int, [c_void_p]),
    ('icqb_stream', c_uint64, [c_void_p]),
    ('icqb_set_stream', c_int, [c_void_p, c_uint64]),
    ('icqb_synchronize', c_int, [c_void_p]),
    ('icqb_release_cached_memory', c_int, [c_int]),
    ('icqb_set_mesh', c_int, [c_void_p, c_double_p, c_int64, c_int64_p, c_int64]),
    ('icqb_num_triangles', c_int64, [c_void_p]),
    ('icqb_set_rows', c_int, [c_void_p, c_int64, c_int64]),
    ('icqb_get_rows', c_int, [c_void_p, c_int64_p, c_int64_p]),
    ('icqb_set_row_blocks', c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64]),
    ('icqb_area2_dev', c_uint64, [c_void_p]),
    ('icqb_area2_host', c_int, [c_void_p, c_double_p]),
    ('icqb_subdivide_dev', c_int, [c_void_p, c_uint64, c_uint64, c_int64, c_int, c_int, c_uint64,
                                   c_uint64]),
    ('icqb_subdivide_host', c_int, [c_void_p, c_double_p, c_int64, c_int64_p, c_int64, c_int, c_int,
                                    c_double_p, c_int64_p]),
    ('icqb_assemble_dev', c_int, [c_void_p, c_int, c_uint64, c_uint64, c_int64]),
    ('icqb_lower_num_tiles', c_int64, [c_void_p]),
    ('icqb_set_assembly_variant', c_int, [c_void_p, c_int]),
    ('icqb_assembly_stats', c_int, [c_void_p, c_int64_p, c_int64_p]),
    ('icqb_assemble_lower_dev', c_int, [c_void_p, c_int, c_uint64, c_uint64, c_uint64]),
    ('icqb_symv_dev', c_int, [c_void_p, c_uint64, c_uint64, c_uint64, c_int]),
    ('icqb_symv_f32_dev', c_int, [c_void_p, c_uint64, c_uint64, c_uint64, c_int]),
    ('icqb_assemble_host', c_int, [c_void_p, c_int, c_double_p, c_double_p]),
    ('icqb_diagonal_host', c_int, [c_void_p, c_int, c_double_p]),
    ('computeOffDiagonalTermsArrays', c_int, [c_double_p, c_int64, c_int64_p, c_int64, c_double_p]),
    ('icqb_gemv_dev', c_int, [c_void_p, c_uint64, c_int64, c_int64, c_int64, c_uint64, c_uint64,
                              c_uint64, c_uint64]),
    ('icqb_gemv_host', c_int, [c_void_p, c_uint64, c_int64, c_int64, c_int64, c_double_p,
                               c_double_p]),
    ('icqb_cg_solve_dev', c_int, [c_void_p, c_uint64, c_int64, c_int64, c_int, c_uint64, c_uint64,
                                  c_uint64, c_uint64, c_double, c_int, c_int64, c_int64_p,
                                  c_double_p, c_double_p]),
    ('icqb_cg_set_preconditioner', c_int, [c_void_p, c_int]),
    ('icqb_cg_set_cap_tiles', c_int, [c_void_p, c_int64, c_int64]),
    ('icqb_cg_set_block_partition', c_int, [c_void_p, c_int64_p, c_int64]),
    ('icqb_cg_set_block_extents', c_int, [c_void_p, c_int64_p, c_int64_p, c_int64]),
    ('icqb_cg_last_stats', c_int, [c_void_p, c_int64_p, c_int64_p]),
    ('icqb_cg_last_blocks', c_int, [c_void_p, c_int64_p, c_int64_p]),
    ('icqb_cg_set_mixed_precision', c_int, [c_void_p, c_double]),
    ('icqb_cg_last_mixed', c_int, [c_void_p, c_int64_p]),
    ('icqb_lu_local_columns', c_int64, [c_void_p, c_int64]),
    ('icqb_lu_factor_dev', c_int, [c_void_p, c_uint64, c_int64, c_int64, c_uint64, POINTER(c_int)]),
    ('icqb_lu_solve_dev', c_int, [c_void_p, c_uint64, c_int64, c_int64, c_uint64, c_uint64, c_uint64]),
    ('icqb_comm_unique_id', c_int, [c_void_p]),
    ('icqb_comm_init', c_int, [c_void_p, c_void_p, c_int, c_int]),
    ('icqb_comm_share', c_int, [c_void_p, c_void_p]),
    ('icqb_comm_rank', c_int, [c_void_p, POINTER(c_int), POINTER(c_int)]),
    ('icqb_peer_create', c_int, [c_void_p, c_int64, c_void_p]),
    ('icqb_peer_attach', c_int, [c_void_p, c_void_p]),
    ('icqb_last_ms', c_double, [c_void_p, c_int]),
    ('icqb_launch_count', c_int64, [c_void_p]),
    ('icqb_measure_fp64_peak', c_int, [c_void_p, c_double_p]),
    ('icqQuadratureInit', None, [POINTER(c_void_p)]),
    ('icqQuadratureSetObserver', None, [POINTER(c_void_p), c_double_p]),
    ('icqQuadratureGetMaxOrder', c_int, [POINTER(c_void_p)]),
    ('icqQuadratureEvaluate', c_double, [POINTER(c_void_p), c_int, c_double_p, c_double_p,
                                         c_double_p]),
    ('icqQuadratureEvaluateDouble', c_double, [POINTER(c_void_p), c_int] + [c_double_p] * 6),
    ('icqQuadratureDel', None, [POINTER(c_void_p)]),
]

_lib = None


def load():
    """Load the CUDA library (raises RuntimeError when it has not been built)."""
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(getSharedLibraryName(LIB_NAME))
        for name, restype, argtypes in SIGNATURES:
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(ctx, status):
    """Turn a status code into the exception the reference would raise:
    KeyError for a diagonal order outside 1..5 (bem/icqQuadrature1D.py:38),
    RuntimeError otherwise."""
    if status == ICQB_OK:
        return
    msg = load().icqb_last_error(ctx)
    msg = msg.decode() if msg else 'status {0}'.format(status)
    if status == ICQB_EORDER:
        raise KeyError(msg)
    raise RuntimeError('icqb error {0}: {1}'.format(status, msg))


class Context:
    """Owns one ``icqb_ctx`` (one GPU, one rank)."""

    def __init__(self, device=0):
        self.lib = load()
        self.handle = c_void_p()
        st = self.lib.icqb_create(byref(self.handle), int(device))
        if st != ICQB_OK:
            msg = self.lib.icqb_last_error(None)
            raise RuntimeError('icqb_create failed ({0}): {1}'.format(st, msg.decode() if msg else ''))
        self.device = int(device)

    def check(self, status):
        check(self.handle, status)

    def close(self):
        if self.handle:
            self.lib.icqb_destroy(self.handle)
            self.handle = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def last_ms(self, which):
        return float(self.lib.icqb_last_ms(self.handle, int(which)))

    def cg_stats(self):
        """(residual replacements, matrix products enqueued) of the last CG solve."""
        rep, prod = c_int64(0), c_int64(0)
        self.check(self.lib.icqb_cg_last_stats(self.handle, byref(rep), byref(prod)))
        return rep.value, prod.value

    def launch_count(self):
        return int(self.lib.icqb_launch_count(self.handle))


_comm_owner = {}
PEER_EXCHANGE_MAX_N = 262144   # vector length the optional peer-exchange buffers are sized for


def attachCommunicator(ctx, rank, nranks, broadcastBytes):
    """Give ``ctx`` the process-wide NCCL communicator for its device, creating it
    on first use: rank 0 draws the ncclUniqueId, ``broadcastBytes(payload, nbytes, src)``
    (torch.distributed in the host layer) carries it to the other ranks."""
    if nranks <= 1:
        return
    owner = _comm_owner.get(ctx.device)
    if owner is None:
        owner = Context(ctx.device)
        uid = ctypes.create_string_buffer(128)
        if rank == 0:
            owner.check(owner.lib.icqb_comm_unique_id(uid))
        raw = broadcastBytes(uid.raw, 128, 0)
        owner.check(owner.lib.icqb_comm_init(owner.handle, ctypes.create_string_buffer(raw, 128),
                                             int(rank), int(nranks)))
        if os.environ.get('ICQB_PEER_EXCHANGE') == '1':
            # optional (include/icqb.h, icqb_peer_create; validated on 4 and 8 B200, 0.4 % faster than
            # the NCCL all-reduce): exchange buffers for the fused product collective of the solver
            # of preconditioner kind 2; the 64-byte IPC handles are
            # gathered with one broadcast per rank
            handle = ctypes.create_string_buffer(64)
            owner.check(owner.lib.icqb_peer_create(owner.handle, PEER_EXCHANGE_MAX_N, handle))
            gathered = b''.join(broadcastBytes(handle.raw, 64, r) for r in range(int(nranks)))
            owner.check(owner.lib.icqb_peer_attach(
                owner.handle, ctypes.create_string_buffer(gathered, 64 * int(nranks))))
        _comm_owner[ctx.device] = owner
    ctx.check(ctx.lib.icqb_comm_share(ctx.handle, owner.handle))
