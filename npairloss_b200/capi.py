"""ctypes binding of libnpair_b200.so (include/npair_b200.h).  Plumbing only: torch supplies device memory and
streams; every computation happens inside the CUDA library.  There is NO CPU fallback: if the library or a B200
is missing, calls fail loudly."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NPAIR_LIB") or os.path.join(_HERE, "lib", "libnpair_b200.so")   # NPAIR_LIB: tuning builds only

GLOBAL, LOCAL = 0, 1
HARD, EASY, RAND, RELATIVE_HARD, RELATIVE_EASY = 0, 1, 2, 3, 4
PREC_FP32_BF16X3, PREC_BF16, PREC_FP32_FP16X2 = 0, 1, 2
GEMM_TCGEN05, GEMM_SIMT_CHECK = 0, 1

ERRORS = {0: "OK", -1: "E_ARG", -2: "E_CUDA", -3: "E_NCCL", -4: "E_EMPTY_LIST", -5: "E_POS_RANGE", -6: "E_STATE"}


class NpairConfig(C.Structure):
    _fields_ = [("Q", C.c_int32), ("D", C.c_int32), ("world", C.c_int32), ("rank", C.c_int32), ("num_tops", C.c_int32),
                ("margin_ident", C.c_float), ("margin_diff", C.c_float), ("identsn", C.c_float), ("diffsn", C.c_float),
                ("ap_region", C.c_int32), ("ap_method", C.c_int32), ("an_region", C.c_int32), ("an_method", C.c_int32),
                ("sim_precision", C.c_int32), ("gemm_backend", C.c_int32), ("device", C.c_int32), ("bwd_exchange", C.c_int32),
                # ABI 2 extensions (0 = reference behaviour)
                ("global_scope", C.c_int32), ("normalize_input", C.c_int32), ("grad_chunk_cols", C.c_int32), ("flags", C.c_int32)]


FLAG_NO_FUSED_GRAD, FLAG_SIM_1CTA, FLAG_GRAD_1CTA, FLAG_NCCL_RECORDS, FLAG_NCCL_FEATURES, FLAG_LSEL_WARP = 1, 2, 4, 8, 16, 32


EXPORTS = ["npair_config_default", "npair_workspace_bytes", "npair_nccl_unique_id", "npair_create", "npair_create_with_comm",
           "npair_destroy", "npair_forward", "npair_backward", "npair_forward_backward", "npair_forward_gathered", "npair_backward_partial", "npair_bwd_exchange_mode", "npair_row_scalars", "npair_backward_gathered", "npair_profile_enable", "npair_profile_read", "npair_kernel_launches", "npair_util_f64_to_f32", "npair_util_f32_to_f64", "npair_last_error", "npair_version", "npair_debug_read",
           "npair_debug_gemm", "npair_debug_mma_symmetric", "npair_l2normalize_forward", "npair_l2normalize_backward"]

_LIB = None


def kernel_launches() -> int:
    """Cumulative number of CUDA kernels libnpair_b200 has launched in this process (include/npair_b200.h)."""
    f = lib().npair_kernel_launches
    f.restype = C.c_ulonglong
    return int(f())


class NpairError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libnpair_b200: {ERRORS.get(code, code)}: {msg}")
        self.code = code


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp, fp = C.c_void_p, C.POINTER(C.c_float)
        L.npair_config_default.argtypes = [C.POINTER(NpairConfig), C.c_int32, C.c_int32]
        L.npair_config_default.restype = None
        L.npair_workspace_bytes.argtypes = [C.POINTER(NpairConfig)]
        L.npair_workspace_bytes.restype = C.c_size_t
        L.npair_nccl_unique_id.argtypes = [vp]
        L.npair_create.argtypes = [C.POINTER(NpairConfig), vp, C.POINTER(vp)]
        L.npair_create_with_comm.argtypes = [C.POINTER(NpairConfig), vp, C.POINTER(vp)]
        L.npair_destroy.argtypes = [vp]
        L.npair_destroy.restype = None
        L.npair_forward.argtypes = [vp, vp, vp, fp, vp]
        L.npair_backward.argtypes = [vp, C.c_float, vp, vp]
        L.npair_forward_gathered.argtypes = [vp, vp, vp, fp, vp]
        L.npair_backward_partial.argtypes = [vp, C.c_float, vp, vp, vp]
        L.npair_bwd_exchange_mode.argtypes = [vp]
        L.npair_row_scalars.argtypes = [vp, vp, vp]
        L.npair_backward_gathered.argtypes = [vp, C.c_float, vp, vp, vp]
        L.npair_profile_enable.argtypes = [vp, C.c_int]
        L.npair_profile_read.argtypes = [vp, fp]
        L.npair_last_error.argtypes = [vp]
        L.npair_last_error.restype = C.c_char_p
        L.npair_version.restype = C.c_char_p
        L.npair_debug_read.argtypes = [vp, C.c_int, fp, C.c_size_t]
        L.npair_debug_gemm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]
        L.npair_l2normalize_forward.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
        L.npair_l2normalize_backward.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, vp]
        _LIB = L
    return _LIB


def make_config(Q, D, world=1, rank=0, num_tops=5, margin_ident=0.0, margin_diff=0.0, identsn=-1.0, diffsn=-1.0,
                ap_region=LOCAL, ap_method=RAND, an_region=LOCAL, an_method=RAND, sim_precision=PREC_FP32_FP16X2,
                gemm_backend=GEMM_TCGEN05, device=-1, bwd_exchange=0, global_scope=0, normalize_input=0, grad_chunk_cols=0,
                flags=0) -> NpairConfig:
    return NpairConfig(Q, D, world, rank, num_tops, margin_ident, margin_diff, identsn, diffsn, ap_region, ap_method,
                       an_region, an_method, sim_precision, gemm_backend, device, bwd_exchange, global_scope, normalize_input,
                       grad_chunk_cols, flags)


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = lib().npair_nccl_unique_id(buf)
    if rc:
        raise NpairError(rc, lib().npair_last_error(None).decode())
    return buf.raw


class Context:
    """One per rank.  forward()/backward() take torch CUDA tensors (device pointers) and return host scalars."""

    def __init__(self, cfg: NpairConfig, nccl_id: bytes | None = None):
        L = lib()
        self.cfg = cfg
        self._h = C.c_void_p()
        idbuf = C.create_string_buffer(nccl_id, 128) if nccl_id is not None else None
        rc = L.npair_create(C.byref(cfg), idbuf, C.byref(self._h))
        if rc:
            raise NpairError(rc, L.npair_last_error(None).decode())

    def close(self):
        if self._h:
            lib().npair_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise NpairError(rc, lib().npair_last_error(self._h).decode())

    def forward_ptr(self, feat_ptr: int, label_ptr: int, stream: int = 0):
        tops = (C.c_float * 5)()
        self._check(lib().npair_forward(self._h, feat_ptr, label_ptr, tops, stream))
        return [tops[i] for i in range(5)]

    def backward_ptr(self, loss_weight: float, diff_ptr: int, stream: int = 0):
        self._check(lib().npair_backward(self._h, C.c_float(loss_weight), diff_ptr, stream))

    def forward_backward_ptr(self, feat_ptr: int, label_ptr: int, loss_weight: float, diff_ptr: int, stream: int = 0):
        tops = (C.c_float * 5)()
        f = lib().npair_forward_backward
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.POINTER(C.c_float), C.c_void_p]
        self._check(f(self._h, feat_ptr, label_ptr, C.c_float(loss_weight), diff_ptr, tops, stream))
        return [tops[i] for i in range(5)]

    def forward_backward(self, feat, label, loss_weight, diff):
        """npair_forward_backward: both passes, one host synchronisation."""
        import torch
        assert feat.is_cuda and label.is_cuda and diff.is_cuda and feat.is_contiguous() and label.is_contiguous() and diff.is_contiguous()
        assert feat.dtype == torch.float32 and label.dtype == torch.float32 and diff.dtype == torch.float32
        return self.forward_backward_ptr(feat.data_ptr(), label.data_ptr(), loss_weight, diff.data_ptr(), torch.cuda.current_stream().cuda_stream)

    def forward(self, feat, label):
        import torch
        assert feat.is_cuda and label.is_cuda and feat.dtype == torch.float32 and label.dtype == torch.float32
        assert feat.is_contiguous() and label.is_contiguous()
        return self.forward_ptr(feat.data_ptr(), label.data_ptr(), torch.cuda.current_stream().cuda_stream)

    def backward(self, loss_weight, diff):
        import torch
        assert diff.is_cuda and diff.dtype == torch.float32 and diff.is_contiguous()
        self.backward_ptr(loss_weight, diff.data_ptr(), torch.cuda.current_stream().cuda_stream)

    def forward_gathered(self, feat_total, label_total):
        import torch
        assert feat_total.is_cuda and feat_total.dtype == torch.float32 and feat_total.is_contiguous()
        assert label_total.is_cuda and label_total.dtype == torch.float32 and label_total.is_contiguous()
        tops = (C.c_float * 5)()
        self._check(lib().npair_forward_gathered(self._h, feat_total.data_ptr(), label_total.data_ptr(), tops,
                                                 torch.cuda.current_stream().cuda_stream))
        return [tops[i] for i in range(5)]

    def backward_partial(self, loss_weight, local_half, total_half=None):
        import torch
        self._check(lib().npair_backward_partial(self._h, C.c_float(loss_weight), local_half.data_ptr(),
                                                 total_half.data_ptr() if total_half is not None else None,
                                                 torch.cuda.current_stream().cuda_stream))

    def bwd_exchange_mode(self):
        return lib().npair_bwd_exchange_mode(self._h)

    def row_scalars(self, out):
        import torch
        self._check(lib().npair_row_scalars(self._h, out.data_ptr(), torch.cuda.current_stream().cuda_stream))

    def backward_gathered(self, loss_weight, rs_total, diff):
        import torch
        self._check(lib().npair_backward_gathered(self._h, C.c_float(loss_weight), rs_total.data_ptr(), diff.data_ptr(),
                                                  torch.cuda.current_stream().cuda_stream))

    def profile_enable(self, on=True):
        self._check(lib().npair_profile_enable(self._h, 1 if on else 0))

    def profile_read(self):
        ms = (C.c_float * 9)()
        self._check(lib().npair_profile_read(self._h, ms))
        return [ms[i] for i in range(9)]

    def debug_read(self, which: int, n: int):
        import numpy as np
        out = np.zeros(n, dtype=np.float32)
        self._check(lib().npair_debug_read(self._h, which, out.ctypes.data_as(C.POINTER(C.c_float)), n))
        return out


def debug_gemm(precision, backend, A, B):
    """C = A @ B.T through the split-operand GEMM (unit test hook)."""
    import torch
    M, K = A.shape
    Nn = B.shape[0]
    Cout = torch.empty((M, Nn), dtype=torch.float32, device=A.device)
    rc = lib().npair_debug_gemm(precision, backend, M, Nn, K, A.data_ptr(), B.data_ptr(), Cout.data_ptr(),
                                torch.cuda.current_stream().cuda_stream)
    if rc:
        raise NpairError(rc, lib().npair_last_error(None).decode())
    return Cout


def l2normalize_forward(x):
    """y = x / ||x||_2 per row (npair_l2normalize_forward); returns (y, inv_norm) as CUDA tensors."""
    import torch
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2
    y = torch.empty_like(x)
    inv = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    rc = lib().npair_l2normalize_forward(x.data_ptr(), x.shape[0], x.shape[1], y.data_ptr(), inv.data_ptr(), torch.cuda.current_stream().cuda_stream)
    if rc:
        raise NpairError(rc, lib().npair_last_error(None).decode())
    return y, inv


def l2normalize_backward(y, inv, dy):
    import torch
    dx = torch.empty_like(dy)
    rc = lib().npair_l2normalize_backward(y.data_ptr(), inv.data_ptr(), dy.data_ptr(), y.shape[0], y.shape[1], dx.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream)
    if rc:
        raise NpairError(rc, lib().npair_last_error(None).decode())
    return dx
