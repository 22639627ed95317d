"""`Engine`: thin Python owner of one libpinn_b200 handle (one GPU).

It plays the role `tf.Session` + the TF graph play in the reference classes
(INF-L2:48-77): it owns the device-resident variables (theta, lambda, Adam
moments, ADMM z/gamma) and runs the fused residual+loss+gradient kernels.  All
arithmetic happens in the CUDA library; this file only moves pointers.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _capi as capi

_PDE = {"burgers": capi.PDE_BURGERS, "euler": capi.PDE_EULER}
_LOSS = {
    "v1": capi.LOSS_V1_INF_L2, "v1_inf_l2": capi.LOSS_V1_INF_L2,
    "v2": capi.LOSS_V2_INF_ADMM, "v2_inf_admm": capi.LOSS_V2_INF_ADMM,
    "v3": capi.LOSS_V3_L1SQ, "v3_l1sq": capi.LOSS_V3_L1SQ,
    "v4": capi.LOSS_V4_MSE, "v4_mse": capi.LOSS_V4_MSE, "euler_mse": capi.LOSS_V4_MSE,
    "v5": capi.LOSS_V5_ADMM, "v5_admm": capi.LOSS_V5_ADMM, "v6_euler_admm": capi.LOSS_V5_ADMM,
}
_PATH = {"auto": capi.PATH_AUTO, "generic": capi.PATH_GENERIC, "fused": capi.PATH_FUSED, "tensor": capi.PATH_TENSOR}


def _is_torch_tensor(a) -> bool:
    return type(a).__module__.startswith("torch") and hasattr(a, "data_ptr")


class _DeviceView:
    """__cuda_array_interface__ holder so torch can wrap a library-owned device buffer."""

    def __init__(self, ptr: int, n: int, owner):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        self._owner = owner


class Engine:
    def __init__(self, layers: Sequence[int], lb, ub, pde: str = "burgers", loss: str = "v4",
                 lambda1: float = 1.0, lambda2: float = 0.0, rho: float = 1.0, trainable_lambda: bool = False,
                 device: int = 0, path: str = "auto"):
        cfg = capi.PinnConfig()
        cfg.abi_version = capi.PINN_B200_ABI_VERSION
        cfg.n_layers = len(layers)
        for i, w in enumerate(layers):
            cfg.layers[i] = int(w)
        cfg.pde = _PDE[pde]
        cfg.loss = _LOSS[loss]
        lb = np.asarray(lb, np.float64).ravel()
        ub = np.asarray(ub, np.float64).ravel()
        cfg.lb[0], cfg.lb[1] = float(lb[0]), float(lb[1])
        cfg.ub[0], cfg.ub[1] = float(ub[0]), float(ub[1])
        cfg.lambda1, cfg.lambda2, cfg.rho = float(lambda1), float(lambda2), float(rho)
        cfg.trainable_lambda = int(bool(trainable_lambda))
        cfg.device = int(device)
        cfg.path = _PATH[path]
        self.layers = [int(w) for w in layers]
        self.device = int(device)
        self.n_out = self.layers[-1]
        self.n_res = 1 if pde == "burgers" else 3
        self.trainable_lambda = bool(trainable_lambda)
        self.loss_kind = loss
        self._h = C.c_void_p()
        capi.check(capi.lib.pinn_create(C.byref(cfg), C.byref(self._h)), None, "pinn_create")
        n = C.c_int64()
        capi.check(capi.lib.pinn_num_params(self._h, C.byref(n)), self._h, "pinn_num_params")
        self.num_params = int(n.value)
        capi.check(capi.lib.pinn_packed_len(self._h, C.byref(n)), self._h, "pinn_packed_len")
        self.packed_len = int(n.value)
        self.n_f = 0
        self._keep = []  # borrowed device tensors that must outlive the handle's use of them

    # ---- lifetime ----
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            capi.lib.pinn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        capi.check(rc, self._h, what)

    def use_torch_stream(self):
        """Issue all work on torch's current stream for this device (plumbing only)."""
        import torch
        s = torch.cuda.current_stream(self.device).cuda_stream
        self._ck(capi.lib.pinn_set_stream(self._h, C.c_void_p(s)), "pinn_set_stream")

    def set_stream(self, cuda_stream: int):
        self._ck(capi.lib.pinn_set_stream(self._h, C.c_void_p(cuda_stream)), "pinn_set_stream")

    def synchronize(self):
        self._ck(capi.lib.pinn_synchronize(self._h), "pinn_synchronize")

    @property
    def kernel_path(self) -> str:
        p = C.c_int32()
        self._ck(capi.lib.pinn_kernel_path(self._h, C.byref(p)), "pinn_kernel_path")
        return {capi.PATH_GENERIC: "generic", capi.PATH_FUSED: "fused", capi.PATH_TENSOR: "tensor"}[p.value]

    @property
    def launch_count(self) -> int:
        n = C.c_int64()
        self._ck(capi.lib.pinn_launch_count(self._h, C.byref(n)), "pinn_launch_count")
        return int(n.value)

    # ---- variables ----
    @staticmethod
    def _f32(a) -> np.ndarray:
        return np.ascontiguousarray(np.asarray(a, dtype=np.float64).astype(np.float32)) \
            if np.asarray(a).dtype != np.float32 else np.ascontiguousarray(a)

    def set_params(self, theta):
        th = self._f32(theta).ravel()
        if th.size != self.num_params:
            raise ValueError(f"expected {self.num_params} parameters, got {th.size}")
        self._ck(capi.lib.pinn_set_params(self._h, th.ctypes.data_as(C.c_void_p), 0), "pinn_set_params")

    def get_params(self) -> np.ndarray:
        out = np.empty(self.num_params, np.float32)
        self._ck(capi.lib.pinn_get_params(self._h, out.ctypes.data_as(C.c_void_p), 0), "pinn_get_params")
        return out

    def set_lambda(self, l1: float, l2: float):
        self._ck(capi.lib.pinn_set_lambda(self._h, float(l1), float(l2)), "pinn_set_lambda")

    def get_lambda(self) -> Tuple[float, float]:
        a, b = C.c_float(), C.c_float()
        self._ck(capi.lib.pinn_get_lambda(self._h, C.byref(a), C.byref(b)), "pinn_get_lambda")
        return float(a.value), float(b.value)

    # ---- feeds ----
    def set_data(self, X_u, u):
        Xu = self._f32(X_u).reshape(-1, 2)
        uu = self._f32(u).reshape(Xu.shape[0], self.n_out)
        self._ck(capi.lib.pinn_set_data(self._h, Xu.ctypes.data_as(C.c_void_p), uu.ctypes.data_as(C.c_void_p),
                                        Xu.shape[0], 0), "pinn_set_data")

    def set_data_weight(self, w: float):
        self._ck(capi.lib.pinn_set_data_weight(self._h, float(w)), "pinn_set_data_weight")

    def set_collocation(self, X_f, nf_global: int = 0):
        """X_f: [N_f,2] host array (float64 is cast like a TF feed) or a float32 CUDA tensor (borrowed)."""
        if _is_torch_tensor(X_f):
            import torch
            if X_f.dtype != torch.float32 or not X_f.is_cuda or not X_f.is_contiguous():
                raise ValueError("device collocation points must be a contiguous float32 CUDA tensor [N,2]")
            self._keep = [X_f]
            n = X_f.shape[0]
            self._ck(capi.lib.pinn_set_collocation(self._h, C.c_void_p(X_f.data_ptr()), n, int(nf_global), 1),
                     "pinn_set_collocation")
        else:
            Xf = self._f32(X_f).reshape(-1, 2)
            n = Xf.shape[0]
            self._ck(capi.lib.pinn_set_collocation(self._h, Xf.ctypes.data_as(C.c_void_p), n, int(nf_global), 0),
                     "pinn_set_collocation")
            # (host memory has been consumed when the call returns: small batches are copied into a pinned staging ring of the
            #  handle, larger ones staged by the runtime -- no synchronisation, the host prepares the next batch meanwhile)
        self.n_f = int(n)

    def set_collocation_ptr(self, ptr: int, n_f: int, nf_global: int = 0, on_device: bool = True):
        """Raw pointer hand-off (pinned host staging buffers, foreign device memory)."""
        self._ck(capi.lib.pinn_set_collocation(self._h, C.c_void_p(ptr), int(n_f), int(nf_global), int(on_device)),
                 "pinn_set_collocation")
        self.n_f = int(n_f)

    def feed_collocation(self, X_f_host, nf_global: int = 0):
        """The per-step feed_dict of the training loop (INF-L2:127-135): X_f_host is a float32 [N,2] HOST
        tensor / array (pinned memory makes the copy asynchronous).  Returns at once; the next loss_grad_device /
        adam_steps consumes the points chunk by chunk while the rest is still on the bus.  The caller keeps the
        buffer untouched until the step's result has been read back."""
        if _is_torch_tensor(X_f_host):
            import torch
            if X_f_host.is_cuda or X_f_host.dtype != torch.float32 or not X_f_host.is_contiguous():
                raise ValueError("feed_collocation needs a contiguous float32 host tensor [N,2]")
            ptr, n = X_f_host.data_ptr(), X_f_host.shape[0]
        else:
            if X_f_host.dtype != np.float32 or not X_f_host.flags["C_CONTIGUOUS"]:
                raise ValueError("feed_collocation needs a C-contiguous float32 array [N,2]")
            ptr, n = X_f_host.ctypes.data, X_f_host.shape[0]
        self._keep = [X_f_host]
        self._ck(capi.lib.pinn_feed_collocation(self._h, C.c_void_p(ptr), int(n), int(nf_global)), "pinn_feed_collocation")
        self.n_f = int(n)

    # ---- peer-memory exchange group (data-parallel sum inside the reduction kernel) ----
    COMM_HANDLE_BYTES = 64

    def comm_export(self) -> bytes:
        """CUDA IPC handle of this engine's receive buffer (to be gathered over all ranks)."""
        buf = C.create_string_buffer(self.COMM_HANDLE_BYTES)
        self._ck(capi.lib.pinn_comm_export(self._h, buf), "pinn_comm_export")
        return buf.raw

    def comm_attach(self, rank: int, world: int, handles):
        """handles: the comm_export() bytes of ranks 0..world-1 in rank order."""
        table = b"".join(handles)
        if len(table) != world * self.COMM_HANDLE_BYTES:
            raise ValueError("need one %d-byte handle per rank" % self.COMM_HANDLE_BYTES)
        self._ck(capi.lib.pinn_comm_attach(self._h, int(rank), int(world), C.c_char_p(table)), "pinn_comm_attach")
        self.comm_attached = True

    def comm_detach(self):
        self._ck(capi.lib.pinn_comm_detach(self._h), "pinn_comm_detach")
        self.comm_attached = False

    def comm_status(self):
        a, hg = C.c_int32(), C.c_int32()
        self._ck(capi.lib.pinn_comm_status(self._h, C.byref(a), C.byref(hg)), "pinn_comm_status")
        return bool(a.value), bool(hg.value)

    def sample_collocation(self, seed: int, first_index: int, n_f: int, nf_global: int = 0):
        self._ck(capi.lib.pinn_sample_collocation(self._h, int(seed), int(first_index), int(n_f), int(nf_global)),
                 "pinn_sample_collocation")
        self.n_f = int(n_f)

    def sample_lhs(self, seed: int, n_f: int, first_index: int = 0, n_design: int = 0, nf_global: int = 0):
        """`lb + (ub - lb) * lhs(2, N_f)` on the device (INF-L2:183): points [first_index, first_index + n_f) of a Latin
        hypercube design of n_design points (default n_f)."""
        self._ck(capi.lib.pinn_sample_lhs(self._h, int(seed), int(first_index), int(n_f), int(n_design), int(nf_global)),
                 "pinn_sample_lhs")
        self.n_f = int(n_f)

    def get_collocation(self) -> np.ndarray:
        out = np.empty((self.n_f, 2), np.float32)
        self._ck(capi.lib.pinn_get_collocation(self._h, out.ctypes.data_as(C.c_void_p), 0), "pinn_get_collocation")
        return out

    def get_collocation_device(self, out):
        """Copies the current collocation batch into `out`, a float32 [n_f, 2] torch tensor on this engine's GPU."""
        assert out.is_cuda and out.is_contiguous() and out.numel() == 2 * self.n_f
        self._ck(capi.lib.pinn_get_collocation(self._h, C.c_void_p(out.data_ptr()), 1), "pinn_get_collocation")
        self.synchronize()
        return out

    # ---- hot path ----
    def loss_grad(self, want_grad: bool = True) -> Tuple[float, Optional[np.ndarray]]:
        """One sess.run([loss, grads]): residual + loss + full parameter gradient."""
        loss = C.c_double()
        n = self.num_params + (2 if self.trainable_lambda else 0)
        g = np.empty(n, np.float32) if want_grad else None
        self._ck(capi.lib.pinn_loss_grad(self._h, C.byref(loss), g.ctypes.data_as(C.c_void_p) if want_grad else None),
                 "pinn_loss_grad")
        return float(loss.value), g

    def loss_grad_device(self):
        self._ck(capi.lib.pinn_loss_grad_device(self._h), "pinn_loss_grad_device")

    def l1_pass1(self) -> int:
        p = C.c_void_p()
        self._ck(capi.lib.pinn_l1_pass1(self._h, C.byref(p)), "pinn_l1_pass1")
        return int(p.value)

    def packed_ptr(self) -> int:
        p = C.c_void_p()
        self._ck(capi.lib.pinn_packed_ptr(self._h, C.byref(p)), "pinn_packed_ptr")
        return int(p.value)

    def packed_tensor(self):
        """The packed [grad | dlambda | partial sums] device buffer as a torch tensor (no copy)."""
        import torch
        return torch.as_tensor(_DeviceView(self.packed_ptr(), self.packed_len, self), device=f"cuda:{self.device}")

    def device_view(self, ptr: int, n: int):
        import torch
        return torch.as_tensor(_DeviceView(ptr, n, self), device=f"cuda:{self.device}")

    def loss_value(self) -> float:
        loss = C.c_double()
        self._ck(capi.lib.pinn_loss_value(self._h, C.byref(loss)), "pinn_loss_value")
        return float(loss.value)

    def admm_misfit(self) -> float:
        """mean |f - z| of the last loss_value() pass: the reference's `admm_misfit` (AB-ADMM:60, "r(w) - z" at :232)."""
        v = C.c_double()
        self._ck(capi.lib.pinn_admm_misfit(self._h, C.byref(v)), "pinn_admm_misfit")
        return float(v.value)

    def loss_from_packed(self, sums, nf_global: int, loss_kind: str) -> float:
        """Assemble the scalar loss from the (all-reduced) partial sums of the packed vector."""
        s = [float(v) for v in sums]
        if _LOSS[loss_kind] == capi.LOSS_V3_L1SQ:
            return s[capi.SUM_DATA] + s[capi.SUM_ABSF] ** 2 / float(nf_global)
        return s[capi.SUM_DATA] + s[capi.SUM_RES]

    # ---- Adam ----
    def adam_config(self, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
        self._ck(capi.lib.pinn_adam_config(self._h, lr, beta1, beta2, eps), "pinn_adam_config")

    def adam_reset(self):
        self._ck(capi.lib.pinn_adam_reset(self._h), "pinn_adam_reset")

    def adam_apply(self):
        self._ck(capi.lib.pinn_adam_apply(self._h), "pinn_adam_apply")

    def adam_steps(self, n: int):
        self._ck(capi.lib.pinn_adam_steps(self._h, int(n)), "pinn_adam_steps")

    # ---- predict ----
    def predict(self, X, want_f: bool = True):
        Xs = self._f32(X).reshape(-1, 2)
        n = Xs.shape[0]
        u = np.empty((n, self.n_out), np.float32)
        f = np.empty((n, self.n_res), np.float32) if want_f else None
        self._ck(capi.lib.pinn_predict(self._h, Xs.ctypes.data_as(C.c_void_p), n, u.ctypes.data_as(C.c_void_p),
                                       f.ctypes.data_as(C.c_void_p) if want_f else None, 0), "pinn_predict")
        return u, f

    # ---- measurement hooks ----
    def kernel_timing(self, enable: bool):
        self._ck(capi.lib.pinn_kernel_timing(self._h, int(enable)), "pinn_kernel_timing")

    def kernel_time(self):
        """(summed ms, launches) of the dominant residual kernel since kernel_timing(True)."""
        ms, n = C.c_double(), C.c_int64()
        self._ck(capi.lib.pinn_kernel_time(self._h, C.byref(ms), C.byref(n)), "pinn_kernel_time")
        return float(ms.value), int(n.value)

    @staticmethod
    def measure_fma_peak(device: int = 0) -> float:
        tf = C.c_double()
        capi.check(capi.lib.pinn_measure_fma_peak(int(device), C.byref(tf)), None, "pinn_measure_fma_peak")
        return float(tf.value)

    # ---- ADMM ----
    def admm_init(self):
        self._ck(capi.lib.pinn_admm_init(self._h), "pinn_admm_init")

    def admm_update(self, inf_admm_quirk: bool = False):
        self._ck(capi.lib.pinn_admm_update(self._h, int(inf_admm_quirk)), "pinn_admm_update")

    def admm_adam_step(self, inf_admm_quirk: bool = False):
        """z/gamma update of the closing epoch + Adam step of the next one in a single training pass (both evaluate
        the same residuals; AB-ADMM:225-226 then :213).  Same bits as admm_update() followed by adam_steps(1)."""
        self._ck(capi.lib.pinn_admm_adam_step(self._h, int(inf_admm_quirk)), "pinn_admm_adam_step")

    def resampled_epochs(self, n_epochs: int, admm: bool, pending: bool, seed: int, first_batch: int, n_f: int, nf_global: int = 0):
        """n_epochs x (Adam step on the current batch [with the owed z/gamma update folded in], new device-sampled batch):
        the Dialect-B batch loops (AB-ADMM:211-226, EUL:227-242) without a host round trip per epoch."""
        self._ck(capi.lib.pinn_resampled_epochs(self._h, int(n_epochs), int(admm), int(pending), int(seed), int(first_batch),
                                               int(n_f), int(nf_global)), "pinn_resampled_epochs")
        self.n_f = int(n_f)

    def admm_state(self):
        z = np.empty((self.n_f, self.n_res), np.float32)
        g = np.empty((self.n_f, self.n_res), np.float32)
        self._ck(capi.lib.pinn_admm_get_state(self._h, z.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p), 0),
                 "pinn_admm_get_state")
        return z, g

    def admm_set_state(self, z, gamma):
        zz = np.ascontiguousarray(np.asarray(z, np.float32)).reshape(self.n_f, self.n_res)
        gg = np.ascontiguousarray(np.asarray(gamma, np.float32)).reshape(self.n_f, self.n_res)
        self._ck(capi.lib.pinn_admm_set_state(self._h, zz.ctypes.data_as(C.c_void_p), gg.ctypes.data_as(C.c_void_p), 0),
                 "pinn_admm_set_state")
