"""
ctypes binding of the C-ABI in include/dcscn_b200.h (libdcscn_b200.so, built from csrc/).

This is what stands where `self.sess.run(...)` stood in the reference
(DCSCN.py:420, :565, :575): PyTorch tensors are only the device containers whose
raw pointers are handed to the library.  There is NO CPU fallback: if the shared
library is missing or no B200 is present, construction raises.
"""

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# DCSCN_B200_LIB points the binding at another build of the same C-ABI (kernel A/B runs); default: the in-tree library
LIB_PATH = os.environ.get("DCSCN_B200_LIB") or os.path.join(os.path.dirname(_HERE), "csrc", "libdcscn_b200.so")

PRECISION_F16X3 = 0
PRECISION_F16X1 = 1


class DcscnConfig(ctypes.Structure):
    """Mirror of `struct dcscn_config` (include/dcscn_b200.h)."""
    _fields_ = [
        ("struct_size", ctypes.c_int32),
        ("scale", ctypes.c_int32),
        ("layers", ctypes.c_int32),
        ("filters", ctypes.c_int32),
        ("min_filters", ctypes.c_int32),
        ("filters_decay_gamma", ctypes.c_float),
        ("use_nin", ctypes.c_int32),
        ("nin_filters", ctypes.c_int32),
        ("nin_filters2", ctypes.c_int32),
        ("cnn_size", ctypes.c_int32),
        ("reconstruct_layers", ctypes.c_int32),
        ("reconstruct_filters", ctypes.c_int32),
        ("pixel_shuffler_filters", ctypes.c_int32),
        ("depthwise_separable", ctypes.c_int32),
        ("channels", ctypes.c_int32),
        ("dropout_keep", ctypes.c_float),
        ("l2_decay", ctypes.c_float),
        ("clipping_norm", ctypes.c_float),
        ("beta1", ctypes.c_float),
        ("beta2", ctypes.c_float),
        ("epsilon", ctypes.c_float),
        ("device_id", ctypes.c_int32),
        ("precision", ctypes.c_int32),
    ]


EXPORTED_SYMBOLS = [
    "dcscn_create", "dcscn_destroy", "dcscn_last_error", "dcscn_num_params", "dcscn_param_info",
    "dcscn_set_param", "dcscn_get_param", "dcscn_forward", "dcscn_forward_host", "dcscn_bicubic_resize", "dcscn_forward_ensemble", "dcscn_forward_ensemble_host", "dcscn_forward_ensemble_partial", "dcscn_get_activation",
    "dcscn_set_option", "dcscn_get_timings", "dcscn_launch_count", "dcscn_device_bytes",
    "dcscn_train_step", "dcscn_train_step_host", "dcscn_get_grad", "dcscn_get_adam_slot", "dcscn_set_adam_slot", "dcscn_get_adam_step",
    "dcscn_set_adam_step", "dcscn_last_grad_norm",
    "dcscn_patch_store_set", "dcscn_train_step_indexed", "dcscn_patch_gather", "dcscn_dropout_mask", "dcscn_grad_buffer", "dcscn_apply_gradients", "dcscn_apply_gradients_avg", "dcscn_reset_optimizer", "dcscn_umma_probe", "dcscn_graph_replays",
]

_lib = None


class EngineError(RuntimeError):
    pass


def load_library(path=None):
    """dlopen the C-ABI library; raises EngineError (never falls back) when it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.isfile(path):
        raise EngineError("CUDA extension %s not found - build it with `python __graft_entry__.py` "
                          "(or `make -C dcscn-super-resolution_b200/csrc`); there is no CPU fallback" % path)
    lib = ctypes.CDLL(path)
    vp, ci, c64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    fp = ctypes.POINTER(ctypes.c_float)
    lib.dcscn_last_error.restype = ctypes.c_char_p
    lib.dcscn_create.argtypes = [ctypes.POINTER(DcscnConfig), ctypes.POINTER(vp)]
    lib.dcscn_destroy.argtypes = [vp]
    lib.dcscn_num_params.argtypes = [vp]
    lib.dcscn_param_info.argtypes = [vp, ci, ctypes.c_char_p, ci, ctypes.POINTER(c64), ctypes.POINTER(ci)]
    lib.dcscn_set_param.argtypes = [vp, ctypes.c_char_p, fp, c64]
    lib.dcscn_get_param.argtypes = [vp, ctypes.c_char_p, fp, c64]
    lib.dcscn_forward.argtypes = [vp, vp, vp, vp, ci, ci, ci, vp]
    lib.dcscn_forward_host.argtypes = [vp, vp, vp, vp, ci, ci, ci]
    lib.dcscn_bicubic_resize.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, vp]
    lib.dcscn_forward_ensemble.argtypes = [vp, vp, vp, vp, ci, ci, ci, vp]
    lib.dcscn_forward_ensemble_host.argtypes = [vp, vp, vp, vp, ci, ci, ci]
    lib.dcscn_forward_ensemble_partial.argtypes = [vp, vp, vp, vp, ci, ci, ci, vp]
    lib.dcscn_get_activation.argtypes = [vp, ctypes.c_char_p, fp, c64]
    lib.dcscn_set_option.argtypes = [vp, ctypes.c_char_p, c64]
    lib.dcscn_get_timings.argtypes = [vp, fp, ci, ctypes.POINTER(ci), ctypes.c_char_p, ci]
    u32, cf = ctypes.c_uint32, ctypes.c_float
    lib.dcscn_train_step.argtypes = [vp, vp, vp, vp, ci, ci, ci, cf, u32, ci, fp, fp, vp]
    lib.dcscn_train_step_host.argtypes = [vp, vp, vp, vp, ci, ci, ci, cf, u32, ci, fp, fp]
    lib.dcscn_get_grad.argtypes = [vp, ctypes.c_char_p, fp, c64]
    lib.dcscn_get_adam_slot.argtypes = [vp, ctypes.c_char_p, ci, fp, c64]
    lib.dcscn_set_adam_slot.argtypes = [vp, ctypes.c_char_p, ci, fp, c64]
    lib.dcscn_get_adam_step.argtypes = [vp, ctypes.POINTER(c64)]
    lib.dcscn_set_adam_step.argtypes = [vp, c64]
    lib.dcscn_last_grad_norm.argtypes = [vp]
    lib.dcscn_last_grad_norm.restype = cf
    lib.dcscn_dropout_mask.argtypes = [vp, ctypes.c_char_p, u32, ci, ci, ci, ctypes.POINTER(ctypes.c_uint8), c64]
    lib.dcscn_grad_buffer.argtypes = [vp, ctypes.POINTER(fp), ctypes.POINTER(c64)]
    lib.dcscn_apply_gradients.argtypes = [vp, cf, vp]
    lib.dcscn_reset_optimizer.argtypes = [vp]
    i32p = ctypes.POINTER(ctypes.c_int32)
    lib.dcscn_patch_store_set.argtypes = [vp, vp, vp, vp, c64, ci, ci]
    lib.dcscn_train_step_indexed.argtypes = [vp, i32p, ci, cf, cf, u32, ci, fp, fp]
    lib.dcscn_patch_gather.argtypes = [vp, i32p, ci, cf, vp, vp, vp]
    lib.dcscn_apply_gradients_avg.argtypes = [vp, cf, cf, fp, fp, vp]
    lib.dcscn_launch_count.argtypes = [vp]
    lib.dcscn_launch_count.restype = c64
    lib.dcscn_device_bytes.argtypes = [vp]
    lib.dcscn_device_bytes.restype = c64
    lib.dcscn_graph_replays.argtypes = [vp]
    lib.dcscn_graph_replays.restype = c64
    lib.dcscn_umma_probe.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)]
    lib.dcscn_umma_probe.restype = ctypes.c_int
    _lib = lib
    return lib


def make_config(scale=2, layers=12, filters=196, min_filters=48, filters_decay_gamma=1.5, use_nin=True,
                nin_filters=64, nin_filters2=32, cnn_size=3, reconstruct_layers=1, reconstruct_filters=32,
                pixel_shuffler_filters=0, depthwise_separable=False, channels=1, dropout_keep=0.8,
                l2_decay=0.0001, clipping_norm=5.0, beta1=0.9, beta2=0.999, epsilon=1e-8, device_id=0,
                precision=PRECISION_F16X3):
    c = DcscnConfig()
    c.struct_size = ctypes.sizeof(DcscnConfig)
    c.scale, c.layers, c.filters, c.min_filters = scale, layers, filters, min_filters
    c.filters_decay_gamma = filters_decay_gamma
    c.use_nin, c.nin_filters, c.nin_filters2, c.cnn_size = int(use_nin), nin_filters, nin_filters2, cnn_size
    c.reconstruct_layers, c.reconstruct_filters = reconstruct_layers, reconstruct_filters
    c.pixel_shuffler_filters, c.depthwise_separable, c.channels = pixel_shuffler_filters, int(depthwise_separable), channels
    c.dropout_keep, c.l2_decay, c.clipping_norm = dropout_keep, l2_decay, clipping_norm
    c.beta1, c.beta2, c.epsilon = beta1, beta2, epsilon
    c.device_id, c.precision = device_id, precision
    return c


class Engine:
    """One DCSCN graph instance on one GPU."""

    def __init__(self, config):
        self.lib = load_library()
        self.config = config
        self.handle = ctypes.c_void_p()
        self._check(self.lib.dcscn_create(ctypes.byref(config), ctypes.byref(self.handle)))

    def _check(self, rc):
        if rc != 0:
            raise EngineError(self.lib.dcscn_last_error().decode("utf-8", "replace"))

    def close(self):
        if self.handle:
            self.lib.dcscn_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- variables ----
    def param_shapes(self):
        out = {}
        buf = ctypes.create_string_buffer(256)
        dims = (ctypes.c_int64 * 4)()
        nd = ctypes.c_int()
        for i in range(self.lib.dcscn_num_params(self.handle)):
            self._check(self.lib.dcscn_param_info(self.handle, i, buf, 256, dims, ctypes.byref(nd)))
            out[buf.value.decode()] = tuple(int(dims[k]) for k in range(nd.value))
        return out

    def set_param(self, name, array):
        a = np.ascontiguousarray(array, dtype=np.float32)
        self._check(self.lib.dcscn_set_param(self.handle, name.encode(), a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                             a.size))

    def get_param(self, name):
        shape = self.param_shapes()[name]
        a = np.empty(shape, dtype=np.float32)
        self._check(self.lib.dcscn_get_param(self.handle, name.encode(), a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                             a.size))
        return a

    def set_params(self, weights):
        shapes = self.param_shapes()
        for name, shape in shapes.items():
            if name not in weights:
                raise EngineError("checkpoint has no variable '%s'" % name)
            w = np.asarray(weights[name])
            if tuple(w.shape) != tuple(shape):
                raise EngineError("variable '%s': checkpoint shape %s != graph shape %s" % (name, w.shape, shape))
            self.set_param(name, w)

    # ---- compute ----
    def forward(self, x, x2, y=None, stream=None):
        """x [n,h,w,1], x2 [n,s*h,s*w,1]: contiguous fp32 CUDA torch tensors.  Asynchronous."""
        import torch
        n, h, w = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
        s = self.config.scale
        assert x.is_cuda and x2.is_cuda and x.dtype == torch.float32 and x2.dtype == torch.float32
        assert x.is_contiguous() and x2.is_contiguous()
        assert tuple(x2.shape[:3]) == (n, s * h, s * w), "x2 must be [n, scale*h, scale*w, 1]"
        if y is None:
            y = torch.empty((n, s * h, s * w, 1), dtype=torch.float32, device=x.device)
        st = stream if stream is not None else torch.cuda.current_stream(x.device).cuda_stream
        self._check(self.lib.dcscn_forward(self.handle, x.data_ptr(), x2.data_ptr(), y.data_ptr(), n, h, w,
                                           ctypes.c_void_p(st)))
        return y

    def forward_host(self, x, x2, y=None):
        """numpy (or pinned torch CPU) fp32 arrays in, numpy out; H2D + forward + D2H, synchronous.  With x2 = None the
        bicubic up-scale of x (util.resize_image_by_pil, bit-exact) is formed on the device and only x is copied."""
        xa = _host_array(x)
        n, h, w = xa.shape[0], xa.shape[1], xa.shape[2]
        s = self.config.scale
        x2p = None
        if x2 is not None:
            x2a = _host_array(x2)
            assert tuple(x2a.shape[:3]) == (n, s * h, s * w)
            x2p = x2a.ctypes.data
        if y is None:
            y = np.empty((n, s * h, s * w, 1), dtype=np.float32)
        ya = _host_array(y)
        self._check(self.lib.dcscn_forward_host(self.handle, xa.ctypes.data, x2p, ya.ctypes.data, n, h, w))
        return y

    def bicubic_resize(self, src, out_height, out_width, out=None, stream=None):
        """Pillow's bicubic `Image.resize` of float images on the device: src [n,h,w] fp32 CUDA tensor -> [n,oh,ow]."""
        import torch
        n, h, w = int(src.shape[0]), int(src.shape[1]), int(src.shape[2])
        assert src.is_cuda and src.dtype == torch.float32 and src.is_contiguous()
        if out is None:
            out = torch.empty((n, int(out_height), int(out_width)), dtype=torch.float32, device=src.device)
        st = stream if stream is not None else torch.cuda.current_stream(src.device).cuda_stream
        self._check(self.lib.dcscn_bicubic_resize(self.handle, src.data_ptr(), out.data_ptr(), n, h, w, int(out_height),
                                                  int(out_width), ctypes.c_void_p(st)))
        return out

    def forward_ensemble_host(self, x, x2, flips):
        """Self-ensemble of one image on the device (DCSCN.py:547-586): x [h,w,(1)], x2 [s*h,s*w,(1)] float32 ->
        float64 [s*h, s*w, 1] mean of the inverse-transformed outputs of the first `flips` transforms."""
        xa = np.ascontiguousarray(x, dtype=np.float32)
        h, w = xa.shape[:2]
        s = int(self.config.scale)
        x2p = None                      # None: the device forms the bicubic up-scale of x itself (bit-exact Pillow)
        if x2 is not None:
            x2a = np.ascontiguousarray(x2, dtype=np.float32)
            if x2a.shape[:2] != (s * h, s * w):
                raise ValueError("x2 must be [%d,%d], got %s" % (s * h, s * w, x2a.shape[:2]))
            x2p = x2a.ctypes.data
        y = np.empty((s * h, s * w, 1), dtype=np.float64)
        self._check(self.lib.dcscn_forward_ensemble_host(self.handle, xa.ctypes.data, x2p, y.ctypes.data, h, w, int(flips)))
        return y

    def forward_ensemble(self, x, x2, flips, out=None, stream=None):
        """Device-resident self-ensemble of one image: x [h,w] / x2 [s*h,s*w] fp32 CUDA tensors -> float64 [s*h,s*w]."""
        import torch
        h, w = int(x.shape[0]), int(x.shape[1])
        s = int(self.config.scale)
        if out is None:
            out = torch.empty((s * h, s * w), dtype=torch.float64, device=x.device)
        st = stream if stream is not None else torch.cuda.current_stream(x.device).cuda_stream
        self._check(self.lib.dcscn_forward_ensemble(self.handle, x.data_ptr(), x2.data_ptr(), out.data_ptr(), h, w, int(flips),
                                                    ctypes.c_void_p(st)))
        return out

    def forward_ensemble_sharded(self, x, x2, flips, out=None):
        """The same ensemble with the transforms spread over the ranks of the current torch.distributed job (rank r
        takes transforms r, r + world, ...): every rank runs its share as batched forwards, writes the float64 SUM of
        its inverse-transformed outputs, ONE all-reduce (NCCL sum over NVLink) combines them and the mean is taken.
        Every rank returns the full result.  x / x2: the SAME image on every rank (fp32 CUDA tensors)."""
        import torch
        import torch.distributed as dist
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)
        h, w = int(x.shape[0]), int(x.shape[1])
        s = int(self.config.scale)
        if out is None:
            out = torch.empty((s * h, s * w), dtype=torch.float64, device=x.device)
        mask = 0
        for t in range(rank, int(flips), world):
            mask |= 1 << t
        st = torch.cuda.current_stream(x.device).cuda_stream
        if mask:
            self._check(self.lib.dcscn_forward_ensemble_partial(self.handle, x.data_ptr(), x2.data_ptr(), out.data_ptr(), h, w,
                                                                mask, ctypes.c_void_p(st)))
        else:
            out.zero_()          # more ranks than transforms: this rank contributes nothing
        if world > 1:
            dist.all_reduce(out, op=dist.ReduceOp.SUM)
        out.div_(float(flips))
        return out

    # ---- training patches resident in HBM ----
    def set_patch_store(self, lr_u8, bicubic_u8, true_u8):
        """uint8 patch arrays [count, ph, pw(, 1)] / [count, s*ph, s*pw(, 1)] -> device memory, once per data set."""
        a = [np.ascontiguousarray(v, dtype=np.uint8) for v in (lr_u8, bicubic_u8, true_u8)]
        count, ph, pw = a[0].shape[:3]
        s = int(self.config.scale)
        if a[1].shape[:3] != (count, s * ph, s * pw) or a[2].shape[:3] != (count, s * ph, s * pw):
            raise ValueError("patch arrays do not match: %s %s %s at scale %d" % (a[0].shape, a[1].shape, a[2].shape, s))
        self._check(self.lib.dcscn_patch_store_set(self.handle, a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, count, ph, pw))
        self._patch_shape = (ph, pw)

    @staticmethod
    def _index_array(indices, mirror=None):
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        if mirror is not None:
            idx = idx | (np.asarray(mirror, dtype=np.int64).astype(bool).astype(np.int64) << 31)
        return np.ascontiguousarray(idx.astype(np.uint32).view(np.int32))

    def train_step_indexed(self, indices, lr, seed, max_value=255.0, mirror=None, apply_update=True):
        """One optimisation step on the patches `indices` of the device store; returns (image_loss, mse)."""
        idx = self._index_array(indices, mirror)
        loss, mse = ctypes.c_float(), ctypes.c_float()
        self._check(self.lib.dcscn_train_step_indexed(self.handle, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), int(idx.size),
                                                      float(max_value), float(lr), int(seed) & 0xFFFFFFFF,
                                                      int(bool(apply_update)), ctypes.byref(loss), ctypes.byref(mse)))
        return float(loss.value), float(mse.value)

    def gather_patches(self, indices, max_value=255.0, mirror=None):
        """The fp32 mini-batch tensors (x, x2, y) the indexed step feeds the network, copied back to the host."""
        idx = self._index_array(indices, mirror)
        ph, pw = self._patch_shape
        s = int(self.config.scale)
        n = int(idx.size)
        x = np.empty((n, ph, pw, 1), np.float32)
        x2 = np.empty((n, s * ph, s * pw, 1), np.float32)
        y = np.empty((n, s * ph, s * pw, 1), np.float32)
        self._check(self.lib.dcscn_patch_gather(self.handle, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), n, float(max_value),
                                                x.ctypes.data, x2.ctypes.data, y.ctypes.data))
        return x, x2, y

    def train_step_host(self, x, x2, y, lr, seed, apply_update=True):
        """One optimisation step on host fp32 arrays x [n,h,w,1], x2 / y [n,sh,sw,1]; returns (image_loss, mse)."""
        xa, x2a, ya = _host_array(x), _host_array(x2), _host_array(y)
        n, h, w = xa.shape[0], xa.shape[1], xa.shape[2]
        s = self.config.scale
        assert tuple(x2a.shape[:3]) == (n, s * h, s * w) and tuple(ya.shape[:3]) == (n, s * h, s * w)
        loss, mse = ctypes.c_float(), ctypes.c_float()
        self._check(self.lib.dcscn_train_step_host(self.handle, xa.ctypes.data, x2a.ctypes.data, ya.ctypes.data, n, h, w,
                                                   float(lr), int(seed) & 0xFFFFFFFF, int(bool(apply_update)),
                                                   ctypes.byref(loss), ctypes.byref(mse)))
        return float(loss.value), float(mse.value)

    def train_step(self, x, x2, y, lr, seed, apply_update=True, stream=None):
        """Same with contiguous fp32 CUDA torch tensors."""
        import torch
        n, h, w = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
        assert x.is_cuda and x2.is_cuda and y.is_cuda and x.is_contiguous() and x2.is_contiguous() and y.is_contiguous()
        st = stream if stream is not None else torch.cuda.current_stream(x.device).cuda_stream
        loss, mse = ctypes.c_float(), ctypes.c_float()
        self._check(self.lib.dcscn_train_step(self.handle, x.data_ptr(), x2.data_ptr(), y.data_ptr(), n, h, w, float(lr),
                                              int(seed) & 0xFFFFFFFF, int(bool(apply_update)), ctypes.byref(loss),
                                              ctypes.byref(mse), ctypes.c_void_p(st)))
        return float(loss.value), float(mse.value)

    def grad_tensor(self):
        """Zero-copy torch view of the flat device gradient buffer (for torch.distributed.all_reduce over NCCL)."""
        import torch
        ptr, cnt = ctypes.POINTER(ctypes.c_float)(), ctypes.c_int64()
        self._check(self.lib.dcscn_grad_buffer(self.handle, ctypes.byref(ptr), ctypes.byref(cnt)))

        class _Buf:
            __cuda_array_interface__ = {"shape": (int(cnt.value),), "typestr": "<f4", "version": 2,
                                        "data": (ctypes.cast(ptr, ctypes.c_void_p).value, False)}
        return torch.as_tensor(_Buf(), device="cuda:%d" % self.config.device_id)

    def apply_gradients(self, lr, stream=None):
        import torch
        st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        self._check(self.lib.dcscn_apply_gradients(self.handle, float(lr), ctypes.c_void_p(st)))

    def apply_gradients_avg(self, lr, grad_scale, stream=None):
        """After the all-reduce(sum) of `grad_tensor()`: scale to the mean, clip + Adam; returns (image_loss, mse) means."""
        import torch
        st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        loss, mse = ctypes.c_float(), ctypes.c_float()
        self._check(self.lib.dcscn_apply_gradients_avg(self.handle, float(lr), float(grad_scale), ctypes.byref(loss),
                                                       ctypes.byref(mse), ctypes.c_void_p(st)))
        return float(loss.value), float(mse.value)

    def train_step_data_parallel(self, x, x2, y, lr, seed, indices=None, max_value=255.0):
        """One optimisation step with the mini-batch sharded over the ranks of the current torch.distributed job (equal
        shards): local gradients -> ONE flat all-reduce carrying [gradients | loss | mse] -> identical mean, clip and
        Adam on every rank.  Returns the job-wide (image_loss, mse).  With `indices` the rank's shard is taken from the
        device patch store (x, x2, y ignored)."""
        import torch.distributed as dist
        if indices is not None:
            loss, mse = self.train_step_indexed(indices, lr, seed, max_value=max_value, apply_update=False)
        else:
            fn = self.train_step if hasattr(x, "is_cuda") and x.is_cuda else self.train_step_host
            loss, mse = fn(x, x2, y, lr, seed, apply_update=False)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            g = self.grad_tensor()
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
            return self.apply_gradients_avg(lr, 1.0 / dist.get_world_size())
        self.apply_gradients(lr)
        return loss, mse

    def get_grad(self, name):
        a = np.empty(self.param_shapes()[name], dtype=np.float32)
        self._check(self.lib.dcscn_get_grad(self.handle, name.encode(), a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), a.size))
        return a

    def get_adam_slot(self, name, slot):
        a = np.empty(self.param_shapes()[name], dtype=np.float32)
        self._check(self.lib.dcscn_get_adam_slot(self.handle, name.encode(), slot,
                                                 a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), a.size))
        return a

    def set_adam_slot(self, name, slot, value):
        a = np.ascontiguousarray(value, dtype=np.float32)
        self._check(self.lib.dcscn_set_adam_slot(self.handle, name.encode(), slot,
                                                 a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), a.size))

    @property
    def adam_step(self):
        """Number of optimizer updates applied so far (TF stores beta^(t+1) as beta1_power / beta2_power)."""
        t = ctypes.c_int64()
        self._check(self.lib.dcscn_get_adam_step(self.handle, ctypes.byref(t)))
        return int(t.value)

    @adam_step.setter
    def adam_step(self, t):
        self._check(self.lib.dcscn_set_adam_step(self.handle, int(t)))

    def reset_optimizer(self):
        """Adam slots back to zero and update count to 0 (what re-running the initializer does in the reference)."""
        self._check(self.lib.dcscn_reset_optimizer(self.handle))

    @property
    def last_grad_norm(self):
        return float(self.lib.dcscn_last_grad_norm(self.handle))

    def dropout_mask(self, tensor, seed, n, h, w, channels):
        a = np.empty((n, h, w, channels), dtype=np.uint8)
        self._check(self.lib.dcscn_dropout_mask(self.handle, tensor.encode(), int(seed) & 0xFFFFFFFF, n, h, w,
                                                a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), a.size))
        return a

    def get_activation(self, tensor, shape):
        a = np.empty(shape, dtype=np.float32)
        self._check(self.lib.dcscn_get_activation(self.handle, tensor.encode(),
                                                  a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), a.size))
        return a

    def set_option(self, key, value):
        self._check(self.lib.dcscn_set_option(self.handle, key.encode(), int(value)))

    def timings(self):
        """[(launch name, ms)] of the last forward (needs set_option("timing", 1) before it)."""
        ms = (ctypes.c_float * 64)()
        cnt = ctypes.c_int()
        names = ctypes.create_string_buffer(1024)
        self._check(self.lib.dcscn_get_timings(self.handle, ms, 64, ctypes.byref(cnt), names, 1024))
        return list(zip(names.value.decode().split(","), [float(ms[i]) for i in range(cnt.value)]))

    @property
    def launch_count(self):
        return int(self.lib.dcscn_launch_count(self.handle))

    @property
    def graph_replays(self):
        return int(self.lib.dcscn_graph_replays(self.handle))

    @property
    def device_bytes(self):
        return int(self.lib.dcscn_device_bytes(self.handle))


def _host_array(a):
    if isinstance(a, np.ndarray):
        assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
        return a
    # torch CPU tensor (possibly pinned): share memory
    return a.numpy()
