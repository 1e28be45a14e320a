"""ctypes mirror of the C++ front-end API (same names and argument meaning as upstream).

``CImageResizer(res_bits, src_bits, params, fpclass).resizeImage(src, NewWidth, NewHeight, k,
vars)`` drives ``avir::CImageResizer<fpclass>::resizeImage`` of ``include/avir_b200.h``
through ``libavirb200_host.so``; ``resizeImageDevice`` takes CUDA device pointers (e.g. from
torch tensors).  Arrays are numpy, shaped (H, W, C), dtype uint8 / uint16 / float32.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))

FP_DEF, FP_FLOAT4, FP_FLOAT8_DIL = 0, 1, 2
# the same classes composed with upstream's error-diffusion ditherer (CImageResizerDithererErrdINL/DIL)
FP_DEF_ERRD, FP_FLOAT4_ERRD, FP_FLOAT8_DIL_ERRD = 3, 4, 5
_T = {np.dtype(np.uint8): 0, np.dtype(np.uint16): 1, np.dtype(np.float32): 2, np.dtype(np.float64): 3}


# avirb200_plan_set_option() keys (include/avirb200.h, avirb200_option)
OPT_KERNEL_FAMILY, OPT_STREAM_VARIANT_H, OPT_STREAM_VARIANT_V, OPT_HOST_BANDS, OPT_ALL_STREAM_CHAINS, OPT_OVERLAP_HALO = range(6)


class AvirB200Error(RuntimeError):
    pass


_LT = (np.dtype(np.uint8), np.dtype(np.uint16), np.dtype(np.float32))  # CLancIR buffer types


_lib = None
_host = None


def lib():
    """libavirb200.so (C ABI, include/avirb200.h)."""
    global _lib
    if _lib is None:
        path = os.path.join(_PKG, "libavirb200.so")
        if not os.path.exists(path):
            raise AvirB200Error("libavirb200.so is not built: run `python avir_b200/build.py` "
                                "(there is no CPU fallback)")
        _lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
        _lib.avirb200_status_string.restype = C.c_char_p
        _lib.avirb200_last_error.restype = C.c_char_p
    return _lib


def host_lib():
    """libavirb200_host.so (the C++ front-ends behind a C API)."""
    global _host
    if _host is None:
        lib()
        path = os.path.join(_PKG, "libavirb200_host.so")
        if not os.path.exists(path):
            raise AvirB200Error("libavirb200_host.so is not built: run `python avir_b200/build.py`")
        h = C.CDLL(path)
        h.avirb200_host_last_error.restype = C.c_char_p
        h.avirb200_host_set_option.argtypes = [C.c_int, C.c_int]
        call = [C.c_int] * 6
        geom = [C.c_int] * 5 + [C.c_double] * 3 + [C.c_int] * 3
        h.avirb200_host_desc_create.restype = C.c_void_p
        h.avirb200_host_desc_create.argtypes = call + geom + [C.c_void_p]
        h.avirb200_host_desc_get.restype = C.c_void_p
        h.avirb200_host_desc_get.argtypes = [C.c_void_p]
        h.avirb200_host_desc_free.argtypes = [C.c_void_p]
        h.avirb200_host_resize.restype = C.c_int
        h.avirb200_host_resize.argtypes = call + [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                                  C.c_int, C.c_int, C.c_int] + [C.c_double] * 3 + [
                                                      C.c_int] * 3
        h.avirb200_host_resize_device.restype = C.c_int
        h.avirb200_host_resize_device.argtypes = h.avirb200_host_resize.argtypes + [C.c_void_p,
                                                                                    C.c_void_p]
        h.avirb200_host_workspace_bytes.restype = C.c_longlong
        h.avirb200_host_workspace_bytes.argtypes = call + geom
        h.lancirb200_host_resize.restype = C.c_int
        h.lancirb200_host_resize.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                             C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_double, C.c_double, C.c_double, C.c_double,
                                             C.c_double]
        h.lancirb200_host_desc_create.restype = C.c_void_p
        h.lancirb200_host_desc_create.argtypes = [C.c_int] * 7 + [C.c_double] * 5
        h.lancirb200_host_desc_get.restype = C.c_void_p
        h.lancirb200_host_desc_get.argtypes = [C.c_void_p]
        h.lancirb200_host_desc_free.argtypes = [C.c_void_p]
        _host = h
    return _host


def device_count():
    return lib().avirb200_device_count()


def set_option(option, value):
    """Test / tuning hook of the ctypes driver: plan option applied to every plan the front-end
    objects behind host_lib() use from now on (value -1: back to the plan's default)."""
    host_lib().avirb200_host_set_option(option, value)


class CImageResizerVars:
    """Upstream avir.h:2516-2547 (input members)."""

    def __init__(self, ox=0.0, oy=0.0, UseSRGBGamma=False, AlphaIndex=-1, BuildMode=-1):
        self.ox, self.oy = ox, oy
        self.UseSRGBGamma, self.AlphaIndex, self.BuildMode = UseSRGBGamma, AlphaIndex, BuildMode


class CImageResizer:
    """avir::CImageResizer<fpclass>(aResBitDepth, aSrcBitDepth, aParams), avir.h:4630."""

    def __init__(self, aResBitDepth=8, aSrcBitDepth=0, aParams=0, fpclass=FP_DEF):
        self.res_bits, self.src_bits, self.params, self.fpclass = (aResBitDepth, aSrcBitDepth,
                                                                  aParams, fpclass)

    def _call(self, tin, tout):
        return (self.fpclass, self.res_bits, self.src_bits, self.params, tin, tout)

    @staticmethod
    def _vars(v):
        v = v or CImageResizerVars()
        return v.ox, v.oy, int(v.UseSRGBGamma), v.AlphaIndex, v.BuildMode

    def resizeImage(self, SrcBuf, NewWidth, NewHeight, k=0.0, aVars=None, out_dtype=None,
                    NewBuf=None):
        """Host buffers in, host buffers out (avir.h:4680-4685)."""
        src = np.ascontiguousarray(SrcBuf)
        sh, sw, ch = src.shape
        out_dtype = np.dtype(out_dtype or src.dtype)
        dst = NewBuf if NewBuf is not None else np.empty((NewHeight, NewWidth, ch), out_dtype)
        ox, oy, g, a, bm = self._vars(aVars)
        r = host_lib().avirb200_host_resize(*self._call(_T[src.dtype], _T[out_dtype]),
                                            src.ctypes.data, sw, sh, 0, dst.ctypes.data, NewWidth,
                                            NewHeight, ch, k, ox, oy, g, a, bm)
        if r != 0:
            raise AvirB200Error(host_lib().avirb200_host_last_error().decode())
        return dst

    def workspaceBytes(self, src_shape, in_dtype, NewWidth, NewHeight, out_dtype, k=0.0, aVars=None):
        sh, sw, ch = src_shape
        ox, oy, g, a, bm = self._vars(aVars)
        n = host_lib().avirb200_host_workspace_bytes(
            *self._call(_T[np.dtype(in_dtype)], _T[np.dtype(out_dtype)]), sw, sh, NewWidth,
            NewHeight, ch, k, ox, oy, g, a, bm)
        if n < 0:
            raise AvirB200Error(host_lib().avirb200_host_last_error().decode())
        return int(n)

    def resizeImageDevice(self, d_src, src_shape, in_dtype, d_dst, NewWidth, NewHeight, out_dtype,
                          d_workspace, k=0.0, aVars=None, stream=0):
        """Device pointers (ints); asynchronous on `stream` (B200 extension)."""
        sh, sw, ch = src_shape
        ox, oy, g, a, bm = self._vars(aVars)
        r = host_lib().avirb200_host_resize_device(
            *self._call(_T[np.dtype(in_dtype)], _T[np.dtype(out_dtype)]), d_src, sw, sh, 0, d_dst,
            NewWidth, NewHeight, ch, k, ox, oy, g, a, bm, d_workspace, stream)
        if r != 0:
            raise AvirB200Error(host_lib().avirb200_host_last_error().decode())

    def descriptor(self, src_shape, in_dtype, NewWidth, NewHeight, out_dtype, k=0.0, aVars=None):
        """Host-only: the C-ABI plan descriptor resizeImage would hand to the GPU library.
        Returns (handle, desc_ptr, (mode_h, mode_v)); free with free_descriptor(handle)."""
        sh, sw, ch = src_shape
        ox, oy, g, a, bm = self._vars(aVars)
        modes = (C.c_int * 2)()
        h = host_lib().avirb200_host_desc_create(
            *self._call(_T[np.dtype(in_dtype)], _T[np.dtype(out_dtype)]), sw, sh, NewWidth,
            NewHeight, ch, k, ox, oy, g, a, bm, modes)
        if not h:
            raise AvirB200Error(host_lib().avirb200_host_last_error().decode())
        return h, host_lib().avirb200_host_desc_get(h), (modes[0], modes[1])

    @staticmethod
    def free_descriptor(handle):
        host_lib().avirb200_host_desc_free(handle)


class CLancIRParams:
    """Upstream lancir.h:260-307."""

    def __init__(self, SrcSSize=0, NewSSize=0, kx=0.0, ky=0.0, ox=0.0, oy=0.0, la=3.0):
        self.SrcSSize, self.NewSSize, self.kx, self.ky, self.ox, self.oy, self.la = (
            SrcSSize, NewSSize, kx, ky, ox, oy, la)


class CLancIR:
    """avir::CLancIR, lancir.h:311 (1-4 channel images, u8 / u16 / float buffers)."""

    def resizeImage(self, SrcBuf, NewWidth, NewHeight, aParams=None, out_dtype=None):
        src = np.ascontiguousarray(SrcBuf)
        sh, sw, ch = src.shape
        out_dtype = np.dtype(out_dtype or src.dtype)
        p = aParams or CLancIRParams()
        dst = np.empty((NewHeight, NewWidth, ch), out_dtype)
        if src.dtype not in _LT or out_dtype not in _LT:
            raise AvirB200Error("CLancIR: uint8 / uint16 / float32 buffers only")
        r = host_lib().lancirb200_host_resize(_T[src.dtype], _T[out_dtype], src.ctypes.data, sw, sh,
                                              dst.ctypes.data, NewWidth, NewHeight, ch, p.SrcSSize,
                                              p.NewSSize, p.kx, p.ky, p.ox, p.oy, p.la)
        return r, dst
