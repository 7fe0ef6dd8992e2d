"""Host-side engine objects over the C ABI: Context, DevicePoints, DeviceScalars.

This is the layer the reference-facing modules (``verifiable_mpc_b200.ac20.pivot`` etc.) call.  Everything that
touches group elements goes through libvmsm.so on the GPU; Python only marshals integers <-> little-endian bytes.
"""
import ctypes
import itertools
import os

from . import _lib
from ._lib import VmsmError, check

ED_P = 2**255 - 19
ED_L = 2**252 + 27742317777372353535851937790883648493
BN_U = 1868033 ** 3
BN_P = 36 * BN_U ** 4 + 36 * BN_U ** 3 + 24 * BN_U ** 2 + 6 * BN_U + 1
BN_N = 36 * BN_U ** 4 + 36 * BN_U ** 3 + 18 * BN_U ** 2 + 6 * BN_U + 1
ORDERS = {_lib.CURVE_ED25519: ED_L, _lib.CURVE_BN256_G1: BN_N, _lib.CURVE_BN256_G2: BN_N}
WIRE_BYTES = {_lib.CURVE_ED25519: 64, _lib.CURVE_BN256_G1: 64, _lib.CURVE_BN256_G2: 128}


def _buf(data):
    """bytes / bytearray / numpy uint8 array -> (ctypes pointer, keepalive)"""
    if isinstance(data, (bytes, bytearray)):
        arr = (ctypes.c_ubyte * len(data)).from_buffer_copy(data) if isinstance(data, bytes) else (
            ctypes.c_ubyte * len(data)).from_buffer(data)
        return ctypes.cast(arr, ctypes.c_void_p), arr
    if hasattr(data, "ctypes") and hasattr(data, "nbytes"):  # numpy array (C-contiguous uint8 expected)
        if not data.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return ctypes.c_void_p(data.ctypes.data), data
    if data is None:
        return ctypes.c_void_p(0), None
    raise TypeError(f"unsupported buffer type {type(data)}")


_REPEAT_32, _REPEAT_LITTLE = itertools.repeat(32), itertools.repeat("little")


def pack_scalars(xs, order=ED_L):
    """Iterable of Python ints (negative / unreduced allowed, as the reference passes them: pivot.py:119-128,
    compressed_pivot.py:66,134) -> n*32 bytes little-endian, reduced below the group order."""
    if type(xs) is list and xs:
        try:  # a list of residues already in [0, order): no per-element int() and %
            if min(xs) >= 0 and max(xs) < order:
                return b"".join(map(int.to_bytes, xs, _REPEAT_32, _REPEAT_LITTLE))
        except (TypeError, AttributeError):
            pass
    return b"".join((int(x) % order).to_bytes(32, "little") for x in xs)


def pack_points(pts):
    """Iterable of canonical affine (x, y) int pairs -> n*64 bytes."""
    return b"".join(int(x).to_bytes(32, "little") + int(y).to_bytes(32, "little") for x, y in pts)


def unpack_points(raw):
    return [(int.from_bytes(raw[i:i + 32], "little"), int.from_bytes(raw[i + 32:i + 64], "little"))
            for i in range(0, len(raw), 64)]


# BN256 wire forms (include/vmsm.h): G1 = x || y (64 B), G2 = x.re || x.im || y.re || y.im (128 B), identity = zeros;
# host-side points are None (identity), (x, y) for G1 and ((xre, xim), (yre, yim)) for G2.
def pack_points_bn(pts, curve):
    out = []
    for p in pts:
        if p is None:
            out.append(bytes(WIRE_BYTES[curve]))
        elif curve == _lib.CURVE_BN256_G2:
            out.append(b"".join(int(v).to_bytes(32, "little") for v in (p[0][0], p[0][1], p[1][0], p[1][1])))
        else:
            out.append(int(p[0]).to_bytes(32, "little") + int(p[1]).to_bytes(32, "little"))
    return b"".join(out)


def unpack_points_bn(raw, curve):
    wb = WIRE_BYTES[curve]
    out = []
    for i in range(0, len(raw), wb):
        v = [int.from_bytes(raw[i + 32 * k:i + 32 * k + 32], "little") for k in range(wb // 32)]
        if not any(v):
            out.append(None)
        elif curve == _lib.CURVE_BN256_G2:
            out.append(((v[0], v[1]), (v[2], v[3])))
        else:
            out.append((v[0], v[1]))
    return out


def pack_any(pts, curve):
    return pack_points(pts) if curve == _lib.CURVE_ED25519 else pack_points_bn(pts, curve)


def unpack_any(raw, curve):
    return unpack_points(raw) if curve == _lib.CURVE_ED25519 else unpack_points_bn(raw, curve)


class DevicePoints:
    """A device-resident vector of group elements (a generator list ``g`` / ``g_hat``)."""

    def __init__(self, ctx, handle, n, curve):
        self.ctx, self.handle, self.n, self.curve = ctx, handle, n, curve
        self.precomputed = False

    def __len__(self):
        return self.n

    def download(self, off=0, n=None):
        n = self.n - off if n is None else n
        wb = WIRE_BYTES[self.curve]
        out = ctypes.create_string_buffer(max(1, wb * n))
        check(self.ctx.lib.vmsm_points_download(self.ctx.h, self.handle, off, n, out))
        return out.raw[:wb * n]

    def tolist(self, off=0, n=None):
        return unpack_any(self.download(off, n), self.curve)

    def wire_view(self, off=0, n=None):
        """Zero-copy view (ctypes char array over the context's pinned buffer) of the canonical wire bytes of a range;
        valid until the next ``*_view`` / ``text_bytes`` call on this context."""
        n = self.n - off if n is None else n
        ptr = ctypes.c_void_p()
        check(self.ctx.lib.vmsm_points_download_ptr(self.ctx.h, self.handle, off, n, ctypes.byref(ptr)))
        nbytes = n * WIRE_BYTES[self.curve]
        return (ctypes.c_char * nbytes).from_address(ptr.value) if nbytes else b""

    def precompute(self, window_bits=0):
        """Build the table of ``2^(c*w) * P_i`` for a vector of FIXED generators (see ``vmsm_points_precompute``):
        later MSMs on this vector run without doublings.  Dropped again by ``fold``."""
        check(self.ctx.lib.vmsm_points_precompute(self.ctx.h, self.handle, window_bits))
        self.precomputed = True
        return self

    def text_bytes(self, off=0, n=None):
        """``b"[x0, y0, 1], [x1, y1, 1], ..."`` formatted on the device (the inside of repr(list of points))."""
        n = self.n - off if n is None else n
        ptr, ln = ctypes.c_void_p(), ctypes.c_uint64()
        check(self.ctx.lib.vmsm_points_text_ptr(self.ctx.h, self.handle, off, n, ctypes.byref(ptr), ctypes.byref(ln)))
        return ctypes.string_at(ptr, ln.value) if ln.value else b""

    def text_view(self, off=0, n=None):
        """As ``text_bytes`` without the copy: a ctypes char array over the context's pinned buffer, valid until the
        next ``*_view`` / ``text_*`` call on this context (hashlib reads it in place)."""
        n = self.n - off if n is None else n
        ptr, ln = ctypes.c_void_p(), ctypes.c_uint64()
        check(self.ctx.lib.vmsm_points_text_ptr(self.ctx.h, self.handle, off, n, ctypes.byref(ptr), ctypes.byref(ln)))
        return (ctypes.c_char * ln.value).from_address(ptr.value) if ln.value else b""

    def text(self, off=0, n=None):
        return self.text_bytes(off, n).decode("ascii")

    def fold(self, c):
        """In place ``P[j] = c*P[j] + P[half+j]`` (compressed_pivot.py:64); the vector shrinks to half."""
        half = self.n // 2
        cb = (int(c) % ED_L).to_bytes(32, "little")
        check(self.ctx.lib.vmsm_fold(self.ctx.h, self.handle, half, cb))
        self.n = half
        self.precomputed = False
        return self

    def free(self):
        if self.handle and self.ctx.h:
            self.ctx.lib.vmsm_points_free(self.ctx.h, self.handle)
        self.handle = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceScalars:
    def __init__(self, ctx, handle, n):
        self.ctx, self.handle, self.n = ctx, handle, n

    def __len__(self):
        return self.n

    def download(self, off=0, n=None):
        n = self.n - off if n is None else n
        out = ctypes.create_string_buffer(max(1, 32 * n))
        check(self.ctx.lib.vmsm_scalars_download(self.ctx.h, self.handle, off, n, out))
        return out.raw[:32 * n]

    def tolist(self, off=0, n=None):
        raw = self.download(off, n)
        return [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)]

    def fold(self, half, c, mode):
        """In place on [0, 2*half): FOLD_WITNESS v[j] += c*v[half+j]; FOLD_FORM v[j] = c*v[j] + v[half+j] (mod l)."""
        check(self.ctx.lib.vmsm_scalars_fold(self.ctx.h, self.handle, half, (int(c) % ED_L).to_bytes(32, "little"), mode))

    def wire_view(self, off=0, n=None):
        """Zero-copy view of n x 32 little-endian bytes (see DevicePoints.wire_view)."""
        n = self.n - off if n is None else n
        ptr = ctypes.c_void_p()
        check(self.ctx.lib.vmsm_scalars_download_ptr(self.ctx.h, self.handle, off, n, ctypes.byref(ptr)))
        return (ctypes.c_char * (32 * n)).from_address(ptr.value) if n else b""

    def axpy(self, c, src=None, mode=_lib.AXPY_ADD_SCALED, off=0, soff=0, n=None):
        """Element-wise on the device: self += c*src (ADD_SCALED), self = c*self + src (SCALE_ADD), self *= c (SCALE)."""
        n = self.n - off if n is None else n
        check(self.ctx.lib.vmsm_scalars_axpy(self.ctx.h, self.handle, off, src.handle if src is not None else 0, soff, n,
                                             (int(c) % ED_L).to_bytes(32, "little"), mode))

    def text_bytes(self, off=0, n=None, signed=True):
        """b"v0, v1, ..." -- the residues as MPyC prints them (signed representatives unless ``signed`` is False)."""
        n = self.n - off if n is None else n
        ptr, ln = ctypes.c_void_p(), ctypes.c_uint64()
        check(self.ctx.lib.vmsm_scalars_text_ptr(self.ctx.h, self.handle, off, n, 1 if signed else 0,
                                                 ctypes.byref(ptr), ctypes.byref(ln)))
        return ctypes.string_at(ptr, ln.value) if ln.value else b""

    def text_view(self, off=0, n=None, signed=True):
        """As ``text_bytes`` without the copy (see ``DevicePoints.text_view``)."""
        n = self.n - off if n is None else n
        ptr, ln = ctypes.c_void_p(), ctypes.c_uint64()
        check(self.ctx.lib.vmsm_scalars_text_ptr(self.ctx.h, self.handle, off, n, 1 if signed else 0,
                                                 ctypes.byref(ptr), ctypes.byref(ln)))
        return (ctypes.c_char * ln.value).from_address(ptr.value) if ln.value else b""

    def free(self):
        if self.handle and self.ctx.h:
            self.ctx.lib.vmsm_scalars_free(self.ctx.h, self.handle)
        self.handle = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedBuffer:
    """Page-locked host memory exposed as a numpy uint8 array (for the end-to-end H2D path)."""

    def __init__(self, lib, nbytes):
        import numpy as np

        self.lib = lib
        p = ctypes.c_void_p()
        check(lib.vmsm_host_alloc(nbytes, ctypes.byref(p)))
        self.ptr, self.nbytes = p, nbytes
        self.array = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_ubyte)), shape=(nbytes,))

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.vmsm_host_free(self.ptr)
            self.ptr = None


class Context:
    """One CUDA device + one stream.  Calls on a context are serialised by the caller."""

    def __init__(self, device=None):
        self.lib = _lib.load()
        self.h = 0
        if device is None:
            device = int(os.environ.get("VMSM_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        h = ctypes.c_uint64()
        check(self.lib.vmsm_ctx_create(device, ctypes.byref(h)))
        self.h = h.value
        self.device = device

    # -- options / timing
    def set_option(self, key, value):
        check(self.lib.vmsm_ctx_set_option(self.h, key, int(value)))

    def sync(self):
        check(self.lib.vmsm_sync(self.h))

    def timer_start(self):
        check(self.lib.vmsm_timer_start(self.h))

    def timer_stop(self):
        ms = ctypes.c_float()
        check(self.lib.vmsm_timer_stop(self.h, ctypes.byref(ms)))
        return ms.value

    def phase_times(self):
        ms = (ctypes.c_double * len(_lib.PHASES))()
        calls = ctypes.c_uint64()
        check(self.lib.vmsm_phase_times(self.h, ms, ctypes.byref(calls)))
        return dict(zip(_lib.PHASES, list(ms))), calls.value

    def launch_count(self):
        n = ctypes.c_uint64()
        check(self.lib.vmsm_launch_count(self.h, ctypes.byref(n)))
        return n.value

    def pinned(self, nbytes):
        return PinnedBuffer(self.lib, nbytes)

    # -- points
    def upload_points(self, pts, curve=_lib.CURVE_ED25519):
        """``pts``: bytes (n*64, canonical affine LE) or a list of (x, y) ints."""
        raw = pts if isinstance(pts, (bytes, bytearray)) or hasattr(pts, "nbytes") else pack_any(pts, curve)
        nbytes = raw.nbytes if hasattr(raw, "nbytes") else len(raw)
        n = nbytes // WIRE_BYTES[curve]
        p, keep = _buf(raw)
        h = ctypes.c_uint64()
        check(self.lib.vmsm_points_upload(self.h, curve, p, n, ctypes.byref(h)))
        return DevicePoints(self, h.value, n, curve)

    def fixed_base(self, scalars=None, seed=0, n=None, curve=_lib.CURVE_ED25519):
        """``g_i = r_i * B`` on the device: explicit scalars (ints or packed bytes) or ``synth(seed, i)``."""
        if scalars is not None:
            raw = scalars if isinstance(scalars, (bytes, bytearray)) or hasattr(scalars, "nbytes") else \
                pack_scalars(scalars, ORDERS[curve])
            nbytes = raw.nbytes if hasattr(raw, "nbytes") else len(raw)
            n = nbytes // 32
            p, keep = _buf(raw)
        else:
            p, keep = ctypes.c_void_p(0), None
        h = ctypes.c_uint64()
        check(self.lib.vmsm_points_fixed_base(self.h, curve, p, seed, n, ctypes.byref(h)))
        return DevicePoints(self, h.value, n, curve)

    # -- scalars
    def upload_scalars(self, scalars, order=ED_L):
        raw = scalars if isinstance(scalars, (bytes, bytearray)) or hasattr(scalars, "nbytes") else pack_scalars(scalars, order)
        nbytes = raw.nbytes if hasattr(raw, "nbytes") else len(raw)
        n = nbytes // 32
        p, keep = _buf(raw)
        h = ctypes.c_uint64()
        check(self.lib.vmsm_scalars_upload(self.h, p, n, ctypes.byref(h)))
        return DeviceScalars(self, h.value, n)

    def synth_scalars(self, seed, n, curve=_lib.CURVE_ED25519):
        h = ctypes.c_uint64()
        check(self.lib.vmsm_scalars_synth(self.h, curve, seed, n, ctypes.byref(h)))
        return DeviceScalars(self, h.value, n)

    # -- MSM
    def msm(self, points, scalars, off=0, n=None):
        """End to end: host scalars (ints / packed bytes / numpy uint8) -> canonical affine (x, y)."""
        raw = scalars if isinstance(scalars, (bytes, bytearray)) or hasattr(scalars, "nbytes") else \
            pack_scalars(scalars, ORDERS[points.curve])
        nbytes = raw.nbytes if hasattr(raw, "nbytes") else len(raw)
        if n is None:
            n = nbytes // 32
        p, keep = _buf(raw)
        out = ctypes.create_string_buffer(WIRE_BYTES[points.curve])
        check(self.lib.vmsm_msm(self.h, points.handle, off, n, p, out))
        return unpack_any(out.raw, points.curve)[0]

    def msm_ext(self, points, off, n, extra, extra_off, n_extra, scalars):
        """Pedersen form in one pass: sum_{i<n} s_i P[off+i] + sum_{j<n_extra} s_{n+j} E[extra_off+j]."""
        raw = scalars if isinstance(scalars, (bytes, bytearray)) or hasattr(scalars, "nbytes") else \
            pack_scalars(scalars, ORDERS[points.curve])
        nbytes = raw.nbytes if hasattr(raw, "nbytes") else len(raw)
        if nbytes != 32 * (n + n_extra):
            raise ValueError("need n + n_extra scalars")
        p, keep = _buf(raw)
        out = ctypes.create_string_buffer(WIRE_BYTES[points.curve])
        check(self.lib.vmsm_msm_ext(self.h, points.handle, off, n, extra.handle, extra_off, n_extra, p, out))
        return unpack_any(out.raw, points.curve)[0]

    def concat(self, a, a_off, a_n, b=None, b_off=0, b_n=0):
        """Device-side copy: a[a_off:a_off+a_n] || b[b_off:b_off+b_n] as a new DevicePoints."""
        h = ctypes.c_uint64()
        check(self.lib.vmsm_points_concat(self.h, a.handle, a_off, a_n, b.handle if b is not None else 0, b_off, b_n,
                                          ctypes.byref(h)))
        return DevicePoints(self, h.value, a_n + b_n, a.curve)

    def msm_raw(self, points, ptr, off, n, out):
        """Zero-marshalling variant for benchmarks: ``ptr`` is a c_void_p to n*32 bytes, ``out`` a 64-byte buffer."""
        check(self.lib.vmsm_msm(self.h, points.handle, off, n, ptr, out))

    def msm_async(self, points, ptr, off, n, slot):
        """Asynchronous end-to-end MSM: ``ptr`` (c_void_p, ideally into ``pinned()`` memory) holds n*32 bytes of
        scalars and must stay valid until ``result(slot)`` returns; H2D overlaps the previous MSM."""
        check(self.lib.vmsm_msm_async(self.h, points.handle, off, n, ptr, slot))

    def msm_dev(self, points, scalars, slot=0, poff=0, soff=0, n=None):
        """Device-resident, asynchronous; fetch with ``result(slot)``."""
        if n is None:
            n = min(points.n - poff, scalars.n - soff)
        check(self.lib.vmsm_msm_dev(self.h, points.handle, poff, n, scalars.handle, soff, slot))

    def msm_dev_ext(self, points, poff, n, scalars, soff, extra, extra_off, extra_scalars, slot=0):
        """Asynchronous ``sum_{i<n} s[soff+i] P[poff+i] + sum_j e_j E[extra_off+j]`` with s resident on the device and
        the few extra scalars e_j (ints) from the host; fetch with ``result(slot)``."""
        raw = pack_scalars(extra_scalars, ORDERS[points.curve])
        check(self.lib.vmsm_msm_dev_ext(self.h, points.handle, poff, n, scalars.handle, soff, extra.handle, extra_off,
                                        len(extra_scalars), raw, slot))

    def msm_dev_ext_dot(self, points, poff, n, scalars, soff, extra, extra_off, dot_a, dot_aoff, dot_b, dot_boff, dot_n,
                        slot=0):
        """As ``msm_dev_ext`` with ONE extra term whose scalar is ``<dot_a[dot_aoff:], dot_b[dot_boff:]>`` (dot_n
        residues), computed on the device and never fetched: a folding round's cross term (compressed_pivot.py:41-42)."""
        check(self.lib.vmsm_msm_dev_ext_dot(self.h, points.handle, poff, n, scalars.handle, soff, extra.handle, extra_off,
                                            dot_a.handle, dot_aoff, dot_b.handle, dot_boff, dot_n, slot))

    def scalars_dot(self, a, aoff, b, boff, n):
        """sum_i a[aoff+i] * b[boff+i] modulo the Ed25519 group order, computed on the device."""
        out = ctypes.create_string_buffer(32)
        check(self.lib.vmsm_scalars_dot(self.h, a.handle, aoff, b.handle, boff, n, out))
        return int.from_bytes(out.raw, "little")

    # -- multi-GPU shards (one context per GPU; rank 0 owns the mailbox)
    def mailbox_create(self, world):
        """Owner side; returns the 64-byte CUDA-IPC handle other processes pass to ``mailbox_open_ipc``."""
        h = ctypes.create_string_buffer(64)
        check(self.lib.vmsm_mailbox_create(self.h, world, h))
        return h.raw

    def mailbox_open_ipc(self, handle, rank, world):
        check(self.lib.vmsm_mailbox_open_ipc(self.h, handle, rank, world))

    def mailbox_open_local(self, owner, rank):
        check(self.lib.vmsm_mailbox_open_local(self.h, owner.h, rank))

    def msm_dev_shard(self, points, scalars, slot, seq, poff=0, soff=0, n=None):
        """This GPU's slice of a sharded MSM; the owner's ``result(slot)`` is the sum over all ranks."""
        if n is None:
            n = min(points.n - poff, scalars.n - soff)
        check(self.lib.vmsm_msm_dev_shard(self.h, points.handle, poff, n, scalars.handle, soff, slot, seq))

    def result(self, slot=0, curve=_lib.CURVE_ED25519):
        out = ctypes.create_string_buffer(128)
        check(self.lib.vmsm_result_affine(self.h, slot, out))
        return unpack_any(out.raw[:WIRE_BYTES[curve]], curve)[0]

    def result_extended(self, slot=0):
        out = ctypes.create_string_buffer(128)
        check(self.lib.vmsm_result_extended(self.h, slot, out))
        return tuple(int.from_bytes(out.raw[32 * i:32 * i + 32], "little") for i in range(4))

    def lincomb(self, pts, scalars, curve=_lib.CURVE_ED25519):
        """sum_i s_i * P_i for a handful (<= 64) of host points; returns canonical affine (x, y)."""
        n = len(pts)
        out = ctypes.create_string_buffer(WIRE_BYTES[curve])
        check(self.lib.vmsm_lincomb(self.h, curve, pack_any(pts, curve), pack_scalars(scalars, ORDERS[curve]), n, out))
        return unpack_any(out.raw, curve)[0]

    def lincomb_async(self, pts, scalars, slot, curve=_lib.CURVE_ED25519):
        """As ``lincomb`` but asynchronous (Ed25519): fetch with ``result(slot)``."""
        check(self.lib.vmsm_lincomb_async(self.h, curve, pack_any(pts, curve), pack_scalars(scalars, ORDERS[curve]),
                                          len(pts), slot))

    def selftest_fe(self, op, a, b):
        n = len(a)
        A = b"".join(int(x).to_bytes(32, "little") for x in a)
        Bb = b"".join(int(x).to_bytes(32, "little") for x in b)
        out = ctypes.create_string_buffer(32 * n)
        check(self.lib.vmsm_selftest_fe(self.h, op, A, Bb, n, out))
        return [int.from_bytes(out.raw[32 * i:32 * i + 32], "little") for i in range(n)]

    def imad_peak(self):
        """Measured IMAD.WIDE.U32 peak of this device in tera limb-products per second."""
        v = ctypes.c_double()
        check(self.lib.vmsm_microbench_imad(self.h, ctypes.byref(v)))
        return v.value

    def close(self):
        if self.h:
            self.lib.vmsm_ctx_destroy(self.h)
            self.h = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default = None


def default_context():
    """Process-wide context on ``VMSM_DEVICE`` / ``LOCAL_RANK`` / device 0.  Raises VmsmError without a GPU."""
    global _default
    if _default is None or not _default.h:
        _default = Context()
    return _default
