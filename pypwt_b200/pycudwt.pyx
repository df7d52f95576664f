# cython: language_level=3
"""pycudwt -- drop-in `Wavelets` class over the B200-native C ABI (include/pwt_b200.h).

Mirrors the reference's Cython wrapper (src/pypwt.pyx:64-616): same constructor signature,
attributes, coefficient list layout, shape / dtype coercions, state rules and exceptions.
All numerics run in libpwt_b200.so on the GPU; there is no CPU fallback.

Extensions (not in the reference): 3D input with ndim=2 is treated as a stack of independent
images (the reference raises NotImplementedError, pypwt.pyx:155-156); `norms()`, `sync()`,
`image_into()`, pinned host buffers and the measurement helpers used by bench.py.
"""
import os
import numpy as np
from cython cimport view
from libc.stdint cimport intptr_t

cdef extern from "pwt_b200.h":
    ctypedef struct pwt_plan:
        pass
    ctypedef struct pwt_info:
        int batch
        int Nr
        int Nc
        int ndims
        int nlevels
        int hlen
        int do_swt
        int do_separable
        int do_cycle_spinning
        int state
        int shift_r
        int shift_c
        int nbands
        int device
    int pwt_create_batch(pwt_plan** out, const float* img, int batch, int Nr, int Nc, const char* wname,
                         int levels, int memisonhost, int do_separable, int do_cycle_spinning,
                         int do_swt, int ndim) nogil
    int pwt_clone(pwt_plan** out, const pwt_plan* src) nogil
    void pwt_destroy(pwt_plan* p) nogil
    int pwt_get_info(const pwt_plan* p, pwt_info* info) nogil
    int pwt_band_shape(const pwt_plan* p, int num, int* nr, int* nc) nogil
    int pwt_forward(pwt_plan* p) nogil
    int pwt_inverse(pwt_plan* p) nogil
    int pwt_soft_threshold(pwt_plan* p, float beta, int app, int normalize) nogil
    int pwt_hard_threshold(pwt_plan* p, float beta, int app, int normalize) nogil
    int pwt_group_soft_threshold(pwt_plan* p, float beta, int app, int normalize) nogil
    int pwt_shrink(pwt_plan* p, float beta, int app) nogil
    int pwt_proj_linf(pwt_plan* p, float beta, int app) nogil
    int pwt_circshift(pwt_plan* p, int sr, int sc, int inplace) nogil
    int pwt_norm1(pwt_plan* p, float* out) nogil
    int pwt_norm2sq(pwt_plan* p, float* out) nogil
    int pwt_norms(pwt_plan* p, double* n1, double* n2) nogil
    int pwt_add_wavelet(pwt_plan* dst, const pwt_plan* src, float alpha) nogil
    int pwt_get_image(pwt_plan* p, float* dst) nogil
    int pwt_set_image(pwt_plan* p, const float* img, int on_device) nogil
    int pwt_get_coeff(pwt_plan* p, float* dst, int num) nogil
    int pwt_set_coeff(pwt_plan* p, const float* src, int num, int on_device) nogil
    intptr_t pwt_image_ptr(pwt_plan* p) nogil
    intptr_t pwt_coeff_ptr(pwt_plan* p, int num) nogil
    long long pwt_coeffs_slab_floats(const pwt_plan* p) nogil
    long long pwt_coeff_offset(const pwt_plan* p, int num) nogil
    int pwt_get_coeffs(pwt_plan* p, float* dst) nogil
    intptr_t pwt_stream_ptr(pwt_plan* p) nogil
    int pwt_wait_stream(pwt_plan* p, intptr_t producer_stream) nogil
    int pwt_set_filters_forward(pwt_plan* p, const char* name, unsigned int len, const float* f1,
                                const float* f2, const float* f3, const float* f4) nogil
    int pwt_set_filters_inverse(pwt_plan* p, const float* f1, const float* f2, const float* f3,
                                const float* f4) nogil
    int pwt_print_informations(pwt_plan* p) nogil
    int pwt_sync(pwt_plan* p) nogil
    const char* pwt_last_error() nogil
    const char* pwt_version() nogil
    int pwt_device_count() nogil
    int pwt_lookup_filters(const char* wname, float* L, float* H, float* IL, float* IH) nogil
    int pwt_host_alloc(void** ptr, size_t nbytes) nogil
    int pwt_host_free(void* ptr) nogil
    int pwt_timer_start(pwt_plan* p) nogil
    int pwt_timer_stop(pwt_plan* p, float* ms) nogil
    int pwt_flush_l2(pwt_plan* p) nogil
    long long pwt_launch_count(const pwt_plan* p) nogil
    int pwt_set_kernel_mode(pwt_plan* p, int mode) nogil
    int pwt_set_device(int device) nogil
    int pwt_profile_enable(pwt_plan* p, int on) nogil
    int pwt_profile_read(pwt_plan* p, float* ms, int* tags, int cap) nogil
    int pwt_comm_unique_id(unsigned char* id) nogil
    int pwt_comm_init(pwt_plan* p, int nranks, int rank, const unsigned char* id) nogil
    int pwt_comm_destroy(pwt_plan* p) nogil
    int pwt_norms_allreduce(pwt_plan* p, double* n1, double* n2) nogil
    int pwt_comm_init_all(pwt_plan** plans, int n) nogil
    int pwt_norms_allreduce_group(pwt_plan** plans, int n, double* n1, double* n2) nogil

    # double precision (the reference's DOUBLEPRECISION build, filters.h:16-30)
    ctypedef struct pwt64_plan:
        pass
    int pwt64_create(pwt64_plan** out, const double* img, int batch, int Nr, int Nc, const char* wname, int levels,
                     int memisonhost, int do_separable, int do_cycle_spinning, int do_swt, int ndim) nogil
    void pwt64_destroy(pwt64_plan* p) nogil
    int pwt64_get_info(const pwt64_plan* p, pwt_info* info) nogil
    int pwt64_band_shape(const pwt64_plan* p, int num, int* nr, int* nc) nogil
    int pwt64_forward(pwt64_plan* p) nogil
    int pwt64_inverse(pwt64_plan* p) nogil
    int pwt64_soft_threshold(pwt64_plan* p, double beta, int app, int normalize) nogil
    int pwt64_hard_threshold(pwt64_plan* p, double beta, int app, int normalize) nogil
    int pwt64_shrink(pwt64_plan* p, double beta, int app) nogil
    int pwt64_norms(pwt64_plan* p, double* n1, double* n2) nogil
    int pwt64_get_image(pwt64_plan* p, double* dst) nogil
    int pwt64_set_image(pwt64_plan* p, const double* img, int on_device) nogil
    int pwt64_get_coeff(pwt64_plan* p, double* dst, int num) nogil
    int pwt64_set_coeff(pwt64_plan* p, const double* src, int num, int on_device) nogil
    int pwt64_sync(pwt64_plan* p) nogil
    int pwt64_timer_start(pwt64_plan* p) nogil
    int pwt64_timer_stop(pwt64_plan* p, float* ms) nogil
    long long pwt64_launch_count(const pwt64_plan* p) nogil
    int pwt64_lookup_filters(const char* wname, double* L, double* H, double* IL, double* IH) nogil
    int pwt64_set_filters_forward(pwt64_plan* p, const char* name, unsigned len, const double* lo, const double* hi) nogil
    int pwt64_set_filters_inverse(pwt64_plan* p, const double* lo, const double* hi) nogil

    # volumetric transform
    ctypedef struct pwt3_plan:
        pass
    int pwt3_create(pwt3_plan** out, const float* vol, int Nz, int Ny, int Nx, const char* wname, int levels, int memisonhost) nogil
    void pwt3_destroy(pwt3_plan* p) nogil
    int pwt3_levels(const pwt3_plan* p) nogil
    int pwt3_band_shape(const pwt3_plan* p, int level, int* nz, int* ny, int* nx) nogil
    int pwt3_forward(pwt3_plan* p) nogil
    int pwt3_inverse(pwt3_plan* p) nogil
    int pwt3_soft_threshold(pwt3_plan* p, float beta, int app) nogil
    int pwt3_hard_threshold(pwt3_plan* p, float beta, int app) nogil
    int pwt3_norms(pwt3_plan* p, double* n1, double* n2) nogil
    int pwt3_get_image(pwt3_plan* p, float* dst) nogil
    int pwt3_set_image(pwt3_plan* p, const float* vol, int on_device) nogil
    int pwt3_get_coeff(pwt3_plan* p, float* dst, int level, int b) nogil
    int pwt3_set_coeff(pwt3_plan* p, const float* src, int level, int b, int on_device) nogil
    int pwt3_sync(pwt3_plan* p) nogil
    int pwt3_timer_start(pwt3_plan* p) nogil
    int pwt3_timer_stop(pwt3_plan* p, float* ms) nogil
    long long pwt3_launch_count(const pwt3_plan* p) nogil

PWT_ERR_UNKNOWN_WAVELET = -2
PWT_ERR_UNSUPPORTED = -6
PWT_ERR_TOO_SMALL = -7


cdef str _errmsg():
    return pwt_last_error().decode("utf-8", "replace")


cdef void _free_pinned(void* ptr) noexcept:
    pwt_host_free(ptr)


def pinned_empty(shape, dtype=np.float32):
    """numpy array backed by page-locked host memory (fast, truly asynchronous H<->D copies).
    Falls back to a regular array if pinning fails or PYCUDWT_PINNED=0."""
    shape = tuple(int(s) for s in (shape if hasattr(shape, "__len__") else (shape,)))
    cdef size_t n = 1
    for s in shape:
        n *= <size_t> s
    dt = np.dtype(dtype)
    if n == 0 or os.environ.get("PYCUDWT_PINNED", "1") == "0":
        return np.empty(shape, dtype=dt)
    cdef void* ptr = NULL
    if pwt_host_alloc(&ptr, n * dt.itemsize) != 0 or ptr == NULL:
        return np.empty(shape, dtype=dt)
    cdef view.array arr = view.array(shape=(n * dt.itemsize,), itemsize=1, format="B", mode="c",
                                     allocate_buffer=False)
    arr.data = <char*> ptr
    arr.callback_free_data = _free_pinned
    return np.asarray(arr).view(dt).reshape(shape)


def pinned_zeros(shape, dtype=np.float32):
    a = pinned_empty(shape, dtype)
    a[...] = 0
    return a


def device_count():
    """Number of usable CUDA devices (0 on a CPU-only box)."""
    return pwt_device_count()


def set_device(int device):
    """Select the CUDA device used by Wavelets objects created afterwards (one process per GPU)."""
    if pwt_set_device(device) != 0:
        raise RuntimeError(_errmsg())


def lookup_filters(str wname):
    """(dec_lo, dec_hi, rec_lo, rec_hi) of a built-in bank, float32 -- the table of filters.cpp."""
    cdef float L[40]
    cdef float H[40]
    cdef float IL[40]
    cdef float IH[40]
    b = wname.encode("ASCII")
    cdef int hlen = pwt_lookup_filters(b, L, H, IL, IH)
    if hlen < 0:
        raise ValueError("unknown wavelet name %r" % wname)
    return tuple(np.array([arr[i] for i in range(hlen)], dtype=np.float32)
                 for arr in (<float[:40]> L, <float[:40]> H, <float[:40]> IL, <float[:40]> IH))


def comm_unique_id():
    """128-byte NCCL unique id (create on rank 0, broadcast to the other ranks)."""
    cdef unsigned char buf[128]
    if pwt_comm_unique_id(buf) != 0:
        raise RuntimeError(_errmsg())
    return bytes(buf[:128])


def comm_init_all(list plans):
    """Single-process multi-GPU: one NCCL communicator over `plans` (Wavelets objects on distinct devices)."""
    cdef pwt_plan* ps[64]
    cdef int n = len(plans), rc
    if n < 1 or n > 64:
        raise ValueError("expected 1..64 plans")
    for i in range(n):
        ps[i] = (<Wavelets?> plans[i]).w
    with nogil:
        rc = pwt_comm_init_all(ps, n)
    if rc != 0:
        raise RuntimeError(_errmsg())


def norms_allreduce_group(list plans):
    """Global (norm1, norm2sq) over the plans of a `comm_init_all` communicator (one NCCL group, one host thread)."""
    cdef pwt_plan* ps[64]
    cdef int n = len(plans), rc
    cdef double a = 0, b = 0
    if n < 1 or n > 64:
        raise ValueError("expected 1..64 plans")
    for i in range(n):
        ps[i] = (<Wavelets?> plans[i]).w
    with nogil:
        rc = pwt_norms_allreduce_group(ps, n, &a, &b)
    if rc != 0:
        raise RuntimeError(_errmsg())
    return a, b


# ---- device-array interop (SURVEY 8f rank 3; the reference only leaks raw addresses, wt.cu:658-665) ------------
cdef extern from "Python.h":
    object PyCapsule_New(void* pointer, const char* name, void (*destructor)(object) noexcept)
    void* PyCapsule_GetPointer(object capsule, const char* name) except? NULL
    int PyCapsule_IsValid(object capsule, const char* name)
    void Py_INCREF(object o)
    void Py_DECREF(object o)

from libc.stdint cimport int64_t, uint64_t, int32_t, uint8_t, uint16_t
from libc.stdlib cimport malloc, free

cdef struct DLDevice:
    int32_t device_type
    int32_t device_id

cdef struct DLDataType:
    uint8_t code
    uint8_t bits
    uint16_t lanes

cdef struct DLTensor:
    void* data
    DLDevice device
    int32_t ndim
    DLDataType dtype
    int64_t* shape
    int64_t* strides
    uint64_t byte_offset

cdef struct DLManagedTensor:
    DLTensor dl_tensor
    void* manager_ctx
    void (*deleter)(DLManagedTensor*) noexcept


cdef void _dl_deleter(DLManagedTensor* t) noexcept with gil:
    if t == NULL:
        return
    if t.manager_ctx != NULL:
        Py_DECREF(<object> t.manager_ctx)
    free(t.dl_tensor.shape)
    free(t)


cdef void _dl_capsule_destructor(object cap) noexcept:
    cdef DLManagedTensor* t
    if PyCapsule_IsValid(cap, "dltensor"):       # never consumed: we still own the tensor
        t = <DLManagedTensor*> PyCapsule_GetPointer(cap, "dltensor")
        _dl_deleter(t)


cdef class DeviceArray:
    """Zero-copy view of device memory owned by a `Wavelets` instance (its image or one coefficient band).

    Exposes `__cuda_array_interface__` (version 3: numba, cupy, torch.as_tensor) and the DLPack protocol
    (`torch.from_dlpack`, `cupy.from_dlpack`).  The view keeps its owner alive.  Work queued by the owner is
    ordered on the owner's stream, which both protocols report, so consumers synchronise correctly; writes through
    the view are seen by the next transform (the fused norm cache is dropped when a band view is handed out)."""
    cdef readonly object owner
    cdef readonly size_t ptr
    cdef readonly tuple shape
    cdef readonly int device
    cdef readonly size_t stream

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": "<f4", "data": (self.ptr, False), "version": 3, "strides": None,
                "stream": self.stream if self.stream else 1}

    @property
    def nbytes(self):
        n = 4
        for d in self.shape:
            n *= d
        return n

    def __dlpack_device__(self):
        return (2, self.device)                  # kDLCUDA

    def __dlpack__(self, stream=None, **kwargs):
        # the producer's work is on `self.stream`; make the consumer's stream wait for it (-1: no synchronisation wanted)
        if stream is not None and stream != -1:
            (<Wavelets> self.owner)._sync_for_consumer()
        cdef DLManagedTensor* t = <DLManagedTensor*> malloc(sizeof(DLManagedTensor))
        if t == NULL:
            raise MemoryError()
        cdef int nd = len(self.shape)
        t.dl_tensor.data = <void*> self.ptr
        t.dl_tensor.device.device_type = 2
        t.dl_tensor.device.device_id = self.device
        t.dl_tensor.ndim = nd
        t.dl_tensor.dtype.code = 2               # kDLFloat
        t.dl_tensor.dtype.bits = 32
        t.dl_tensor.dtype.lanes = 1
        t.dl_tensor.shape = <int64_t*> malloc(sizeof(int64_t) * (nd if nd > 0 else 1))
        for i in range(nd):
            t.dl_tensor.shape[i] = self.shape[i]
        t.dl_tensor.strides = NULL
        t.dl_tensor.byte_offset = 0
        Py_INCREF(self)
        t.manager_ctx = <void*> self
        t.deleter = _dl_deleter
        return PyCapsule_New(t, "dltensor", _dl_capsule_destructor)


cdef object _as_device_input(obj):
    """(ptr, shape, producer stream) of a float32 C-contiguous device array exposing __cuda_array_interface__
    (torch / cupy / numba / DeviceArray), or None for host data."""
    if isinstance(obj, np.ndarray):
        return None
    cai = getattr(obj, "__cuda_array_interface__", None)
    if cai is None:
        return None
    if np.dtype(cai["typestr"]) != np.float32:
        raise ValueError("device arrays must be float32 (got %s)" % cai["typestr"])
    shape = tuple(int(x) for x in cai["shape"])
    strides = cai.get("strides")
    if strides is not None:
        expect, acc = [], 4
        for d in reversed(shape):
            expect.append(acc)
            acc *= d
        if tuple(strides) != tuple(reversed(expect)):
            raise ValueError("device arrays must be C-contiguous")
    stream = cai.get("stream")
    return (int(cai["data"][0]), shape, int(stream) if stream else 0)


cdef class Wavelets:
    """
    Initializes the Wavelet transform from an image and given parameters.

    img: numpy.ndarray (float32 after coercion)
        2D image, 1D signal, or (extension) 3D stack of images
    wname: string
        Name of the wavelet
    levels: int
        Number of decomposition levels
    do_separable: int
        if not 0, perform a separable transform
    do_cycle_spinning: int
        if not 0, perform a random shift on the image (useful for iterative algorithms)
    do_swt: int
        if not 0, perform a Stationary (non-decimated) wavelet transform
    ndim: int
        1 on 2D input = batched 1D transform of the rows
    """
    cdef pwt_plan* w
    cdef readonly int Nr
    cdef readonly int Nc
    cdef readonly list sizes
    cdef readonly str wname
    cdef readonly int levels
    cdef readonly int do_cycle_spinning
    cdef readonly int hlen
    cdef readonly int do_swt
    cdef readonly int do_separable
    cdef readonly int ndim
    cdef readonly int batched1d
    cdef readonly int batch
    cdef list _coeffs
    cdef object _slab          # pinned host slab that every array of _coeffs is a view of (one D2H fills them all)
    cdef tuple shape
    cdef int _is1d

    def __cinit__(self, img, str wname, int levels, int do_separable=1, int do_cycle_spinning=0,
                  int do_swt=0, int ndim=2, Wavelets copy=None):
        self.w = NULL
        if copy is not None:            # shell for Wavelets.copy(): the cloned plan is attached by the caller
            self.Nr, self.Nc, self.sizes, self.wname, self.levels = copy.Nr, copy.Nc, list(copy.sizes), copy.wname, copy.levels
            self.do_cycle_spinning, self.hlen, self.do_swt, self.do_separable = copy.do_cycle_spinning, copy.hlen, copy.do_swt, copy.do_separable
            self.ndim, self.batched1d, self.batch, self.shape, self._is1d = copy.ndim, copy.batched1d, copy.batch, copy.shape, copy._is1d
            return
        # a device array (anything exposing __cuda_array_interface__: torch, cupy, numba, DeviceArray) is taken
        # where it lies: the reference's `memisonhost=0` constructor argument (wt.cu:84,145-150), which its wrapper
        # never passes (pypwt.pyx:169)
        dev_in = _as_device_input(img)
        if dev_in is None:
            img = self._checkarray(img)
            ishape = tuple(int(x) for x in img.shape)
        else:
            ishape = dev_in[1]
        indim = len(ishape)
        ndim = min(ndim, 2)                                       # pypwt.pyx:145
        self.batched1d = 0
        self.batch = 1
        if indim == 2:
            self.Nr = ishape[0]
            self.Nc = ishape[1]
            if indim != ndim:
                self.batched1d = 1
        elif indim == 1:
            self.Nr = 1
            self.Nc = ishape[0]
        elif indim == 3 and ndim == 2:
            # extension: stack of independent 2D images (SURVEY 8e)
            self.batch = ishape[0]
            self.Nr = ishape[1]
            self.Nc = ishape[2]
        else:
            raise NotImplementedError("Wavelets(): Only 1D and 2D transforms are supported for now")
        self.shape = ishape
        self.wname = wname
        py_wname = wname.encode("ASCII")
        self.do_cycle_spinning = do_cycle_spinning
        self.do_swt = do_swt
        self.ndim = indim

        if pwt_device_count() < 1:
            raise RuntimeError("pycudwt: no CUDA device available (there is no CPU fallback)")
        cdef const float* src = <const float*> <size_t> (img.ctypes.data if dev_in is None else dev_in[0])
        cdef const char* c_wname = py_wname
        cdef int rc, c_ndim = ndim, c_levels = levels, c_sep = do_separable, c_onhost = 1 if dev_in is None else 0
        with nogil:
            rc = pwt_create_batch(&self.w, src, self.batch, self.Nr, self.Nc, c_wname, c_levels, c_onhost,
                                  c_sep, self.do_cycle_spinning, self.do_swt, c_ndim)
        if rc != 0:
            self.w = NULL
            msg = _errmsg()
            if rc in (PWT_ERR_UNKNOWN_WAVELET, PWT_ERR_TOO_SMALL, PWT_ERR_UNSUPPORTED, -1):
                raise ValueError(msg)
            raise RuntimeError(msg)
        # Retrieve the possibly updated attributes (pypwt.pyx:181-183)
        cdef pwt_info info
        pwt_get_info(self.w, &info)
        self.levels = info.nlevels
        self.hlen = info.hlen
        self.do_separable = info.do_separable
        self._is1d = 1 if info.ndims == 1 else 0
        self.sizes = self._compute_sizes()

        self._alloc_host_coeffs()

    cdef _alloc_host_coeffs(self):
        # persistent host buffers: [A, [H1, V1, D1], ...] or [A, D1, ...]  (pypwt.pyx:191-205), all of them views of
        # ONE pinned slab laid out like the device's coefficient region, so `coeffs` is a single D2H copy
        lead = (self.batch,) if len(self.shape) == 3 else ()
        self._slab = pinned_zeros(int(pwt_coeffs_slab_floats(self.w)))

        def view(num, shp):
            shp = lead + tuple(shp)
            off = int(pwt_coeff_offset(self.w, num))
            return self._slab[off:off + int(np.prod(shp))].reshape(shp)

        self._coeffs = [view(0, self.sizes[-1])]
        for i in range(self.levels):
            if self._is1d:
                self._coeffs.append(view(i + 1, self.sizes[i]))
            else:
                self._coeffs.append([view(3 * i + 1 + j, self.sizes[i]) for j in range(3)])

    def info(self):
        """Print some information on the current ``Wavelets`` instance."""
        pwt_print_informations(self.w)

    def __repr__(self):
        self.info()
        return ""

    def __str__(self):
        self.info()
        return ""

    @staticmethod
    def _checkarray(arr, shp=None):
        arr = np.asarray(arr)
        res = arr
        if arr.dtype != np.float32 or not arr.flags["C_CONTIGUOUS"]:
            res = np.ascontiguousarray(arr, dtype=np.float32)
        if shp is not None:
            if arr.ndim != len(shp):
                raise ValueError("Invalid number of dimensions (expected %d, got %d)" % (len(shp), arr.ndim))
            for i in range(arr.ndim):
                if arr.shape[i] != shp[i]:
                    raise ValueError("The image does not have the correct shape (expected %s, got %s)"
                                     % (str(shp), str(arr.shape)))
        return res

    @staticmethod
    def div2(n):
        """Returns (N + (N%2))/2: image size at the next scale."""
        return (n + (n & 1)) // 2

    def _compute_sizes(self):
        cdef int nr = 0, nc = 0
        res = []
        for i in range(self.levels):
            num = (3 * i + 1) if not self._is1d else (i + 1)
            pwt_band_shape(self.w, num, &nr, &nc)
            res.append((nr, nc))
        return res

    cdef object _coeff_ref(self, int num):
        if num == 0:
            return self._coeffs[0]
        if not self._is1d:
            return self._coeffs[(num - 1) // 3 + 1][(num - 1) % 3]
        return self._coeffs[num]

    def coeff_only(self, int num):
        """
        Get only the coeff "num".  2D : [0: A, 1: H1, 2: V1, 3: D1, 4: H2, ...] ; 1D : [0: A, 1: D1, ...]
        The returned array is the persistent host buffer of that band, refreshed in place.
        """
        nb = (3 * self.levels + 1) if not self._is1d else (self.levels + 1)
        if num < 0 or num >= nb:
            raise IndexError("coefficient number %d out of range" % num)
        coeff_ref = self._coeff_ref(num)
        cdef float* dst = <float*> <size_t> coeff_ref.ctypes.data
        cdef int numc
        with nogil:
            numc = pwt_get_coeff(self.w, dst, num)
        if numc != min(coeff_ref.size, 0x7fffffff):
            raise RuntimeError("Wavelets.coeff_only(): something went wrong when retrieving coefficients numbef %d, expected %d coeffs, got %d"
                               % (num, coeff_ref.size, numc))
        return coeff_ref

    @property
    def coeffs(self):
        """[A, [H1, V1, D1], [H2, V2, D2], ...] (2D) or [A, D1, ...] (1D): every band, moved from the device with ONE
        copy into the persistent pinned host buffers (the reference issues 3L+1 blocking copies, pypwt.pyx:290-306)."""
        cdef float* dst = <float*> <size_t> self._slab.ctypes.data
        cdef int rc
        with nogil:
            rc = pwt_get_coeffs(self.w, dst)
        if rc != 0:
            raise RuntimeError("Wavelets.coeffs: something went wrong when retrieving the coefficients (%s)"
                               % ("inverse() has been performed" if rc == 1 else _errmsg()))
        return self._coeffs

    def _img_shape(self):
        return ((self.batch,) if len(self.shape) == 3 else ()) + (self.Nr, self.Nc)

    @property
    def image(self):
        res = np.empty(self._img_shape(), dtype=np.float32)
        return self.image_into(res)

    def image_into(self, out):
        """Copy the device image into `out` (float32, C-contiguous, e.g. from `pinned_empty`)."""
        if out.dtype != np.float32 or not out.flags["C_CONTIGUOUS"] or out.size != self.batch * self.Nr * self.Nc:
            raise ValueError("image_into(): expected a C-contiguous float32 array of %d elements" % (self.batch * self.Nr * self.Nc))
        cdef float* dst = <float*> <size_t> out.ctypes.data
        cdef int numc
        with nogil:
            numc = pwt_get_image(self.w, dst)
        if numc != min(out.size, 0x7fffffff):
            raise RuntimeError("Wavelets.image(): something went wrong when retrieving image, expected %d coeffs, got %d" % (out.size, numc))
        return out

    cdef _upload_image(self, img, shp):
        """host array -> H2D, device array (`__cuda_array_interface__`) -> D2D on the plan's stream, ordered after the
        producer's stream when the array names one (wt.cu:425-432 with mem_is_on_device = 0 / 1)."""
        cdef const float* src
        cdef int rc, on_device = 0
        cdef intptr_t producer = 0
        dev_in = _as_device_input(img)
        if dev_in is None:
            img = self._checkarray(img, shp)
            src = <const float*> <size_t> img.ctypes.data
        else:
            if tuple(dev_in[1]) != tuple(shp):
                raise ValueError("The image does not have the correct shape (expected %s, got %s)" % (str(shp), str(dev_in[1])))
            src = <const float*> <size_t> dev_in[0]
            on_device = 1
            producer = dev_in[2]
        with nogil:
            rc = 0
            if producer > 2:
                rc = pwt_wait_stream(self.w, producer)
            if rc == 0:
                rc = pwt_set_image(self.w, src, on_device)
        if rc != 0:
            raise RuntimeError(_errmsg())

    def set_image(self, img):
        """Replace the device image (does not update the coefficients).  `img`: numpy array or device array."""
        self._upload_image(img, self._img_shape())

    def forward(self, img=None):
        """Forward transform of `img` (if given, checked against the original shape) or of the current image."""
        cdef int rc
        if img is not None:
            self._upload_image(img, self.shape)
        with nogil:
            rc = pwt_forward(self.w)
        if rc < 0:
            raise RuntimeError(_errmsg())

    def inverse(self):
        """Inverse transform; consumes the coefficients (a second call is refused until forward())."""
        cdef int rc
        with nogil:
            rc = pwt_inverse(self.w)
        if rc < 0:
            raise RuntimeError(_errmsg())

    def soft_threshold(self, float beta, int do_threshold_appcoeffs=0, int normalize=0):
        """ST(x, t) = (|x| - t)_+ . sign(x) on the detail (and optionally approximation) coefficients."""
        if pwt_soft_threshold(self.w, beta, do_threshold_appcoeffs, normalize) < 0:
            raise RuntimeError(_errmsg())

    def hard_threshold(self, float beta, int do_threshold_appcoeffs=0, int normalize=0):
        """HT(x, t) = x . 1_{|x| > t}."""
        if pwt_hard_threshold(self.w, beta, do_threshold_appcoeffs, normalize) < 0:
            raise RuntimeError(_errmsg())

    def group_soft_threshold(self, float beta, int do_threshold_appcoeffs=0, int normalize=0):
        """Joint (h, v, d[, a]) shrinkage (C++ only in the reference, wt.h:58)."""
        if pwt_group_soft_threshold(self.w, beta, do_threshold_appcoeffs, normalize) < 0:
            raise RuntimeError(_errmsg())

    def shrink(self, float beta, int do_threshold_appcoeffs=1):
        """shrink(x, t) = x / (1 + t)."""
        if pwt_shrink(self.w, beta, do_threshold_appcoeffs) < 0:
            raise RuntimeError(_errmsg())

    def proj_linf(self, float beta, int do_threshold_appcoeffs=1):
        """Projection onto the L-infinity ball of radius beta (C++ only in the reference, wt.h:60)."""
        if pwt_proj_linf(self.w, beta, do_threshold_appcoeffs) < 0:
            raise RuntimeError(_errmsg())

    def circshift(self, int sr, int sc):
        """In-place circular shift of the device image (np.roll(img, (sr, sc)))."""
        if pwt_circshift(self.w, sr, sc, 1) < 0:
            raise RuntimeError(_errmsg())

    def norm1(self):
        """L1 norm of all coefficients."""
        cdef float r = 0
        if pwt_norm1(self.w, &r) != 0:
            raise RuntimeError(_errmsg())
        return r

    def norm2sq(self):
        """Squared L2 norm of all coefficients."""
        cdef float r = 0
        if pwt_norm2sq(self.w, &r) != 0:
            raise RuntimeError(_errmsg())
        return r

    def norms(self):
        """(norm1, norm2sq) in one pass, double precision (local to this GPU)."""
        cdef double a = 0, b = 0
        cdef int rc
        with nogil:
            rc = pwt_norms(self.w, &a, &b)
        if rc != 0:
            raise RuntimeError(_errmsg())
        return a, b

    def add_wavelet(self, Wavelets W, alpha=1.0):
        """coeffs += alpha * W.coeffs."""
        cdef float c_alpha = alpha
        return pwt_add_wavelet(self.w, W.w, c_alpha)

    def copy(self):
        """Deep copy (device state, custom filters included): the reference's C++ copy constructor (wt.cu:191-222)."""
        cdef pwt_plan* fresh = NULL
        cdef int rc
        with nogil:
            rc = pwt_clone(&fresh, self.w)
        if rc != 0:
            raise RuntimeError(_errmsg())
        # the Python shell is built around the cloned plan directly: no throw-away plan, and a custom bank's name
        # (unknown to the filter table) is not looked up again
        cdef Wavelets other = Wavelets.__new__(Wavelets, None, self.wname, self.levels, copy=self)
        other.w = fresh
        other._alloc_host_coeffs()
        return other

    def set_coeff(self, coeff, int num, check=False):
        """Overwrite coefficient `num` on the device.  `coeff`: numpy array or device array."""
        cdef const float* src
        cdef int rc, on_device = 0
        cdef intptr_t producer = 0
        ref = self._coeff_ref(num)
        dev_in = _as_device_input(coeff)
        if dev_in is None:
            coeff = self._checkarray(coeff)
            cshape, csize = coeff.shape, coeff.size
            src = <const float*> <size_t> coeff.ctypes.data
        else:
            cshape, csize = dev_in[1], int(np.prod(dev_in[1]))
            src = <const float*> <size_t> dev_in[0]
            on_device = 1
            producer = dev_in[2]
        if check:
            dcoeff = self.coeff_only(num)
            if dcoeff.shape != tuple(cshape):
                raise ValueError("set_coefInvalid coefficient shape : expected %s, got %s" % (str(dcoeff.shape), str(cshape)))
        if csize != ref.size:
            raise ValueError("set_coeff(): expected %d elements, got %d" % (ref.size, csize))
        with nogil:
            rc = 0
            if producer > 2:
                rc = pwt_wait_stream(self.w, producer)
            if rc == 0:
                rc = pwt_set_coeff(self.w, src, num, on_device)
        if rc != 0:
            raise RuntimeError(_errmsg())

    def set_wavelets_filters(self, filter_name, lowpass, highpass, i_lowpass, i_highpass,
                             LH=None, HL=None, i_LH=None, i_HL=None):
        """Custom filter bank (re-defines the transform).  Non-separable mode takes the four 2D
        filters LL=lowpass, LH, HL, HH=highpass (and their inverse counterparts)."""
        if any(len(arr) != len(lowpass) for arr in [lowpass, highpass, i_lowpass, i_highpass, LH, HL, i_LH, i_HL] if arr is not None):
            raise ValueError("All filters must have the same length")
        lp = self._checkarray(lowpass); hp = self._checkarray(highpass)
        ilp = self._checkarray(i_lowpass); ihp = self._checkarray(i_highpass)
        name = filter_name.encode("ASCII")
        cdef unsigned int flen = len(lowpass)
        cdef int rc
        if self.do_separable:
            rc = pwt_set_filters_forward(self.w, name, flen, <const float*> <size_t> lp.ctypes.data,
                                         <const float*> <size_t> hp.ctypes.data, NULL, NULL)
            if rc == 0:
                rc = pwt_set_filters_inverse(self.w, <const float*> <size_t> ilp.ctypes.data,
                                             <const float*> <size_t> ihp.ctypes.data, NULL, NULL)
        else:
            if LH is None or HL is None or i_LH is None or i_HL is None:
                raise ValueError("Expected LH and HL filters for non-separable transform")
            lh = self._checkarray(LH); hl = self._checkarray(HL)
            ilh = self._checkarray(i_LH); ihl = self._checkarray(i_HL)
            rc = pwt_set_filters_forward(self.w, name, flen, <const float*> <size_t> lp.ctypes.data,
                                         <const float*> <size_t> lh.ctypes.data,
                                         <const float*> <size_t> hl.ctypes.data,
                                         <const float*> <size_t> hp.ctypes.data)
            if rc == 0:
                rc = pwt_set_filters_inverse(self.w, <const float*> <size_t> ilp.ctypes.data,
                                             <const float*> <size_t> ilh.ctypes.data,
                                             <const float*> <size_t> ihl.ctypes.data,
                                             <const float*> <size_t> ihp.ctypes.data)
        if rc != 0:
            raise ValueError("set_wavelets_filters() failed with code %d" % rc)
        cdef pwt_info info
        pwt_get_info(self.w, &info)
        self.hlen = info.hlen
        self.wname = filter_name

    def image_int_ptr(self):
        """Address of the device image."""
        return pwt_image_ptr(self.w)

    def coeff_int_ptr(self, int num):
        """Address of a device coefficient band."""
        return pwt_coeff_ptr(self.w, num)

    cdef DeviceArray _device_view(self, size_t ptr, tuple shape):
        cdef pwt_info info
        pwt_get_info(self.w, &info)
        cdef DeviceArray d = DeviceArray.__new__(DeviceArray)
        d.owner = self
        d.ptr = ptr
        d.shape = shape
        d.device = info.device
        d.stream = <size_t> pwt_stream_ptr(self.w)
        return d

    def _sync_for_consumer(self):
        self.sync()

    @property
    def image_device(self):
        """Zero-copy device view of the image (`__cuda_array_interface__` + DLPack), e.g. `torch.as_tensor(W.image_device,
        device="cuda")` -- the end-to-end path without the PCIe round trip of `image`."""
        return self._device_view(<size_t> pwt_image_ptr(self.w), self._img_shape())

    def coeff_device(self, int num):
        """Zero-copy device view of coefficient band `num` (numbering of `coeff_only`)."""
        nb = (3 * self.levels + 1) if not self._is1d else (self.levels + 1)
        if num < 0 or num >= nb:
            raise IndexError("coefficient number %d out of range" % num)
        return self._device_view(<size_t> pwt_coeff_ptr(self.w, num), tuple(self._coeff_ref(num).shape))

    @property
    def coeffs_device(self):
        """Device views in the layout of `coeffs`: [A, [H1, V1, D1], ...] (2D) or [A, D1, ...] (1D)."""
        out = [self.coeff_device(0)]
        for i in range(self.levels):
            if self._is1d:
                out.append(self.coeff_device(i + 1))
            else:
                out.append([self.coeff_device(3 * i + 1 + j) for j in range(3)])
        return out

    # ---- extensions ------------------------------------------------------------------------
    def sync(self):
        """Block until all work queued by this instance has finished (forward/inverse are asynchronous)."""
        cdef int rc
        with nogil:
            rc = pwt_sync(self.w)
        if rc != 0:
            raise RuntimeError(_errmsg())

    @property
    def state(self):
        cdef pwt_info info
        pwt_get_info(self.w, &info)
        return info.state

    @property
    def current_shift(self):
        cdef pwt_info info
        pwt_get_info(self.w, &info)
        return (info.shift_r, info.shift_c)

    @property
    def launch_count(self):
        return pwt_launch_count(self.w)

    def set_kernel_mode(self, int mode):
        """0 = auto, 1 = force the generic tiled kernels (used by the parity tests)."""
        pwt_set_kernel_mode(self.w, mode)

    def profile_enable(self, int on=1):
        """Bracket every transform kernel with CUDA events (see pwt_profile_enable)."""
        if pwt_profile_enable(self.w, on) != 0:
            raise RuntimeError(_errmsg())

    def profile_read(self):
        """[(tag, ms), ...] in launch order; tag = 100*level + (1 forward | 2 inverse)."""
        cdef float ms[512]
        cdef int tags[512]
        cdef int n
        with nogil:
            n = pwt_profile_read(self.w, ms, tags, 512)
        if n < 0:
            raise RuntimeError(_errmsg())
        return [(tags[i], ms[i]) for i in range(n)]

    def timer_start(self):
        if pwt_timer_start(self.w) != 0:
            raise RuntimeError(_errmsg())

    def timer_stop(self):
        cdef float ms = 0
        cdef int rc
        with nogil:
            rc = pwt_timer_stop(self.w, &ms)
        if rc != 0:
            raise RuntimeError(_errmsg())
        return ms

    def flush_l2(self):
        if pwt_flush_l2(self.w) != 0:
            raise RuntimeError(_errmsg())

    def comm_init(self, int nranks, int rank, bytes unique_id):
        if len(unique_id) != 128:
            raise ValueError("unique_id must be 128 bytes")
        cdef const unsigned char* idp = unique_id
        cdef int rc
        with nogil:
            rc = pwt_comm_init(self.w, nranks, rank, idp)
        if rc != 0:
            raise RuntimeError(_errmsg())

    def comm_destroy(self):
        pwt_comm_destroy(self.w)

    def norms_allreduce(self):
        """Global (norm1, norm2sq) over all ranks of the communicator (fused reduction + NCCL all-reduce)."""
        cdef double a = 0, b = 0
        cdef int rc
        with nogil:
            rc = pwt_norms_allreduce(self.w, &a, &b)
        if rc != 0:
            raise RuntimeError(_errmsg())
        return a, b

    def __dealloc__(self):
        self.cleanup()

    def cleanup(self):  # should not be called manually
        if self.w is not NULL:
            pwt_destroy(self.w)
            self.w = NULL

    @classmethod
    def version(cls):
        """Version string of the library this is a drop-in for."""
        return pwt_version().decode("ASCII")


def lookup_filters64(str wname):
    """(L, H, IL, IH) of a built-in bank at the table's full (double) precision."""
    cdef double L[40]
    cdef double H[40]
    cdef double IL[40]
    cdef double IH[40]
    b = wname.encode("ASCII")
    cdef int hlen = pwt64_lookup_filters(b, L, H, IL, IH)
    if hlen < 0:
        raise ValueError("unknown wavelet name '%s'" % wname)
    return (np.array([L[k] for k in range(hlen)], dtype=np.float64), np.array([H[k] for k in range(hlen)], dtype=np.float64),
            np.array([IL[k] for k in range(hlen)], dtype=np.float64), np.array([IH[k] for k in range(hlen)], dtype=np.float64))


cdef class Wavelets64:
    """Double-precision counterpart of `Wavelets`: the reference's DOUBLEPRECISION build (libpdwtd.so,
    pdwt/src/filters.h:16-30), which its Python wrapper cannot reach.  Same constructor arguments and attribute /
    method names (pypwt.pyx:64-118); images and coefficients are float64.  Not carried over: non-separable custom
    banks, add_wavelet, device-array interop."""
    cdef pwt64_plan* w
    cdef readonly int Nr
    cdef readonly int Nc
    cdef readonly list sizes
    cdef readonly str wname
    cdef readonly int levels
    cdef readonly int do_cycle_spinning
    cdef readonly int hlen
    cdef readonly int do_swt
    cdef readonly int do_separable
    cdef readonly int ndim
    cdef readonly int batched1d
    cdef readonly int batch
    cdef tuple shape
    cdef int _is1d
    cdef int _nbands

    def __cinit__(self, img, str wname, int levels, int do_separable=1, int do_cycle_spinning=0, int do_swt=0, int ndim=2):
        self.w = NULL
        img = np.ascontiguousarray(img, dtype=np.float64)
        ishape = tuple(int(x) for x in img.shape)
        indim = len(ishape)
        ndim = min(ndim, 2)
        self.batched1d = 0
        self.batch = 1
        if indim == 2:
            self.Nr, self.Nc = ishape
            if indim != ndim:
                self.batched1d = 1
        elif indim == 1:
            self.Nr, self.Nc = 1, ishape[0]
        elif indim == 3 and ndim == 2:
            self.batch, self.Nr, self.Nc = ishape
        else:
            raise NotImplementedError("Wavelets64(): Only 1D and 2D transforms are supported for now")
        self.shape = ishape
        self.wname = wname
        self.do_cycle_spinning = do_cycle_spinning
        self.do_swt = do_swt
        self.ndim = indim
        if pwt_device_count() < 1:
            raise RuntimeError("pycudwt: no CUDA device available (there is no CPU fallback)")
        py_wname = wname.encode("ASCII")
        cdef const double* src = <const double*> <size_t> img.ctypes.data
        cdef const char* c_wname = py_wname
        cdef int rc, c_ndim = ndim, c_levels = levels, c_sep = do_separable
        with nogil:
            rc = pwt64_create(&self.w, src, self.batch, self.Nr, self.Nc, c_wname, c_levels, 1, c_sep,
                              self.do_cycle_spinning, self.do_swt, c_ndim)
        if rc != 0:
            self.w = NULL
            msg = _errmsg()
            if rc in (PWT_ERR_UNKNOWN_WAVELET, PWT_ERR_TOO_SMALL, PWT_ERR_UNSUPPORTED, -1):
                raise ValueError(msg)
            raise RuntimeError(msg)
        cdef pwt_info info
        pwt64_get_info(self.w, &info)
        self.levels = info.nlevels
        self.hlen = info.hlen
        self.do_separable = info.do_separable
        self._is1d = 1 if info.ndims == 1 else 0
        self._nbands = info.nbands
        cdef int nr = 0, nc = 0
        per = 1 if self._is1d else 3
        self.sizes = []
        for i in range(self.levels):
            pwt64_band_shape(self.w, per * i + 1, &nr, &nc)
            self.sizes.append((nr, nc))

    def _band(self, int num, shp):
        lead = (self.batch,) if len(self.shape) == 3 else ()
        out = np.empty(lead + tuple(shp), dtype=np.float64)
        cdef double* dst = <double*> <size_t> out.ctypes.data
        cdef int n
        with nogil:
            n = pwt64_get_coeff(self.w, dst, num)
        if n == 0:
            raise RuntimeError("Wavelets64.coeffs: the coefficients were consumed by inverse()")
        return out

    @property
    def coeffs(self):
        """[A, [H1, V1, D1], ...] (2D) or [A, D1, ...] (1D / batched 1D), float64, level 1 = finest (pypwt.pyx:261-306)."""
        res = [self._band(0, self.sizes[-1])]
        for i in range(self.levels):
            if self._is1d:
                res.append(self._band(i + 1, self.sizes[i]))
            else:
                res.append([self._band(3 * i + 1 + j, self.sizes[i]) for j in range(3)])
        return res

    @property
    def image(self):
        out = np.empty(self.shape, dtype=np.float64)
        cdef double* dst = <double*> <size_t> out.ctypes.data
        cdef int n
        with nogil:
            n = pwt64_get_image(self.w, dst)
        if n == 0:
            raise RuntimeError(_errmsg())
        return out

    def set_image(self, img):
        img = np.ascontiguousarray(img, dtype=np.float64)
        if tuple(img.shape) != self.shape:
            raise ValueError("Wavelets64.set_image(): shape mismatch %s != %s" % (img.shape, self.shape))
        cdef const double* src = <const double*> <size_t> img.ctypes.data
        cdef int rc
        with nogil:
            rc = pwt64_set_image(self.w, src, 0)
        if rc != 0:
            raise RuntimeError(_errmsg())

    def set_coeff(self, arr, int num):
        cdef int nr = 0, nc = 0
        if num < 0 or num >= self._nbands:
            raise ValueError("Wavelets64.set_coeff(): band %d out of range" % num)
        pwt64_band_shape(self.w, num, &nr, &nc)
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        if arr.size != self.batch * nr * nc:
            raise ValueError("Wavelets64.set_coeff(): band %d holds %d values, got %d" % (num, self.batch * nr * nc, arr.size))
        cdef const double* src = <const double*> <size_t> arr.ctypes.data
        cdef int rc
        with nogil:
            rc = pwt64_set_coeff(self.w, src, num, 0)
        if rc != 0:
            raise RuntimeError(_errmsg())

    def forward(self, img=None):
        if img is not None:
            self.set_image(img)
        cdef int rc
        with nogil:
            rc = pwt64_forward(self.w)
        if rc != 0:
            raise RuntimeError(_errmsg())

    def inverse(self):
        cdef int rc
        with nogil:
            rc = pwt64_inverse(self.w)
        if rc < 0:
            raise RuntimeError(_errmsg())

    def soft_threshold(self, double beta, int do_threshold_appcoeffs=0, int normalize=0):
        if pwt64_soft_threshold(self.w, beta, do_threshold_appcoeffs, normalize) < 0:
            raise RuntimeError(_errmsg())

    def hard_threshold(self, double beta, int do_threshold_appcoeffs=0, int normalize=0):
        if pwt64_hard_threshold(self.w, beta, do_threshold_appcoeffs, normalize) < 0:
            raise RuntimeError(_errmsg())

    def shrink(self, double beta, int do_threshold_appcoeffs=1):
        if pwt64_shrink(self.w, beta, do_threshold_appcoeffs) < 0:
            raise RuntimeError(_errmsg())

    def norms(self):
        cdef double a = 0, b = 0
        cdef int rc
        with nogil:
            rc = pwt64_norms(self.w, &a, &b)
        if rc != 0:
            raise RuntimeError(_errmsg())
        return a, b

    def norm1(self):
        return self.norms()[0]

    def norm2sq(self):
        return self.norms()[1]

    def set_wavelets_filters(self, filter_name, lowpass, highpass, i_lowpass, i_highpass):
        """Custom separable filter bank in double precision (pypwt.pyx:487-575 -> wt.cu:558-600 with DTYPE = double);
        re-defines the transform.  Non-separable plans are refused (the double-precision plans have no F x F stencils)."""
        if any(len(arr) != len(lowpass) for arr in (highpass, i_lowpass, i_highpass)):
            raise ValueError("All filters must have the same length")
        lp = np.ascontiguousarray(lowpass, dtype=np.float64); hp = np.ascontiguousarray(highpass, dtype=np.float64)
        ilp = np.ascontiguousarray(i_lowpass, dtype=np.float64); ihp = np.ascontiguousarray(i_highpass, dtype=np.float64)
        name = filter_name.encode("ASCII")
        cdef unsigned int flen = len(lowpass)
        cdef int rc = pwt64_set_filters_forward(self.w, name, flen, <const double*> <size_t> lp.ctypes.data,
                                                <const double*> <size_t> hp.ctypes.data)
        if rc == 0:
            rc = pwt64_set_filters_inverse(self.w, <const double*> <size_t> ilp.ctypes.data, <const double*> <size_t> ihp.ctypes.data)
        if rc != 0:
            raise ValueError("set_wavelets_filters() failed with code %d" % rc)
        cdef pwt_info info
        pwt64_get_info(self.w, &info)
        self.hlen = info.hlen
        self.wname = filter_name

    def sync(self):
        pwt64_sync(self.w)

    def timer_start(self):
        if pwt64_timer_start(self.w) != 0:
            raise RuntimeError(_errmsg())

    def timer_stop(self):
        cdef float ms = 0
        cdef int rc
        with nogil:
            rc = pwt64_timer_stop(self.w, &ms)
        if rc != 0:
            raise RuntimeError(_errmsg())
        return ms

    @property
    def launch_count(self):
        return int(pwt64_launch_count(self.w))

    @property
    def current_shift(self):
        cdef pwt_info info
        pwt64_get_info(self.w, &info)
        return (info.shift_r, info.shift_c)

    def __dealloc__(self):
        if self.w is not NULL:
            pwt64_destroy(self.w)
            self.w = NULL


VOLUME_BAND_KEYS = ("aad", "ada", "add", "daa", "dad", "dda", "ddd")   # band index 1..7 = 4 dz + 2 dy + dx, (z, y, x) key order


cdef class Wavelets3D:
    """Separable 3D DWT of a volume [Nz][Ny][Nx] (float32) -- the extension the reference names as missing ("3D is not
    handled", pdwt/README.md:29).  Same conventions as `Wavelets` (periodisation, ceil halving, level clipping).
    `coeffs` = [A, {key: band} of level 1 (finest), ..., level L], keys as in pywt.wavedecn ('aad' = low-pass along
    z and y, high-pass along x); the values equal pywt.wavedecn(vol, wname, mode='periodization', level=L)."""
    cdef pwt3_plan* w
    cdef readonly tuple shape
    cdef readonly str wname
    cdef readonly int levels
    cdef readonly list sizes        # (nz, ny, nx) of the bands of level 1 .. L

    def __cinit__(self, vol, str wname, int levels):
        self.w = NULL
        vol = np.ascontiguousarray(vol, dtype=np.float32)
        if vol.ndim != 3:
            raise ValueError("Wavelets3D(): a 3D volume is required, got %d dimensions" % vol.ndim)
        self.shape = tuple(int(x) for x in vol.shape)
        self.wname = wname
        if pwt_device_count() < 1:
            raise RuntimeError("pycudwt: no CUDA device available (there is no CPU fallback)")
        py_wname = wname.encode("ASCII")
        cdef const float* src = <const float*> <size_t> vol.ctypes.data
        cdef const char* c_wname = py_wname
        cdef int rc, nz = self.shape[0], ny = self.shape[1], nx = self.shape[2], c_levels = levels
        with nogil:
            rc = pwt3_create(&self.w, src, nz, ny, nx, c_wname, c_levels, 1)
        if rc != 0:
            self.w = NULL
            msg = _errmsg()
            if rc in (PWT_ERR_UNKNOWN_WAVELET, PWT_ERR_TOO_SMALL, PWT_ERR_UNSUPPORTED, -1):
                raise ValueError(msg)
            raise RuntimeError(msg)
        self.levels = pwt3_levels(self.w)
        self.sizes = []
        for l in range(1, self.levels + 1):
            pwt3_band_shape(self.w, l, &nz, &ny, &nx)
            self.sizes.append((nz, ny, nx))

    def _band(self, int level, int b):
        out = np.empty(self.sizes[level - 1], dtype=np.float32)
        cdef float* dst = <float*> <size_t> out.ctypes.data
        cdef int rc
        with nogil:
            rc = pwt3_get_coeff(self.w, dst, level, b)
        if rc != 0:
            raise RuntimeError(_errmsg())
        return out

    @property
    def coeffs(self):
        res = [self._band(self.levels, 0)]
        for l in range(1, self.levels + 1):
            res.append({k: self._band(l, b + 1) for b, k in enumerate(VOLUME_BAND_KEYS)})
        return res

    def set_coeff(self, arr, int level, key):
        b = 0 if key in (0, "aaa") else (VOLUME_BAND_KEYS.index(key) + 1 if isinstance(key, str) else int(key))
        arr = np.ascontiguousarray(arr, dtype=np.float32)
        if level < 1 or level > self.levels or tuple(arr.shape) != self.sizes[level - 1]:
            raise ValueError("Wavelets3D.set_coeff(): wrong level or shape")
        cdef const float* src = <const float*> <size_t> arr.ctypes.data
        cdef int rc, bb = b
        with nogil:
            rc = pwt3_set_coeff(self.w, src, level, bb, 0)
        if rc != 0:
            raise ValueError(_errmsg())

    @property
    def image(self):
        out = np.empty(self.shape, dtype=np.float32)
        cdef float* dst = <float*> <size_t> out.ctypes.data
        cdef int rc
        with nogil:
            rc = pwt3_get_image(self.w, dst)
        if rc != 0:
            raise RuntimeError(_errmsg())
        return out

    def set_image(self, vol):
        vol = np.ascontiguousarray(vol, dtype=np.float32)
        if tuple(vol.shape) != self.shape:
            raise ValueError("Wavelets3D.set_image(): shape mismatch %s != %s" % (vol.shape, self.shape))
        cdef const float* src = <const float*> <size_t> vol.ctypes.data
        cdef int rc
        with nogil:
            rc = pwt3_set_image(self.w, src, 0)
        if rc != 0:
            raise RuntimeError(_errmsg())

    def forward(self, vol=None):
        if vol is not None:
            self.set_image(vol)
        cdef int rc
        with nogil:
            rc = pwt3_forward(self.w)
        if rc != 0:
            raise RuntimeError(_errmsg())

    def inverse(self):
        cdef int rc
        with nogil:
            rc = pwt3_inverse(self.w)
        if rc < 0:
            raise RuntimeError(_errmsg())

    def soft_threshold(self, float beta, int do_threshold_appcoeffs=0):
        if pwt3_soft_threshold(self.w, beta, do_threshold_appcoeffs) < 0:
            raise RuntimeError(_errmsg())

    def hard_threshold(self, float beta, int do_threshold_appcoeffs=0):
        if pwt3_hard_threshold(self.w, beta, do_threshold_appcoeffs) < 0:
            raise RuntimeError(_errmsg())

    def norms(self):
        cdef double a = 0, b = 0
        cdef int rc
        with nogil:
            rc = pwt3_norms(self.w, &a, &b)
        if rc != 0:
            raise RuntimeError(_errmsg())
        return a, b

    def norm1(self):
        return self.norms()[0]

    def norm2sq(self):
        return self.norms()[1]

    def sync(self):
        pwt3_sync(self.w)

    def timer_start(self):
        if pwt3_timer_start(self.w) != 0:
            raise RuntimeError(_errmsg())

    def timer_stop(self):
        cdef float ms = 0
        cdef int rc
        with nogil:
            rc = pwt3_timer_stop(self.w, &ms)
        if rc != 0:
            raise RuntimeError(_errmsg())
        return ms

    @property
    def launch_count(self):
        return int(pwt3_launch_count(self.w))

    def __dealloc__(self):
        if self.w is not NULL:
            pwt3_destroy(self.w)
            self.w = NULL
