"""Literal per-output emulation of the reference kernels' index arithmetic -- TEST INFRASTRUCTURE.

Pure-Python loops (small inputs only).  Each function walks the same window / wrap / tap-index
logic as the cited CUDA kernel, one "thread" (output sample) at a time, so that the vectorised
closed forms in `pdwt_oracle.py` can be pinned to the reference's actual arithmetic without a GPU.
References are relative to /root/reference/pdwt/src/.
"""
import numpy as np


def _centre_fwd(hlen):
    # separable.cu:98-107: odd -> hlen/2 ; even -> hlen/2 - 1.  Window length is hlen either way.
    return hlen // 2 if hlen & 1 else hlen // 2 - 1


def _wrap_dwt(idx, N):
    # separable.cu:116-121 (odd N: virtual extension by one repeated sample, period N+1)
    odd = N & 1
    if idx < 0:
        idx += N + odd
    if idx > N - 1:
        if idx == N and odd:
            idx -= 1
        else:
            idx -= N + odd
    return idx


def fwd_rows(img, fL, fH):
    """w_kern_forward_pass1, separable.cu:91-131."""
    Nr, Nc = img.shape
    hlen = len(fL)
    Nc2 = (Nc + (Nc & 1)) // 2
    c = _centre_fwd(hlen)
    lo = np.zeros((Nr, Nc2))
    hi = np.zeros((Nr, Nc2))
    for y in range(Nr):
        for x in range(Nc2):
            for j in range(hlen):
                v = img[y, _wrap_dwt(2 * x - c + j, Nc)]
                lo[y, x] += v * fL[hlen - 1 - j]
                hi[y, x] += v * fH[hlen - 1 - j]
    return lo, hi


def fwd_cols(t1, t2, fL, fH):
    """w_kern_forward_pass2, separable.cu:135-176."""
    Nr, Nc = t1.shape
    hlen = len(fL)
    Nr2 = (Nr + (Nr & 1)) // 2
    c = _centre_fwd(hlen)
    out = [np.zeros((Nr2, Nc)) for _ in range(4)]
    for y in range(Nr2):
        for x in range(Nc):
            for j in range(hlen):
                iy = _wrap_dwt(2 * y - c + j, Nr)
                out[0][y, x] += t1[iy, x] * fL[hlen - 1 - j]
                out[1][y, x] += t1[iy, x] * fH[hlen - 1 - j]
                out[2][y, x] += t2[iy, x] * fL[hlen - 1 - j]
                out[3][y, x] += t2[iy, x] * fH[hlen - 1 - j]
    return out


def _inv_geometry(hlen, g):
    # separable.cu:250-267 / 297-314: half-kernel parity decides centre and the +1 "virtual id"
    h2 = hlen // 2
    if h2 & 1:
        c, span, shift = h2 // 2, 2 * (h2 // 2), 0
    else:
        c, span, shift = h2 // 2, 2 * (h2 // 2) - 1, 1
    return c, span, shift, g + shift


def inv_rows(t1, t2, fIL, fIH, Nc_out):
    """w_kern_inverse_pass2, separable.cu:293-328 (Nc = coefficient width)."""
    Nr, Nc = t1.shape
    hlen = len(fIL)
    img = np.zeros((Nr, Nc_out))
    for y in range(Nr):
        for gx in range(Nc_out):
            c, span, shift, g = _inv_geometry(hlen, gx)
            j1 = c - g // 2
            j2 = Nc - 1 - g // 2 + c
            off = 1 - (g & 1)
            acc = 0.0
            for j in range(span + 1):
                ix = g // 2 - c + j
                if j < j1:
                    ix += Nc
                if j > j2:
                    ix -= Nc
                t = hlen - 1 - (2 * j + off)
                acc += t1[y, ix] * fIL[t] + t2[y, ix] * fIH[t]
            if g - shift < Nc_out:   # the store index is gidx (or gidx-1 after the shift)
                img[y, g - shift] = acc
    return img


def inv_cols(cA, cH, cV, cD, fIL, fIH, Nr_out):
    """w_kern_inverse_pass1, separable.cu:246-289."""
    Nr, Nc = cA.shape
    hlen = len(fIL)
    t1 = np.zeros((Nr_out, Nc))
    t2 = np.zeros((Nr_out, Nc))
    for gy in range(Nr_out):
        c, span, shift, g = _inv_geometry(hlen, gy)
        j1 = c - g // 2
        j2 = Nr - 1 - g // 2 + c
        off = 1 - (g & 1)
        for x in range(Nc):
            a1 = a2 = 0.0
            for j in range(span + 1):
                iy = g // 2 - c + j
                if j < j1:
                    iy += Nr
                if j > j2:
                    iy -= Nr
                t = hlen - 1 - (2 * j + off)
                a1 += cA[iy, x] * fIL[t] + cH[iy, x] * fIH[t]
                a2 += cV[iy, x] * fIL[t] + cD[iy, x] * fIH[t]
            t1[g - shift, x] = a1
            t2[g - shift, x] = a2
    return t1, t2


def _swt_wrap(g, j, c, factor, N):
    # separable.cu:431-438: one conditional wrap each side
    idx = g + j * factor - c
    if factor * j < c - g:
        idx += N
    if factor * j > N - 1 - g + c:
        idx -= N
    return idx


def swt_rows(img, fL, fH, level):
    """w_kern_forward_swt_pass1, separable.cu:409-449."""
    Nr, Nc = img.shape
    hlen = len(fL)
    factor = 1 << (level - 1)
    c = _centre_fwd(hlen) * factor
    lo = np.zeros((Nr, Nc))
    hi = np.zeros((Nr, Nc))
    for y in range(Nr):
        for x in range(Nc):
            for j in range(hlen):
                v = img[y, _swt_wrap(x, j, c, factor, Nc)]
                lo[y, x] += v * fL[hlen - 1 - j]
                hi[y, x] += v * fH[hlen - 1 - j]
    return lo, hi


def iswt_rows(t1, t2, fIL, fIH, level):
    """w_kern_inverse_swt_pass2, separable.cu:593-626 (taps divided by 2)."""
    Nr, Nc = t1.shape
    hlen = len(fIL)
    factor = 1 << (level - 1)
    if hlen & 1:
        c, span = hlen // 2, 2 * (hlen // 2)
    else:
        c, span = hlen // 2, 2 * (hlen // 2) - 1
    c *= factor
    img = np.zeros((Nr, Nc))
    for y in range(Nr):
        for x in range(Nc):
            acc = 0.0
            for j in range(span + 1):
                ix = _swt_wrap(x, j, c, factor, Nc)
                acc += t1[y, ix] * fIL[hlen - 1 - j] / 2 + t2[y, ix] * fIH[hlen - 1 - j] / 2
            img[y, x] = acc
    return img


def ns_filters(f_lo, f_hi):
    """w_compute_filters, nonseparable.cu:70-74: LL, LH(=lo (x) hi), HL, HH with res[i*len+j]=a[i]*b[j]."""
    return (np.outer(f_lo, f_lo), np.outer(f_lo, f_hi), np.outer(f_hi, f_lo), np.outer(f_hi, f_hi))


def ns_forward(img, K):
    """w_kern_forward, nonseparable.cu:114-171.  K = (LL, LH, HL, HH) analysis; returns (a, h, v, d)."""
    Nr, Nc = img.shape
    hlen = K[0].shape[0]
    Nr2, Nc2 = (Nr + (Nr & 1)) // 2, (Nc + (Nc & 1)) // 2
    c = _centre_fwd(hlen)
    out = [np.zeros((Nr2, Nc2)) for _ in range(4)]
    for y in range(Nr2):
        for x in range(Nc2):
            for jy in range(hlen):
                iy = _wrap_dwt(2 * y - c + jy, Nr)
                for jx in range(hlen):
                    v = img[iy, _wrap_dwt(2 * x - c + jx, Nc)]
                    for b in range(4):
                        out[b][y, x] += v * K[b][hlen - 1 - jy, hlen - 1 - jx]
    return out


def ns_inverse(cA, cH, cV, cD, K, shape):
    """w_kern_inverse, nonseparable.cu:176-225.  K = synthesis (LL, LH, HL, HH)."""
    Nr, Nc = cA.shape
    Nr2, Nc2 = shape
    hlen = K[0].shape[0]
    img = np.zeros(shape)
    for gy0 in range(Nr2):
        for gx0 in range(Nc2):
            c, span, shift, gy = _inv_geometry(hlen, gy0)
            gx = gx0 + shift
            ox, oy = 1 - (gx & 1), 1 - (gy & 1)
            acc = 0.0
            for jy in range(span + 1):
                iy = gy // 2 - c + jy
                if jy < c - gy // 2:
                    iy += Nr
                if jy > Nr - 1 - gy // 2 + c:
                    iy -= Nr
                for jx in range(span + 1):
                    ix = gx // 2 - c + jx
                    if jx < c - gx // 2:
                        ix += Nc
                    if jx > Nc - 1 - gx // 2 + c:
                        ix -= Nc
                    ty, tx = hlen - 1 - (2 * jy + oy), hlen - 1 - (2 * jx + ox)
                    acc += (cA[iy, ix] * K[0][ty, tx] + cH[iy, ix] * K[1][ty, tx]
                            + cV[iy, ix] * K[2][ty, tx] + cD[iy, ix] * K[3][ty, tx])
            img[gy - shift, gx - shift] = acc
    return img


def ns_forward_swt(img, K, level):
    """w_kern_forward_swt, nonseparable.cu:304-355."""
    Nr, Nc = img.shape
    hlen = K[0].shape[0]
    factor = 1 << (level - 1)
    c = _centre_fwd(hlen) * factor
    out = [np.zeros((Nr, Nc)) for _ in range(4)]
    for y in range(Nr):
        for x in range(Nc):
            for jy in range(hlen):
                iy = _swt_wrap(y, jy, c, factor, Nr)
                for jx in range(hlen):
                    v = img[iy, _swt_wrap(x, jx, c, factor, Nc)]
                    for b in range(4):
                        out[b][y, x] += v * K[b][hlen - 1 - jy, hlen - 1 - jx]
    return out


def ns_inverse_swt(cA, cH, cV, cD, K, level):
    """w_kern_inverse_swt, nonseparable.cu:360-401 (taps divided by 4)."""
    Nr, Nc = cA.shape
    hlen = K[0].shape[0]
    factor = 1 << (level - 1)
    c = (hlen // 2) * factor
    span = 2 * (hlen // 2) if hlen & 1 else 2 * (hlen // 2) - 1
    img = np.zeros((Nr, Nc))
    for y in range(Nr):
        for x in range(Nc):
            acc = 0.0
            for jy in range(span + 1):
                iy = _swt_wrap(y, jy, c, factor, Nr)
                for jx in range(span + 1):
                    ix = _swt_wrap(x, jx, c, factor, Nc)
                    ty, tx = hlen - 1 - jy, hlen - 1 - jx
                    acc += (cA[iy, ix] * K[0][ty, tx] + cH[iy, ix] * K[1][ty, tx]
                            + cV[iy, ix] * K[2][ty, tx] + cD[iy, ix] * K[3][ty, tx]) / 4
            img[y, x] = acc
    return img


def haar2d_fwd(img):
    """kern_haar2d_fwd, haar.cu:10-38."""
    Nr, Nc = img.shape
    Nr2, Nc2 = (Nr + (Nr & 1)) // 2, (Nc + (Nc & 1)) // 2
    out = [np.zeros((Nr2, Nc2)) for _ in range(4)]
    for y in range(Nr2):
        for x in range(Nc2):
            x1 = 2 * x + 1 if not (Nc & 1 and 2 * x + 1 == Nc) else 2 * x
            y1 = 2 * y + 1 if not (Nr & 1 and 2 * y + 1 == Nr) else 2 * y
            a, b, c, d = img[2 * y, 2 * x], img[2 * y, x1], img[y1, 2 * x], img[y1, x1]
            out[0][y, x] = 0.5 * ((a + c) + (b + d))
            out[2][y, x] = 0.5 * ((a + c) - (b + d))   # V
            out[1][y, x] = 0.5 * ((a - c) + (b - d))   # H
            out[3][y, x] = 0.5 * ((a - c) - (b - d))
    return out


def haar2d_inv(cA, cH, cV, cD, shape):
    """kern_haar2d_inv, haar.cu:41-58."""
    img = np.zeros(shape)
    for y in range(shape[0]):
        for x in range(shape[1]):
            a, b, c, d = cA[y // 2, x // 2], cV[y // 2, x // 2], cH[y // 2, x // 2], cD[y // 2, x // 2]
            sy = 1.0 if y & 1 == 0 else -1.0
            sx = 1.0 if x & 1 == 0 else -1.0
            img[y, x] = 0.5 * ((a + sy * c) + sx * (b + sy * d))
    return img
