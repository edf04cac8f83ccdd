"""Coefficients and error bound of the polynomial in fm_dev_fast (tfrec_b200/csrc/demod_dev.cuh).

atan(q) = q * R(q^2) on q in [0, 1]; R is interpolated at Chebyshev nodes (degree 12 in z = q^2).  Prints the
coefficients as hex doubles and the maximum error, in radians and in fm_dev output units (x 16384/pi), of the
double-precision Horner evaluation against numpy's arctan over 2,000,001 points.  The kernel sends every sample
whose scaled angle is within 1e-6 of a truncation boundary to the exact path, 30x this error.
"""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P


def R(z):
    q = np.sqrt(z)
    out = np.ones_like(z)
    m = q > 1e-8
    out[m] = np.arctan(q[m]) / q[m]
    return out


def main(deg=12):
    c = C.Chebyshev.interpolate(R, deg, domain=[0, 1]).convert(kind=P.Polynomial).coef
    for k, v in enumerate(c):
        print("z^%-2d %s" % (k, float(v).hex()))
    q = np.linspace(0, 1, 2000001)
    z = q * q
    acc = np.full_like(z, c[-1])
    for k in range(len(c) - 2, -1, -1):
        acc = acc * z + c[k]
    err = float(np.max(np.abs(q * acc - np.arctan(q))))
    print("max error %.3e rad = %.3e output units" % (err, err * 16384 / np.pi))


if __name__ == "__main__":
    main()
