"""PLONK / fflonk polynomial wrappers — the only places those provers touch the bulk curve operations
(SURVEY.md §8a a10).  Mirrors of src/polynomial/polynomial.js and src/polynomial/evaluations.js for exactly those
methods; everything else in those classes is per-element host arithmetic and out of scope.

    Polynomial.fromEvaluations      polynomial.js:31-35     Fr.ifft
    Polynomial.to4T                 polynomial.js:111-119   Fr.ifft -> zero-pad x4 -> Fr.fft   (blinding handled by caller)
    Polynomial.multiExponentiation  polynomial.js:970-977   Fr.batchFromMontgomery -> G1.multiExpAffine -> toAffine
    Evaluations.fromPolynomial      evaluations.js:29-36    zero-pad x extension -> Fr.fft
"""
from __future__ import annotations

import numpy as np

from .curve import Curve, _arr


class Polynomial:
    def __init__(self, coefficients, curve: Curve, logger=None):
        self.coef = _arr(coefficients)
        self.curve = curve
        self.Fr = curve.Fr
        self.G1 = curve.G1
        self.logger = logger

    @staticmethod
    def fromEvaluations(buffer, curve: Curve, logger=None) -> "Polynomial":
        return Polynomial(curve.Fr.ifft(buffer), curve, logger)

    @staticmethod
    def to4T(buffer, domainSize: int, Fr):
        """-> (coefficients a, evaluations A4 over the 4x extended domain) without blinding factors."""
        a = Fr.ifft(buffer)
        a4 = np.zeros(domainSize * 4 * Fr.n8, np.uint8)
        a4[:a.size] = a
        return a, Fr.fft(a4)

    def length(self) -> int:
        return self.coef.size // self.Fr.n8

    def multiExponentiation(self, PTau, name: str = ""):
        n = self.coef.size // self.Fr.n8
        PTauN = _arr(PTau)[:n * self.G1.n8 * 2]
        bm = self.Fr.batchFromMontgomery(self.coef)
        res = self.G1.multiExpAffine(PTauN, bm, self.logger, name)
        return self.G1.toAffine(res)


class Evaluations:
    def __init__(self, evaluations, curve: Curve, logger=None):
        self.eval = _arr(evaluations)
        self.curve = curve
        self.Fr = curve.Fr
        self.logger = logger

    @staticmethod
    def fromPolynomial(polynomial: Polynomial, extension: int, curve: Curve, logger=None) -> "Evaluations":
        coefficientsN = np.zeros(polynomial.length() * extension * curve.Fr.n8, np.uint8)
        coefficientsN[:polynomial.coef.size] = polynomial.coef
        return Evaluations(curve.Fr.fft(coefficientsN), curve, logger)
