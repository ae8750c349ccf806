"""Krylov factorizations (initialize / expand! / shrink!) — mirror of src/factorizations/."""
from .arnoldi import ArnoldiFactorization, ArnoldiIterator
from .gkl import GKLFactorization, GKLIterator
from .lanczos import LanczosFactorization, LanczosIterator

__all__ = ["LanczosIterator", "LanczosFactorization", "ArnoldiIterator", "ArnoldiFactorization",
           "GKLIterator", "GKLFactorization"]
