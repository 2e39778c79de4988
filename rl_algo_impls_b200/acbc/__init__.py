from .acbc import ACBC

__all__ = ["ACBC"]
