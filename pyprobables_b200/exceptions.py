"""Exception types of the batch engine: same class names, hierarchy and `.message` attribute as
probables/exceptions.py:4-92 so that `except` clauses written for pyprobables keep working."""


class ProbablesBaseException(Exception):
    """root of every error this package raises on purpose"""

    def __init__(self, message: str) -> None:
        self.message = message
        super().__init__(message)

    def __str__(self) -> str:
        return self.message


def _make(name: str, doc: str):
    return type(name, (ProbablesBaseException,), {"__doc__": doc, "__module__": __name__})


InitializationError = _make("InitializationError", "a constructor was given unusable parameters")
NotSupportedError = _make("NotSupportedError", "the operation is not available on this variant")
SimilarityError = _make("SimilarityError", "two structures are not comparable (size or hash function differ)")
CuckooFilterFullError = _make("CuckooFilterFullError", "a fingerprint found no slot within max_swaps evictions")
RotatingBloomFilterError = _make("RotatingBloomFilterError", "rotating bloom filter queue error")
CountMinSketchError = _make("CountMinSketchError", "count-min sketch parameter mismatch")
QuotientFilterError = _make("QuotientFilterError", "quotient filter error")
