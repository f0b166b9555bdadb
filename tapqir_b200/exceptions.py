"""
Error types of the drop-in boundary.  The reference's callers catch these by name and read ``.msg`` / ``.name`` /
``.path`` (tapqir/exceptions.py:8-39, raised at models/model.py:233-235,341 and utils/dataset.py:219-221), so the class
names, constructor arguments, attributes and message texts are part of the contract; ``NativeLibraryError`` is this
repository's own (no CPU fallback exists, a missing sm_100a library is an error).
"""

__all__ = ["TapqirException", "TapqirFileNotFoundError", "CudaOutOfMemoryError", "NativeLibraryError",
           "NonFiniteParameterError", "PeerTimeoutError"]


class TapqirException(Exception):
    """Root of the hierarchy: carries a non-empty message in ``msg``."""

    def __init__(self, msg, *extra):
        if not msg:
            raise AssertionError("a TapqirException needs a message")
        Exception.__init__(self, msg, *extra)
        self.msg = msg


class TapqirFileNotFoundError(TapqirException):
    """A file of kind ``name`` ("data", "model", "parameter", "summary") expected at ``path`` does not exist."""

    template = "Unable to find {name} file '{path}'"

    def __init__(self, name, path):
        TapqirException.__init__(self, self.template.format(name=name, path=path))
        self.name, self.path = name, path


class CudaOutOfMemoryError(TapqirException):
    """The device ran out of memory during a step or while computing statistics."""

    message = "CUDA out of memory. Try to use smaller AOI/frame batch size"

    def __init__(self):
        TapqirException.__init__(self, self.message)


class NativeLibraryError(TapqirException):
    """libtapqir_b200.so (sm_100a) is missing or could not be loaded."""


class NonFiniteParameterError(ValueError):
    """NaN / Inf found in a variational parameter at checkpoint time.  The reference raises a plain ``ValueError`` there
    (models/model.py:246-250) and ``run`` answers it by restoring the last checkpoint with a new seed (:220-232); this
    subclass keeps callers that catch ``ValueError`` working while ``Model.run`` catches ONLY it -- an argument error of
    the C ABI is a ``ValueError`` too and must not be mistaken for a divergence."""


class PeerTimeoutError(TapqirException):
    """A rank stopped pushing its accumulators over NVLink peer memory (csrc/p2p_allreduce.cu gave up waiting): the sums
    of that step -- and every parameter update since -- are garbage."""

    def __init__(self, seq):
        TapqirException.__init__(self, f"peer-memory all-reduce timed out at call #{seq}: a rank died or stalled; "
                                       "parameters since the last checkpoint are not to be trusted")
        self.seq = seq
