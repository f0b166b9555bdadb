"""
Error types of the drop-in boundary.  The reference's callers catch these by name and read ``.msg`` / ``.name`` /
``.path`` (tapqir/exceptions.py:8-39, raised at models/model.py:233-235,341 and utils/dataset.py:219-221), so the class
names, constructor arguments, attributes and message texts are part of the contract; ``NativeLibraryError`` is this
repository's own (no CPU fallback exists, a missing sm_100a library is an error).
"""

__all__ = ["TapqirException", "TapqirFileNotFoundError", "CudaOutOfMemoryError", "NativeLibraryError"]


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
