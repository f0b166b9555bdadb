"""
Error types of the drop-in boundary (reference: tapqir/exceptions.py:8-39).  Same class names,
constructor arguments and messages so callers that catch them keep working.
"""

from pathlib import Path
from typing import Union


class TapqirException(Exception):
    """Root of the error hierarchy; ``msg`` must be non-empty."""

    def __init__(self, msg, *args):
        assert msg
        self.msg = msg
        super().__init__(msg, *args)


class TapqirFileNotFoundError(TapqirException):
    """A data / model / parameter / summary file expected at ``path`` is missing."""

    def __init__(self, name: str, path: Union[str, Path]):
        self.name = name
        self.path = path
        super().__init__(f"Unable to find {name} file '{path}'")


class CudaOutOfMemoryError(TapqirException):
    """The device ran out of memory; the fix is a smaller AOI / frame minibatch."""

    def __init__(self):
        super().__init__("CUDA out of memory. Try to use smaller AOI/frame batch size")


class NativeLibraryError(TapqirException):
    """The sm_100a shared library is missing or could not be loaded.  There is no CPU fallback."""
