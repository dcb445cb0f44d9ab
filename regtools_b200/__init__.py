"""regtools_b200 — B200-native drop-in for the `regtools junctions extract` hot path.

Only what that path needs lives here: csrc/ (CUDA kernels for sm_100a, native BAM feeder, the C ABI
libregtools_jx.so, the C++ JunctionsExtractor shim and the `regtools` CLI) and the Python mirror of
the reference's JunctionsExtractor interface.  Importing this package loads the shared library and
raises if it has not been built — there is no fallback implementation.
"""
from . import _lib  # noqa: F401  (raises ImportError when libregtools_jx.so is missing)
from .extractor import (CmdlineHelpException, Junction, JunctionsAnnotator, JunctionsExtractor, JUNCTION_DTYPE,  # noqa: F401
                        junctions_annotate, junctions_extract, plan_shards)

__all__ = ["JunctionsExtractor", "JunctionsAnnotator", "Junction", "CmdlineHelpException", "junctions_extract",
           "junctions_annotate", "plan_shards", "JUNCTION_DTYPE"]
