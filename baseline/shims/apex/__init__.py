"""Import stand-in for NVIDIA apex (cannot be built offline).  The reference's ``optimization.py:24-34`` imports
``FusedAdam`` / ``multi_tensor_applier`` / ``amp_C`` at module level and never calls them; its
``BertLayerNorm`` falls back to its own Python layer norm when ``apex.normalization`` is not importable -- which is what
happens here, because this package deliberately has no ``normalization`` sub-module."""
