"""Shadow of the reference's top-level package: same module paths and names,
backed by the B200 implementation.  Put this repository ahead of the reference
checkout on ``sys.path`` and the reference's ``utilities/`` and ``examples/``
drive the CUDA controller unchanged."""
