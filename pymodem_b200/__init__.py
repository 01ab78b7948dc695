"""pymodem_b200: the demod_chain hot path of pymodem (demodulate -> slice -> unscramble -> decode) on B200 GPUs.

The compute lives in libpymodem_b200.so (pymodem_b200/csrc, C ABI in include/pymodem_b200.h); this package holds the
ctypes binding (engine.py), the host mirror of the reference's builder/execute interface (modems_codecs/) and the
multi-GPU host side (sharded.py).  There is no CPU fallback: without the CUDA library the engine raises."""
