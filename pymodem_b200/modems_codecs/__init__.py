"""Host-side mirror of the reference's modems_codecs package for the
demod_chain path: same module and class names, same constructor / option
semantics, but the blocks only hold parameters and filter taps -- the work is
done by the CUDA engine (pymodem_b200.engine)."""
