"""B200-native CSR5 SpMV behind the reference's ``anonymouslibHandle<>`` API.

* ``handle.anonymouslibHandle`` -- ctypes mirror of the reference class on top of ``libcsr5_b200.so``
  (C ABI: ``include/csr5_b200.h``; CUDA sources: ``csrc/``);
* ``sharded.ShardedCsr5`` -- row-range sharded multi-GPU handle (one process per GPU);
* ``mmio`` / ``cli`` -- Matrix-Market ingest and the ``./spmv file.mtx`` report of the reference;
* ``matrices`` -- synthetic CSR generators for the benchmark configurations.

Nothing here falls back to a CPU implementation: without the CUDA library the handle raises
``_lib.Csr5LibraryMissing``.
"""
__all__ = ["handle", "sharded", "mmio", "cli", "matrices"]
