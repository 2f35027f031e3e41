"""Test / bench tooling for outline_rag_b200 -- NOT part of the product.

* ``orx_testkit.synth``   the counter-based synthetic "bge-m3-shaped" table and query generator
  (NumPy).  Importing it loads no native library, so the CPU reference arm of
  ``bench.py`` runs without ``liborx.so`` in its process.
* ``orx_testkit.device``  the bit-identical CUDA generator (``liborx_synth.so``, built by
  ``orx_testkit/Makefile``) that fills HBM buffers for the 1M..100M-row benches and tests.
"""
