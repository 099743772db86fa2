"""TEST INFRASTRUCTURE — the CPU oracle for the particle-stepping hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or the
timed CPU baseline.  The product (``ptoy_b200``) never imports it.

* ``oracle.ref``  — ctypes binding of ``oracle/_ref/libptoy_ref*.so``: the
  UNMODIFIED reference physics compiled from ``/root/reference/src`` (prebuilt
  in the build container; travels to the GPU box as a .so).
* ``oracle.port`` — ctypes binding of ``oracle/libptoy_oracle.so``: a plain-C
  restatement of the same algorithm (``oracle/ptoy_oracle.c``), pinned bit-for-bit
  against ``oracle.ref`` and against ``tests/golden/``.
"""
