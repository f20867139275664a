"""TEST INFRASTRUCTURE ONLY -- the correctness oracle for the TLR update path.

Nothing in ``hcore_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs use it, and only as the checker or the timed CPU baseline.

* ``oracle.tlr_oracle``  -- numpy/SciPy-LAPACK restatement of the reference algorithm
* ``oracle.ref_binding`` -- ctypes binding of ``oracle/_ref/libhcore_ref.so``, the
  UNMODIFIED reference compiled by ``oracle/Makefile`` (present when it was built in
  the dev container; it travels to the GPU box as a prebuilt file)
"""
