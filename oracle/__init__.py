"""CPU oracle for the style-transfer hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in numpy / SciPy-BLAS, the arithmetic that the reference
(crowsonkb/style_transfer) performs on its per-tile hot path.  It exists to *check* the CUDA
engine in ``style_transfer_b200``; it is never the thing that is shipped or measured as the
product.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product package never does.

PARITY STATUS
-------------
* ``numeric.py`` and ``optimizers.py`` restate ``num_utils.py`` / ``optimizers.py`` of the
  reference and ARE pinned: ``tests/golden/make_golden.py`` imported the reference's own modules
  (with import stubs for the absent ``pywt`` / ``average`` packages) and wrote
  ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks the restatement against them.
* ``caffe_ops.py`` / ``caffe_net.py`` restate BVLC Caffe (unpinned ``git clone --depth 1`` at
  reference ``docker/Dockerfile:29``; source NOT under /root/reference, not installable here).
  For that part **parity is unpinned** against Caffe itself; it is anchored instead on the
  reference's call sites (``style_transfer.py:421-427, 556-612``), on per-op finite differences and
  on a cross-check against ``torch.nn.functional`` on CPU (``tests/test_oracle_caffe_ops.py``).
* ``average.EWMA`` (third-party, ``requirements.txt:3``, absent) is restated from its published
  behaviour in ``optimizers.py``; pinned only through the reference's own optimizer control flow.
"""
