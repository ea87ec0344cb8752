"""CPU oracle for the SSD hot path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy / torch-CPU restatement of the reference algorithms
(pierluigiferrari/ssd_keras).  It exists to CHECK the CUDA path; it is never
the thing measured or shipped.  Only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
The product package ``ssd_keras_b200`` never imports ``oracle``.

Parity pins
-----------
* NumPy half (box math, matching, anchors, SSDInputEncoder, decode_detections,
  decode_detections_fast): the restatement is pinned against outputs of the
  REAL reference code imported from ``/root/reference`` in the build container.
  The generating script is ``tests/golden/make_golden.py``; its outputs are the
  committed fixtures ``tests/golden/*.npz`` / ``*.json``.
* TF/Keras half (AnchorBoxes, L2Normalization, DecodeDetections(+Fast) layers,
  SSDLoss, model graphs): TensorFlow 1.x / Keras 2.x cannot be installed here,
  and the reference has no tests or golden vectors of its own for them
  (SURVEY.md section 4).  They cannot be checked against a running TensorFlow;
  they ARE pinned against the reference's own source: ``tests/golden/make_tf_golden.py``
  imports the reference's loss / layer classes and model builders unmodified and
  executes them eagerly over ``tests/golden/tf_shim.py``, a NumPy/torch-CPU stand-in
  for the TensorFlow / Keras primitives they call; the outputs are the committed
  fixture ``tests/golden/ref_tf_shim_golden.npz`` (tests/test_oracle_tf_shim_golden.py).
  What stays assumed is the semantics of those primitives (listed in tf_shim.py).

Every function cites the reference file:line it restates (paths relative to the
reference repository root).
"""
