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
  (SURVEY.md section 4).  These restatements follow the cited reference lines
  and the published semantics of the TF ops they call; they are **parity
  unpinned** against a running TF and say so in their module headers.

Every function cites the reference file:line it restates (paths relative to the
reference repository root).
"""
