"""CPU oracle for the EfficientDet hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Everything under ``oracle/`` is a plain CPU restatement (numpy / torch-CPU / C)
of the algorithm the reference (Ely-S/EfficientDet, mounted read-only at
/root/reference while this repo was built) runs for the path named in
BASELINE.json.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and
only as the checker or the timed CPU baseline -- never as a fallback for
``efficientdet_b200`` (which fails loudly when its CUDA library is missing).

Pinning status (see DESIGN.md "Oracle"):
  * anchors / IoU / anchor targets : PINNED -- bit-exact against vectors
    produced by executing the reference's own utils/anchors.py and the
    re-cythonized utils/compute_overlap.pyx (tests/golden/make_golden.py).
  * decode / clip / score-threshold+NMS : PINNED on the reference's own test
    vectors (test_RegressBoxes.py, test_ClipBoxes.py, test_FilterDetections.py).
  * network forward (EfficientNet / BiFPN / heads), top-k/pad tail, losses,
    SGD : PARITY UNPINNED -- TensorFlow is not installable here and the
    reference has no tests or fixtures for them; the restatement follows the
    cited lines plus the documented TF semantics listed in SURVEY.md App. A.
"""
