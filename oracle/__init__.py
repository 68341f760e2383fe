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
  * network forward (EfficientNet / BiFPN / heads; oracle/graph.py) : WIRING
    PINNED -- fixtures produced by executing the reference's own model.py /
    efficientnet.py / layers.py / initializers.py UNMODIFIED under a torch-backed
    stand-in for tensorflow.keras (tests/golden/keras_stub.py,
    make_golden_graph.py): topology, layer names, widths, skip / drop conditions,
    fusion-input order, head order, BN epsilons, activations, and the Keras layer
    counts of train_tpu.py:24.  The arithmetic INSIDE TensorFlow's kernels (SAME
    padding, BN formula, nearest upsampling, ...) stays a restatement of TF's
    documented behaviour: TensorFlow is not installable here.
  * losses (oracle/losses.py) and the fast-normalised fusion : PINNED on values and
    gradients produced by executing the reference's utils/tpu.py and
    layers.wBiFPNAdd under the same stand-in (make_golden_losses.py); only
    keras.backend.binary_crossentropy is restated (App. A.8).
  * top-k / pad tail, NMS tie order, SGD, BN training-mode statistics : PARITY
    UNPINNED -- TF semantics from SURVEY.md App. A; the reference has no tests or
    fixtures for them.
"""
