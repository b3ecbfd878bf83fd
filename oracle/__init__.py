"""TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy / cv2 / scipy, the reference's own numeric libraries) of the
rain-rendering hot path of astra-vision/rain-rendering, plus the harness that runs the
untouched reference in this container to produce golden vectors.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import this package -- and only as the checker / baseline.  The
product (``rain_rendering_b200``) never imports it and has no CPU fallback.

Parity pin: see ``oracle/README.md`` -- the restatement is bit-exact against the reference's
own Python on the committed golden inputs (``tests/golden``), with the third-party packages
that are absent here (pyclipper, imutils) restated from their published behaviour
("parity unpinned" for those two call sites).
"""
