"""CPU oracle for the eegldm hot path -- TEST INFRASTRUCTURE ONLY.

Plain-PyTorch (fp32, CPU) restatements of the reference algorithms on the
sampling / autoencoder path.  Nothing in the product package may import this
module: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker or the
timed CPU baseline.

Pinning status (see DESIGN.md "Oracle"):

* ``oracle.unet``      -- PINNED: checked bit-for-bit (max |diff| == 0 up to fp32
  summation order, asserted <= 1e-5) against the reference's own
  ``src/models/unet.py`` imported from ``/root/reference`` (tests/test_oracle_unet.py)
  and against committed golden vectors produced by that module
  (tests/golden/make_golden.py).
* ``oracle.schedulers`` -- known-answer scalars from SURVEY.md section 4 are
  asserted; the upstream package (monai-generative, version unpinned in
  ``requirements.txt:12``) is NOT installable here -> **parity unpinned** against
  upstream itself.
* ``oracle.aekl`` / ``oracle.jukebox`` -- restated from the published upstream
  algorithm (monai-generative ``generative/networks/nets/autoencoderkl.py``,
  ``generative/losses/spectral_loss.py``), cross-checked structurally against the
  in-tree ancestor ``src/models/ae_kl.py``; **parity unpinned** against upstream.
"""
