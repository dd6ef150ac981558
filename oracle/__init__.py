"""CPU oracle for the SM_HPSS_MTL feature front-end.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker / the CPU
arm being timed.  The product package ``sm_hpss_mtl_b200`` never imports it.

Pinning status (see DESIGN.md "Oracle"):

* median filters      -- pinned: ``scipy.ndimage.median_filter`` is the very
  routine ``librosa.decompose.hpss`` calls; our restatement is checked
  bit-exact against it.
* patches / scale_data -- pinned: the reference's own Cython leaf
  (``lib/cython_impl/tools.pyx``) is compiled into ``oracle/_ref`` and the
  restatement is checked against it.
* glue (feature-name dispatch, H/P concat, per-file StandardScaler, global
  stats) -- pinned: ``tests/golden/make_golden.py`` runs the reference's own
  ``lib/preprocessing.py`` (imported from /root/reference with a ``librosa``
  shim made of the leaf restatements below) and commits its outputs.
* STFT, softmask, Slaney mel basis, power_to_db -- **parity unpinned** by the
  reference (librosa is an un-vendored, unpinned dependency that is not
  installable here and the reference ships no tests or golden vectors); they
  restate the published librosa (~0.8) algorithm and are cross-checked against
  independent implementations available in the image (``torch.stft`` in
  float64, ``torchaudio.functional.melscale_fbanks``,
  ``transformers.audio_utils``).
"""
