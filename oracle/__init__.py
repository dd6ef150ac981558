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
* signal preparation (normalize_signal, removeSilence, the doubling below
  0.1 s, mix_signals) and get_data_statistics -- pinned: the same script runs
  the reference's load_and_preprocess_signal / mix_signals and its compiled
  Cython removeSilence / get_data_statistics; signals, frame markers, sample
  markers and statistics are reproduced bit for bit by the restatement.
* STFT, softmask, Slaney mel basis, power_to_db -- **parity unpinned by
  librosa** (an un-vendored, unpinned dependency that is not installable here;
  the reference ships no tests or golden vectors).  These four rows -- and
  ``librosa.feature.rms`` inside the signal preparation, which the golden
  script shims with the restatement too -- are the ones no librosa run backs.
  They restate the published librosa (~0.8) algorithm and are pinned by what
  IS available: analytic known answers that do not pass through this package
  (tests/test_kat.py: impulse and bin-centred sinusoid spectra, exact rational
  soft masks, a hand-computed Slaney basis and the 6400 Hz = 42 mel anchor,
  power_to_db edge values), a whole-pipeline cross-check against
  ``transformers.audio_utils.spectrogram``, and per-leaf cross-checks against
  ``torch.stft`` in float64, ``torchaudio.functional.melscale_fbanks`` and
  ``transformers.audio_utils``.
"""
