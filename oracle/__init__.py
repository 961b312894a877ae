"""CPU oracle for the audio front-end hot path.  TEST INFRASTRUCTURE ONLY.

This package is the *checker*, never the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.  Nothing under ``heart_murmur_detection_b200/``
imports it, and that package fails loudly when its CUDA library is missing.

What it restates (reference = /root/reference, carla-biermann/heart-murmur-detection):

* ``oracle.librosa_restated`` - numpy restatement of the librosa 0.10.1 calls the
  reference makes (``src/util.py:172,242,340,484-494,600``).  librosa itself is a
  third-party dependency pinned at 0.10.1 in ``environment.yml:117`` and is NOT
  installed in the build container, so this half of the oracle is
  **parity unpinned** against librosa itself; it is cross-checked against
  torchaudio's librosa-compatible ops (``tests/test_oracle.py``).
* ``oracle.frontend`` - restatement of the reference's own Python
  (``src/util.py:26-51,113-126,481-620,794-860`` and
  ``src/benchmark/baseline/extract_feature.py:213-259``) on in-memory arrays.
  Pinned: ``tests/golden/make_golden.py`` executes the *real* reference
  ``src/util.py`` (with ``librosa`` shimmed by ``oracle.librosa_restated``) and the
  committed fixtures are compared with this restatement.
* live third-party oracles: ``scipy.signal.butter/lfilter/sosfiltfilt``,
  ``torchaudio.compliance.kaldi.fbank`` and ``torchaudio.functional.resample``
  (the same libraries the reference calls).
"""
