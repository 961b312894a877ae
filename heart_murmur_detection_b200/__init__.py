"""heart_murmur_detection_b200 - B200-native audio front-end for heart-murmur-detection.

A from-scratch sm_100a implementation of the reference's preprocessing / spectrogram hot
path (resample, Butterworth band-pass, silence trim, pad/split/crop, STFT power, mel,
power_to_db + min-max, Kaldi fbank) behind a C ABI (``include/hmfe.h``).

* ``frontend``  - batched, device-resident API (ragged batches of clips on the GPU)
* ``util`` / ``extract_feature`` - drop-in mirrors of the reference's
  ``src/util.py`` / ``src/benchmark/baseline/extract_feature.py`` call signatures
* ``build``     - in-tree nvcc build of ``libhmfe.so``

There is no CPU fallback anywhere in this package.
"""
__version__ = "0.1.0"
