"""passt_b200 — B200-native (sm_100a) hot path of PaSST: waveform -> log-mel -> patchout-ViT fwd/bwd.

Public surface mirrors the reference (kkoutini/PaSST):
  passt_b200.preprocess.AugmentMelSTFT      (models/preprocess.py:19)
  passt_b200.passt.get_model / PaSST        (models/passt.py:957, :383)
  passt_b200.wrapper.get_basic_model / get_model_passt   (hear21passt-style, README.md:49-85)
The compute lives in the C-ABI library built from passt_b200/csrc (include/passt_b200.h).
"""
__version__ = "0.1.0"
