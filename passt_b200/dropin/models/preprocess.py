"""Drop-in for reference ``models/preprocess.py``: ``model_ing`` ("spectrograms") with the ``AugmentMelSTFT`` command
(:10, :18-19)."""
from passt_b200.preprocess import AugmentMelSTFT as _AugmentMelSTFT
from ._ingredient import make_ingredient

model_ing = make_ingredient("spectrograms")
AugmentMelSTFT = model_ing.command(_AugmentMelSTFT)
