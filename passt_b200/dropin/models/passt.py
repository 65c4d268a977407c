"""Drop-in for reference ``models/passt.py``: same ``model_ing`` ingredient + command names (:917-919, :922, :932,
:957, :1039), implementation from passt_b200."""
from passt_b200 import passt as _impl
from passt_b200.passt import PaSST, EnsembelerModel, checkpoint_filter_fn  # noqa: F401
from ._ingredient import make_ingredient

model_ing = make_ingredient("passt")
model_ing.add_config(instance_cmd="get_model")

fix_embedding_layer = model_ing.command(_impl.fix_embedding_layer)
lighten_model = model_ing.command(_impl.lighten_model)
get_model = model_ing.command(_impl.get_model)
get_ensemble_model = model_ing.command(_impl.get_ensemble_model)
