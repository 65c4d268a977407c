"""Put this directory first on sys.path to make ``import models.passt`` / ``import models.preprocess`` resolve to the
B200 implementation, so that the reference's ex_audioset.py (which wires ``models.passt.model_ing`` and
``models.preprocess.model_ing`` by dotted path, ex_audioset.py:61-70) runs unchanged.  See INTEGRATION.md."""
