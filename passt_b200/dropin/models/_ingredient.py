"""``model_ing`` provider: the real ba3l Ingredient when the reference's experiment stack (sacred, munch, ba3l) is
importable, otherwise a minimal stand-in whose ``command`` leaves the factories callable as plain functions
(the documented standalone use, README.md:289-292)."""


def make_ingredient(name):
    try:
        from ba3l.ingredients.ingredient import Ingredient  # reference ba3l/ingredients/ingredient.py:18
        return Ingredient(name)
    except Exception:
        class _Standalone:
            def __init__(self, path):
                self.path = path
                self.commands = {}
                self.config = {}

            def command(self, fn=None, **kw):
                def reg(f):
                    self.commands[getattr(f, "__name__", str(f))] = f
                    return f
                return reg(fn) if fn is not None else reg

            def add_config(self, *a, **kw):
                self.config.update(kw)

        return _Standalone(name)
