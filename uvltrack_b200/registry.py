"""Name -> builder registries with the reference's names (lib/registry.py:6-49): ``MODELS['uvltrack']``,
``BACKBONES['modality_unified_feature_extractor']``, ``HEADS['modality_adaptive_box_head']``."""


class Registry(dict):
    """dict with a ``register(name)`` decorator / ``register(name, obj)`` call, duplicate names rejected."""

    def register(self, name, obj=None):
        def _add(o):
            if name in self:
                raise AssertionError(f"{name!r} is already registered")
            self[name] = o
            return o

        if obj is None:
            return _add
        _add(obj)
        return None


MODELS = Registry()
BACKBONES = Registry()
HEADS = Registry()
