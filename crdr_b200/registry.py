"""Name -> class registries with the reference's names (src/utils/registry.py:80-95): classes register
themselves by ``__name__`` with ``@X_REGISTRY.register()`` and configs select them by ``type:``."""


class Registry:
    def __init__(self, name):
        self._name = name
        self._classes = {}

    def register(self):
        def wrap(obj):
            key = obj.__name__
            if key in self._classes:
                raise AssertionError(f"An object named '{key}' was already registered in '{self._name}' registry!")
            self._classes[key] = obj
            return obj
        return wrap

    def get(self, class_name, display_name=None):
        try:
            return self._classes[class_name]
        except KeyError:
            raise KeyError(f"No object named '{class_name}' found in '{self._name}' registry!") from None

    def __contains__(self, name):
        return name in self._classes

    def keys(self):
        return list(self._classes)


TRAINER_REGISTRY = Registry("trainer")
OPTIMIZER_REGISTRY = Registry("optimizer")
SCHEDULER_REGISTRY = Registry("scheduler")
MODEL_REGISTRY = Registry("comp_model")
ENCODER_REGISTRY = Registry("encoder")
DECODER_REGISTRY = Registry("decoder")
HYPERENCODER_REGISTRY = Registry("hyperencoder")
HYPERDECODER_REGISTRY = Registry("hyperdecoder")
CONTEXTMODEL_REGISTRY = Registry("context_model")
ENTROPYMODEL_REGISTRY = Registry("entropy_model")
DISCRIMINATOR_REGISTRY = Registry("discriminator")
DATASET_REGISTRY = Registry("dataset")
LOSS_REGISTRY = Registry("loss")
METRIC_REGISTRY = Registry("metric")
