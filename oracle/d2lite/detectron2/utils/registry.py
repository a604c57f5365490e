class Registry:
    """name -> object map with a decorator-style register() (fvcore.common.registry.Registry)."""

    def __init__(self, name):
        self._name = name
        self._obj_map = {}

    def _do_register(self, name, obj):
        assert name not in self._obj_map, f"'{name}' already registered in '{self._name}'"
        self._obj_map[name] = obj

    def register(self, obj=None):
        if obj is None:
            def deco(fn):
                self._do_register(fn.__name__, fn)
                return fn
            return deco
        self._do_register(obj.__name__, obj)
        return obj

    def get(self, name):
        if name not in self._obj_map:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return self._obj_map[name]

    def __contains__(self, name):
        return name in self._obj_map
