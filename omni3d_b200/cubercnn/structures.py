"""Minimal result containers with the field/attribute surface the reference's callers use
(detectron2.structures.Boxes / Instances as consumed by demo/demo.py:90-126 and the evaluator)."""
import torch


class Boxes:
    def __init__(self, tensor):
        self.tensor = tensor.reshape(-1, 4).float()

    def __len__(self):
        return self.tensor.shape[0]

    def __getitem__(self, i):
        return Boxes(self.tensor[i].reshape(-1, 4))

    def to(self, *a, **k):
        return Boxes(self.tensor.to(*a, **k))

    def area(self):
        t = self.tensor
        return (t[:, 2] - t[:, 0]) * (t[:, 3] - t[:, 1])

    def get_centers(self):
        return (self.tensor[:, :2] + self.tensor[:, 2:]) / 2


class Instances:
    def __init__(self, image_size, **fields):
        object.__setattr__(self, "_image_size", image_size)
        object.__setattr__(self, "_fields", {})
        for k, v in fields.items():
            self.set(k, v)

    @property
    def image_size(self):
        return self._image_size

    def set(self, name, value):
        self._fields[name] = value

    def __setattr__(self, name, value):
        self.set(name, value)

    def __getattr__(self, name):
        f = object.__getattribute__(self, "_fields")
        if name not in f:
            raise AttributeError(f"Cannot find field '{name}' in the given Instances!")
        return f[name]

    def has(self, name):
        return name in self._fields

    def get(self, name):
        return self._fields[name]

    def get_fields(self):
        return self._fields

    def __len__(self):
        for v in self._fields.values():
            return len(v)
        return 0

    def __getitem__(self, item):
        return Instances(self._image_size, **{k: v[item] for k, v in self._fields.items()})

    def to(self, *a, **k):
        return Instances(self._image_size, **{n: (v.to(*a, **k) if hasattr(v, "to") else v)
                                              for n, v in self._fields.items()})
