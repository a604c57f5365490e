"""Name -> builder registries with the reference's registry names (rcnn3d.py:25, dla.py:484,
resnet.py:66, rpn.py:19, roi_heads.py:39, cube_head.py:17)."""


class Registry(dict):
    def __init__(self, name):
        super().__init__()
        self.name = name

    def register(self, obj=None, name=None):
        def deco(o):
            key = name or o.__name__
            if key in self:
                raise KeyError(f"{key} already registered in {self.name}")
            self[key] = o
            return o
        return deco if obj is None else deco(obj)

    def get(self, key):
        if key not in self:
            raise KeyError(f"No object named '{key}' found in '{self.name}' registry!")
        return self[key]


META_ARCH_REGISTRY = Registry("META_ARCH")
BACKBONE_REGISTRY = Registry("BACKBONE")
PROPOSAL_GENERATOR_REGISTRY = Registry("PROPOSAL_GENERATOR")
ROI_HEADS_REGISTRY = Registry("ROI_HEADS")
ROI_CUBE_HEAD_REGISTRY = Registry("ROI_CUBE_HEAD")
