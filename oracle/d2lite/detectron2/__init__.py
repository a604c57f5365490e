"""d2lite: ORACLE restatement of the detectron2 API surface used by cubercnn (see ../README.md)."""
__version__ = "0.6-d2lite"
