"""d2lite restatement of the pytorch3d.transforms functions on the Cube R-CNN path."""
